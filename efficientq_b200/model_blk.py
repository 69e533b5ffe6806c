"""The 3D residual U-Net that drives the calibration path layer by layer.

Re-statement of the reference graph (src/models/model_blk.py:49-207 with the blocks of
src/models/factoryQ.py:66-81,182-236 and src/models/factory_blk.py:18-166) with
IDENTICAL module names, so reference checkpoints (``conv0.conv.weight``,
``u_blocks.UResBlock1.Layer1.block1.conv.weight``, ``trans_ups.TransUp4.upsampler.block.conv``
...) load unchanged.  Every conv is created through the ``QConv`` factory -- the
reference's plug-in point (src/definer.py:286-329).  The glue between quantizer layers
(ReLU, MaxPool3d, trilinear Upsample, residual add) is stock PyTorch by design
(out of scope for custom kernels).
"""
from __future__ import annotations

from collections.abc import Iterable

import torch
import torch.nn as nn

__all__ = ["UResQ", "get_conv_wrapper"]


class PassModule(nn.Module):
    def forward(self, x):
        return x


def ReLU(inplace=True):
    def make(inp=None):
        return nn.ReLU(inplace if inp is None else inp)
    return make


def get_conv_wrapper(QConv, q_weight, qlvl, q_act, qlvl_act, **kw):
    """factoryQ.py:182-192: curry the quantisation arguments into the conv constructor."""
    if QConv in (nn.Conv2d, nn.Conv3d):
        return QConv

    def make(in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        return QConv(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                     q_weight=q_weight, qlvl=qlvl, q_act=q_act, qlvl_act=qlvl_act, **kw)
    return make


class _Unit(nn.Module):
    """One conv with its norm / activation / dropout in one of three orders
    (factoryQ.py:30-81): pre = BN-ReLU-(DO)-Conv, mid = ReLU-(DO)-Conv-BN, post = (DO)-Conv-BN-ReLU."""
    ORDER = {"pre": ("bn", "relu", "do", "conv"), "mid": ("relu", "do", "conv", "bn"),
             "post": ("do", "conv", "bn", "relu")}

    def __init__(self, kind, cin, cout, kernel, stride, padding, dilation, Conv, bn, nla, drop_rate):
        super().__init__()
        self.kind = kind
        mods = {"relu": nla(), "do": nn.Dropout3d(drop_rate) if drop_rate > 0 else PassModule(),
                "conv": Conv(cin, cout, kernel, stride, padding, dilation, 1, False),
                "bn": bn(cin if kind == "pre" else cout)}
        for key in self.ORDER[kind]:                  # registration order = execution order (BN folding relies on it)
            setattr(self, key, mods[key])

    def forward(self, x):
        for key in self.ORDER[self.kind]:
            x = getattr(self, key)(x)
        return x


class ResBlockWithType(nn.Module):
    """factory_blk.py:147-166."""

    def __init__(self, cin, cout, drop_rate, dilation, nla, Conv, bn, blk_type):
        super().__init__()
        self.change_dim = cin != cout
        self.block1 = _Unit(blk_type, cin, cout, 3, 1, dilation, dilation, Conv, bn, nla, 0)
        self.block2 = _Unit(blk_type, cout, cout, 3, 1, dilation, dilation, Conv, bn, nla, drop_rate)
        self.projection = Conv(cin, cout, 1, 1, 0, bias=False) if self.change_dim else PassModule()

    def forward(self, x):
        return self.block2(self.block1(x)) + self.projection(x)


def _down(kernel, Conv, nla, bn, blk_type):
    """factory_blk.py:18-42: MaxPool then a 1x1x1 unit."""
    def make(cin, cout):
        seq = nn.Sequential()
        seq.add_module("pool", nn.MaxPool3d(kernel, kernel))
        seq.add_module("block", _Unit(blk_type, cin, cout, 1, 1, 0, 1, Conv, bn, nla, 0))
        return seq
    return make


def _up(scale, Conv, nla, bn, blk_type):
    """factory_blk.py:45-69: 1x1x1 unit (if the width changes) then trilinear upsampling."""
    def make(cin, cout):
        seq = nn.Sequential()
        if cin != cout:
            seq.add_module("block", _Unit(blk_type, cin, cout, 1, 1, 0, 1, Conv, bn, nla, 0))
        seq.add_module("trilinear", nn.Upsample(scale_factor=scale, mode="trilinear"))
        return seq
    return make


class _Fuser(nn.Module):
    """factory_blk.py:72-93: upsample and add the skip connection."""

    def __init__(self, upsampler):
        super().__init__()
        self.upsampler = upsampler

    def forward(self, x, skip):
        return self.upsampler(x) + skip


def _scale_tuple(t, f):
    return tuple(v * f for v in t) if isinstance(t, Iterable) else t * f


class ModelQ(nn.Module):
    """model_blk.py:17-37."""

    def load_state_dict(self, state_dict, strict=True, init=True):
        res = super().load_state_dict(state_dict, strict)
        if init:
            self.qparam_init()
        return res

    def qparam_init(self):
        for m in self.modules():
            if "QConv" in m.__class__.__name__ and hasattr(m, "qparam_init"):
                m.qparam_init()

    def perform_quantization(self):
        for m in self.modules():
            if "QConv" in m.__class__.__name__:
                m.perform_quantization()


class UResQ(ModelQ):
    def __init__(self, QConv, num_mod, num_classes, depth_config, width_config, dilation_config,
                 init_stride=1, stride=2, drop_rate=0.25, nla=ReLU(True), bn=nn.BatchNorm3d, ds=False,
                 blk_type="pre", q_weight=True, qlvl=8, q_act=True, qlvl_act=8, q_first=None, q_last=None,
                 rb=None, hetero_param=None, save_mem=False, fuse_bn=False, init_kernel=3, is_infer=False, **kwQ):
        super().__init__()
        assert len(depth_config) == len(width_config) == len(dilation_config)
        assert len(depth_config) % 2 == 1, "Can only have odd number of UBlocks"
        hp = hetero_param or {}
        ani_depth = hp.get("aniso_pool_depth", 99999)
        ani_stride = hp.get("aniso_pool_stride", (2, 2, 1))
        drop_cut = hp.get("drop_cut_thres", -1)
        ds_limit = hp.get("ds_depth_limit", 99999)
        self.init_stride = init_stride
        nb = len(depth_config)

        ConvQ = get_conv_wrapper(QConv, q_weight, qlvl, q_act, qlvl_act, **kwQ)
        up_nla = ReLU(False) if blk_type == "mid" else nla
        ConvFirst = get_conv_wrapper(QConv, q_first[0] > 0, q_first[0], q_first[1] > 0, q_first[1], **kwQ) \
            if q_first else nn.Conv3d
        ConvLast = get_conv_wrapper(QConv, q_last[0] > 0, q_last[0], q_last[1] > 0, q_last[1], **kwQ) \
            if q_last else nn.Conv3d

        self.conv0 = nn.Sequential()
        self.conv0.add_module("conv", ConvFirst(num_mod, width_config[0], init_kernel, init_stride,
                                                 (init_kernel - 1) // 2, bias=False))
        if blk_type != "pre":
            self.conv0.add_module("bn", bn(width_config[0]))
        if blk_type == "post":
            self.conv0.add_module("relu", nla())

        self.u_blocks, self.trans_downs = nn.Sequential(), nn.Sequential()
        self.trans_ups, self.classifiers = nn.Sequential(), nn.Sequential()
        for i in range(nb):
            dr = drop_rate
            if dr > 0 and width_config[i] < drop_cut:
                dr = min(drop_rate / 2, 0.2)
            stage = nn.Sequential() if depth_config[i] > 0 else PassModule()
            for j in range(depth_config[i]):
                stage.add_module(f"Layer{j + 1}", ResBlockWithType(width_config[i], width_config[i], dr,
                                                                   dilation_config[i], nla, ConvQ, bn, blk_type))
            self.u_blocks.add_module(f"UResBlock{i + 1}", stage)
            if i < nb // 2:
                k = stride if i < ani_depth else ani_stride
                self.trans_downs.add_module(f"TransDown{i + 1}",
                                            _down(k, ConvQ, nla, bn, blk_type)(width_config[i], width_config[i + 1]))
            elif i < nb - 1:
                k = stride if i >= nb - 1 - ani_depth else ani_stride
                self.trans_ups.add_module(f"TransUp{i + 1}", _Fuser(
                    _up(k, ConvQ, up_nla, bn, blk_type)(width_config[i], width_config[i + 1])))
                if ds:
                    if ds != "simple":
                        raise NotImplementedError("only ds='simple' (both reference PTQ configs) is supported")
                    head = None
                    if nb - i <= ds_limit:
                        head = nn.Sequential()
                        head.add_module("classifier", nn.Conv3d(width_config[i], num_classes, 1, 1, 0))
                        extra = _scale_tuple(init_stride, 2 ** len(width_config[i + 1:]))
                        if extra not in (1, (1, 1), (1, 1, 1)):
                            head.add_module("extra_up", nn.Upsample(scale_factor=extra, mode="trilinear"))
                    self.classifiers.add_module(f"AuxClassifier{i + 1}", head)

        self.final_cls = nn.Sequential()
        self.final_cls.add_module("cls", ConvLast(width_config[-1], num_classes, 1, 1, 0))
        if init_stride not in (1, (1, 1), (1, 1, 1)):
            self.final_cls.add_module("extra_up", nn.Upsample(scale_factor=init_stride, mode="trilinear"))

    def forward(self, x, feature_out=False):
        nb, nd = len(self.u_blocks), len(self.trans_downs)
        f = self.conv0(x)
        skips, outs = [], []
        for i in range(nb):
            f = self.u_blocks[i](f)
            if i < nd:
                skips.append(f)
                f = self.trans_downs[i](f)
            elif i < nb - 1:
                if len(self.classifiers) and self.classifiers[i - nd] is not None:
                    outs.append(self.classifiers[i - nd](f))
                f = self.trans_ups[i - nd](f, skips[-(i - nd + 1)])
        if feature_out:
            return f
        outs.append(self.final_cls(f))
        return torch.stack(outs, dim=0)
