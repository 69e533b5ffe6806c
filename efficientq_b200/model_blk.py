"""The 3D residual U-Net that drives the calibration path layer by layer.

Re-statement of the reference graph (src/models/model_blk.py:49-207 with the blocks of
src/models/factoryQ.py:66-81,182-236 and src/models/factory_blk.py:18-166) with
IDENTICAL module names, so reference checkpoints (``conv0.conv.weight``,
``u_blocks.UResBlock1.Layer1.block1.conv.weight``, ``trans_ups.TransUp4.upsampler.block.conv``
...) load unchanged.  Every conv is created through the ``QConv`` factory -- the
reference's plug-in point (src/definer.py:286-329).  The glue between quantizer layers
(ReLU, MaxPool3d, trilinear Upsample, residual add) runs on the
repo's own kernels (csrc/glue.cu) whenever the tensors live on the GPU and autograd is not
recording: the modules below subclass the stock ones (same names, no parameters, so the
state-dict layout is unchanged) and fuse the neighbouring elementwise op into the same pass
(MaxPool + the ReLU of the unit behind it, Upsample + the skip connection).  CPU tensors
(model definition / FP checkpoints on the host) and autograd (FP training, alpha refinement)
keep the stock PyTorch ops; ``EFFQ_GLUE=lib`` restores them on the GPU for A/B comparisons.
"""
from __future__ import annotations

import os
from collections.abc import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["UResQ", "get_conv_wrapper"]


class PassModule(nn.Module):
    def forward(self, x):
        return x


def _own(*tensors) -> bool:
    """True when the glue op runs on the repo's kernels: CUDA fp32, 5-D, nothing to differentiate."""
    if os.environ.get("EFFQ_GLUE", "own") == "lib":
        return False
    recording = torch.is_grad_enabled()
    return all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 5 and not (recording and t.requires_grad)
               for t in tensors)


def _int_triple(v):
    """(a, b, c) if v is an integer or a triple of integers (floats with integral value count), else None."""
    t = tuple(v) if isinstance(v, Iterable) else (v,) * 3
    if len(t) != 3 or any(float(e) != int(e) or int(e) < 1 for e in t):
        return None
    return tuple(int(e) for e in t)


class GlueReLU(nn.ReLU):
    def forward(self, x):
        if _own(x) and x.is_contiguous():
            from . import ops
            return ops.relu(x, self.inplace)
        return super().forward(x)


class GlueMaxPool3d(nn.MaxPool3d):
    """MaxPool3d(k, k); ``fuse_relu`` applies the ReLU of the unit behind it in the same pass."""

    def __init__(self, kernel, stride, fuse_relu=False):
        super().__init__(kernel, stride)
        self.fuse_relu = fuse_relu

    def forward(self, x):
        k = _int_triple(self.kernel_size)
        if _own(x) and k is not None and k == _int_triple(self.stride) and k[2] <= 2 and self.padding == 0 \
                and self.dilation == 1 and not self.ceil_mode and not self.return_indices:
            from . import ops
            return ops.maxpool3d(x, k, relu_after=self.fuse_relu)
        y = super().forward(x)
        return F.relu(y) if self.fuse_relu else y


class GlueUpsample(nn.Upsample):
    """nn.Upsample; ``forward(x, skip)`` adds the skip connection in the same pass."""

    def forward(self, x, skip=None):
        f = _int_triple(self.scale_factor) if self.scale_factor is not None else None
        if f is not None and self.mode == "trilinear" and not self.align_corners and not self.recompute_scale_factor \
                and _own(x) and (skip is None or _own(skip)):
            from . import ops
            return ops.upsample_trilinear(x, f, skip)
        y = super().forward(x)
        return y if skip is None else y + skip


def _add(a, b):
    """Residual add (factory_blk.py:166)."""
    if a.shape == b.shape and _own(a, b):
        from . import ops
        return ops.add(a, b)
    return a + b


def ReLU(inplace=True):
    def make(inp=None):
        return GlueReLU(inplace if inp is None else inp)
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
        y = self.block2(self.block1(x))          # block1's in-place ReLU has rewritten x by now (reference behaviour)
        return _add(y, self.projection(x))


def _down(kernel, Conv, nla, bn, blk_type):
    """factory_blk.py:18-42: MaxPool then a 1x1x1 unit."""
    def make(cin, cout):
        seq = nn.Sequential()
        unit = _Unit(blk_type, cin, cout, 1, 1, 0, 1, Conv, bn, nla, 0)
        # "mid" units start with the ReLU: the pooling pass applies it (relu(max) == max(relu), exactly) and the
        # unit's own ReLU module, which would re-read and re-write the pooled tensor, becomes a pass-through
        fuse = blk_type == "mid" and isinstance(unit.relu, nn.ReLU)
        if fuse:
            unit.relu = PassModule()
        seq.add_module("pool", GlueMaxPool3d(kernel, kernel, fuse_relu=fuse))
        seq.add_module("block", unit)
        return seq
    return make


def _up(scale, Conv, nla, bn, blk_type):
    """factory_blk.py:45-69: 1x1x1 unit (if the width changes) then trilinear upsampling."""
    def make(cin, cout):
        seq = nn.Sequential()
        if cin != cout:
            seq.add_module("block", _Unit(blk_type, cin, cout, 1, 1, 0, 1, Conv, bn, nla, 0))
        seq.add_module("trilinear", GlueUpsample(scale_factor=scale, mode="trilinear"))
        return seq
    return make


class _Fuser(nn.Module):
    """factory_blk.py:72-93: upsample and add the skip connection."""

    def __init__(self, upsampler):
        super().__init__()
        self.upsampler = upsampler

    def forward(self, x, skip):
        for name, m in self.upsampler.named_children():
            x = m(x, skip) if name == "trilinear" else m(x)      # the upsampling pass adds the skip connection
        return x


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
                            head.add_module("extra_up", GlueUpsample(scale_factor=extra, mode="trilinear"))
                    self.classifiers.add_module(f"AuxClassifier{i + 1}", head)

        self.final_cls = nn.Sequential()
        self.final_cls.add_module("cls", ConvLast(width_config[-1], num_classes, 1, 1, 0))
        if init_stride not in (1, (1, 1), (1, 1, 1)):
            self.final_cls.add_module("extra_up", GlueUpsample(scale_factor=init_stride, mode="trilinear"))

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
