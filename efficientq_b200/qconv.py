"""Per-layer quantizer modules: the drop-in boundary of the reference.

``PTQConv`` mirrors reference src/models/PTQConv.py:11-175 (same constructor, the
``alpha_act`` / ``alpha_w`` 0-dim parameters and state-dict keys, the four-state
mode machine, ``store_int_weight`` / ``restore_fp_weight``).  ``EfficientQConv``
mirrors src/models/EfficientQConv.py:12-166; its ``ptq(x)`` runs the layer's ADMM
calibration on the GPU through the C-ABI kernels (``layer_engine``).  Keep
``QConv`` in the class names: the reference duck-types on it
(src/models/model_blk.py:26-34).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .layer_engine import LayerCalibrator

__all__ = ["PTQConv", "EfficientQConv"]


class PTQConv(nn.Conv3d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, q_weight=True, qlvl=8, q_act=True, qlvl_act=8, **kwQ):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        if self.dilation != (1, 1, 1) or self.groups != 1:
            raise NotImplementedError("effq_b200 quantizer layers support dilation 1, groups 1 "
                                      "(all the reference's configs use)")
        self.conv_param = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)
        self.q_act, self.q_weight = q_act, q_weight
        self.qlvl_w, self.qlvl_act = qlvl, qlvl_act
        self.kwQ = kwQ
        self.alpha_act = nn.Parameter(torch.tensor(1.))
        self.alpha_w = nn.Parameter(torch.tensor(1.))
        self.act_in = self.output_fp = self.grad_in = self.grad_out = None
        self.name = None
        self.snap_dir = kwQ.get("snap_dir", None)
        self.w_backup = self.b_backup = None
        self._fp, self._quantizing, self._quantized, self._init_act = True, False, False, False
        self._act_inited = False
        self._wcodes_cache = None
        self._tune_ctx = None            # set by tune.tune_activation_range while alpha_act is being refined

    # -- mode machine (PTQConv.py:44-72) ------------------------------------------------
    def _mode(self, fp=False, quantizing=False, quantized=False, init_act=False):
        self._fp, self._quantizing, self._quantized, self._init_act = fp, quantizing, quantized, init_act

    def set_fp(self):
        self._mode(fp=True)

    def set_quantizing(self):
        self._mode(quantizing=True)

    def set_quantized(self):
        self._mode(quantized=True)

    def set_init_act(self):
        self._mode(init_act=True)

    def qweight_init_iter(self):
        pass

    def qparam_init(self):
        pass

    def perform_quantization(self):
        pass

    def backup_weight(self):
        self.w_backup = self.weight.data.cpu().clone()
        if self.bias is not None:
            self.b_backup = self.bias.data.cpu().clone()

    # -- fake-quant (PTQConv.py:110-116) through the fused CUDA kernel --------------------
    def _quantize_w(self):
        y, _ = ops.fakequant(self.weight.data, self.alpha_w.data, self.qlvl_w, -1.0, 1.0)
        return y

    def _quantize_act(self, x):
        y, _ = ops.fakequant(x, self.alpha_act.data, self.qlvl_act, 0.0, 1.0)
        return y

    def init_alpha_act(self, x):
        """PTQConv.py:74-78."""
        st = ops.ScaleState(x.device)
        ops.scale_search(x.contiguous(), self.qlvl_act, 0.0, 1.0, st)
        self.alpha_act.data = st.a_f32()
        self._act_inited = True
        return ops.fakequant_state(x, st, self.qlvl_act, 0.0, 1.0)

    def ptq(self, x):
        raise NotImplementedError

    # -- integer export (PTQConv.py:125-152) -------------------------------------------------
    def _aw(self):
        """alpha_w broadcast over the weight: a scalar (reference) or one value per output channel (extension)."""
        a = self.alpha_w.data
        return a.view(-1, 1, 1, 1, 1) if a.dim() == 1 else a

    def store_int_weight(self):
        a = self._aw().to(self.weight.device)
        b = self.weight.data / a
        delta = 2 / (self.qlvl_w - 1)
        w_int = torch.round((b + 1) / delta)
        w_int = w_int.to(torch.uint8) if self.qlvl_w <= 256 else w_int.to(torch.int32)
        self.weight.requires_grad = False
        self.weight.data = w_int.data.cpu()

    def restore_fp_weight(self):
        self._wcodes_by_dtype = self._wcodes_cache = self._dgrad_cache = None
        delta = 2 / (self.qlvl_w - 1)
        self.weight.data = self._aw().to(self.weight.device) * (self.weight.data.float() * delta - 1)

    # -- forward dispatch (PTQConv.py:154-174) ------------------------------------------------
    def _conv(self, x):
        out, _ = ops.conv3d_f32(x, self.weight.data, self.bias.data if self.bias is not None else None,
                                self.stride, self.padding)
        return out

    def _weight_codes(self, code_dtype=ops.CODE_BF16):
        """(codes, scale) of the stored fake-quant weights for the tensor-core conv: odd integers
        2c-(L-1) in the tensor-core layout and the grid scale a with  weight == a * codes / (L-1)  exactly,
        or None when the weights do not lie on a symmetric L-level grid.  The scale is recovered from the
        weights themselves, NOT from alpha_w: after calibration alpha_w is the LAST iterate's scale while
        the weights come from the BEST iterate (reference quirk, EfficientQConv.py:155-158), and the
        reference's quantized forward uses the stored weights as they are (PTQConv.py:163-167).
        Cached until the weights change (one host read per weight version)."""
        key = (self.weight.data_ptr(), self.weight._version, code_dtype)
        if getattr(self, "_wcodes_by_dtype", None) is None:
            self._wcodes_by_dtype = {}                    # one entry per operand type (forward: e4m3, dgrad: bf16)
        self._wcodes_cache = self._wcodes_by_dtype.get(code_dtype)
        if self._wcodes_cache is None or self._wcodes_cache[0] != key:
            w = self.weight.data.float()
            lm1 = float(self.qlvl_w - 1)
            found = None
            if getattr(self, "channel_wise", False):
                # one grid per output channel: every row is tested against the candidates, rows may differ in theirs
                c2 = w.shape[0]
                wr = w.reshape(c2, -1)
                wmax = wr.abs().max(1, keepdim=True).values
                a_vec = torch.zeros_like(wmax)
                codes = torch.zeros_like(wr)
                done = wmax.squeeze(1) == 0                     # an all-zero row has no scale: codes stay 0 -> reject below
                ok_rows = torch.zeros(c2, dtype=torch.bool, device=w.device)
                for k in range(self.qlvl_w - 1, max(self.qlvl_w - 9, 0), -2):
                    a = wmax * (lm1 / k)
                    cd = torch.round(wr / a.clamp_min(1e-30) * lm1)
                    good = ((torch.remainder(cd + lm1, 2.0) == 0).all(1) &
                            ((cd * (a / lm1) - wr).abs().max(1).values <= 1e-6 * wmax.squeeze(1))) & ~ok_rows & ~done
                    a_vec[good], codes[good] = a[good], cd[good]
                    ok_rows |= good
                if bool(ok_rows.all().item()):
                    found = (ops.pack_weight_codes(codes.view_as(w), code_dtype), a_vec.reshape(-1).clone(), codes.view_as(w))
                self._wcodes_cache = self._wcodes_by_dtype[code_dtype] = (key, found)
                return self._wcodes_cache[1]
            wmax = w.abs().max()
            for k in range(self.qlvl_w - 1, max(self.qlvl_w - 9, 0), -2):     # largest |code| present: L-1, L-3, ...
                a = wmax * (lm1 / k)
                codes = torch.round(w / a * lm1)
                odd = torch.remainder(codes + lm1, 2.0) == 0
                exact = (codes * (a / lm1) - w).abs().max() <= 1e-6 * wmax
                if bool((odd.all() & exact & (wmax > 0)).item()):
                    found = (ops.pack_weight_codes(codes, code_dtype), a.reshape(1).clone(), codes)
                    break
            self._wcodes_cache = self._wcodes_by_dtype[code_dtype] = (key, found)
        return self._wcodes_cache[1]

    def _dgrad_operands(self):
        """(packed codes, scale) of the TRANSPOSED, spatially flipped weights for the tensor-core dgrad of a
        stride-1 layer (d qact = conv(d out, W^T flipped)), or None when the weights are on no exact grid.
        Cached with the weights."""
        wc = self._weight_codes(ops.CODE_BF16)
        if wc is None:
            return None
        key = self._wcodes_cache[0]
        if getattr(self, "_dgrad_cache", None) is None or self._dgrad_cache[0] != key:
            codes_t = wc[2].flip(2, 3, 4).transpose(0, 1).contiguous()              # [C1][C2][kd][kh][kw]
            scale = (wc[1].double() / (self.qlvl_w - 1)).float().reshape(1)
            self._dgrad_cache = (key, (ops.pack_weight_codes(codes_t, ops.CODE_BF16), scale))
        return self._dgrad_cache[1]

    def _quantized_forward(self, x):
        """Deployment forward (PTQConv.py:163-167): fake-quant activations + conv with the stored
        quantized weights.  3x3x3 / 1x1x1 stride-1 layers run on the tcgen05 integer-code conv
        (codes from the same fp32 arithmetic as `_quantize_act`, so the result equals the
        reference's conv3d(qact, qweight) up to fp32 summation order); other layers use the
        generic fp32 kernels."""
        if self.q_act and self.q_weight and self.qlvl_act <= 256 and self.qlvl_w <= 256 and \
                ops.conv3d_tc_supported(x.shape, self.out_channels, self.kernel_size, self.stride, self.padding):
            fp8 = ops.fp8_codes_enabled() and self.qlvl_act <= 16 and self.qlvl_w <= 16 and \
                ops.conv3d_tc_supported(x.shape, self.out_channels, self.kernel_size, self.stride, self.padding,
                                        ops.CODE_E4M3)
            wc = self._weight_codes(ops.CODE_E4M3 if fp8 else ops.CODE_BF16)
            if wc is not None:                  # weights on an exact L-level grid (always, after calibration)
                wcodes, w_scale, _ = wc
                if fp8:
                    _, xcodes = ops.quantize_act_ndhwc(x, self.qlvl_act, alpha=self.alpha_act.data, bf16=False, e4m3=True)
                else:
                    xcodes = ops.quantize_act_ndhwc(x, self.qlvl_act, alpha=self.alpha_act.data)
                scale = (self.alpha_act.data.double() / (self.qlvl_act - 1) *
                         w_scale.double() / (self.qlvl_w - 1)).float().reshape(-1)
                out, _ = ops.conv3d_tc(xcodes, wcodes, self.bias.data if self.bias is not None else None,
                                       scale[:1], self.out_channels, self.kernel_size, want_out=True,
                                       scale_vec=scale if scale.numel() > 1 else None)
                return out
        qact = self._quantize_act(x) if self.q_act else x
        return self._conv(qact)

    def _fp_forward(self, x):
        """FP pass (PTQConv.py:157-158, F.conv3d): on the GPU it runs on the repo's own kernels -- the tcgen05 conv on
        fixed-point digit planes (exact integer accumulation, ops.conv3d_fp) where the geometry allows (3x3x3 / 1x1x1, stride 1, channels in multiples of 16),
        the generic fp32 kernel otherwise (conv0: 4 input channels, stride 2; final_cls: 3 output channels).
        EFFQ_FP_CONV=lib restores the library conv (bring-up comparison).  Autograd (training, alpha refinement) and
        CPU tensors keep F.conv3d."""
        if not x.is_cuda or torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad) or \
                os.environ.get("EFFQ_FP_CONV", "own") == "lib" or self.dilation != (1, 1, 1) or self.groups != 1:
            return F.conv3d(x, self.weight, self.bias, self.stride, self.padding)
        if ops.conv3d_fp_supported(x.shape, self.out_channels, self.kernel_size, self.stride, self.padding):
            return ops.conv3d_fp(x, self.weight.data, self.bias.data if self.bias is not None else None, self.kernel_size)
        return self._conv(x.contiguous())

    def forward(self, x):
        if self._fp:
            return self._fp_forward(x)
        if self._quantizing:
            return self.ptq(x)                      # returns conv3d(qact, weight*, bias*) of the calibrated layer
        if self._quantized:
            if self._tune_ctx is not None:          # differentiable (STE) forward, ptqer.py:238-272
                return self._tune_ctx["forward"](self, x)
            return self._quantized_forward(x.contiguous())
        if self._init_act:
            return self._conv(self.init_alpha_act(x))
        raise RuntimeError(f"Unknown FP/Quant setting: FP={self._fp}, "
                           f"Quantizing={self._quantizing}, Quantized={self._quantized}")


class EfficientQConv(PTQConv):
    """EfficientQ layer: ADMM with the closed-form proximal step, on the B200."""
    _engines = {}

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, q_weight=True, qlvl=8, q_act=True, qlvl_act=8, **kwQ):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         q_weight, qlvl, q_act, qlvl_act, **kwQ)
        self.lwq_iter, self.lwq_rho, self.lwq_rho_max, self.lwq_eta = 200, 10, 1000, 1   # EfficientQConv.py:23-26
        self.lwq_fold_bn = True
        self.lwq_verbose = kwQ.get("lwq_verbose", False)
        # optional extension (north star: "weights per output channel"): one alpha_w per output channel.  The
        # reference's live path is per-tensor (PTQConv.py:26-27) and that stays the default
        self.channel_wise = bool(kwQ.get("lwq_channel_wise", False))
        self.mask_pyramid = None
        self.layer_loss = None
        self.report = None
        self.dist = None                 # set by the orchestrator for sharded calibration
        self.keep_history = False

    def _engine(self, device) -> LayerCalibrator:
        key = (device.index, self.lwq_iter, id(self.dist))
        eng = EfficientQConv._engines.get(key)
        if eng is None:
            eng = LayerCalibrator(device, dist=self.dist, n_iter=self.lwq_iter, rho0=self.lwq_rho,
                                  rho_max=self.lwq_rho_max, eta0=self.lwq_eta)
            EfficientQConv._engines = {key: eng}      # one live engine (its scratch is large)
        eng.keep_history = self.keep_history
        return eng

    def ptq(self, x):
        """EfficientQConv.py:33-166.  Afterwards weight/bias hold the fake-quant values and
        alpha_act/alpha_w the scales; returns the calibrated layer's output."""
        if not x.is_cuda:
            raise ops.EffqError("EfficientQConv.ptq needs CUDA tensors: there is no CPU path")
        if self.output_fp is None:
            raise RuntimeError("output_fp missing: run the FP pass with forward hooks first")
        print(f"Calibrating {self.name}")                 # unconditional in the reference (EfficientQConv.py:52)
        out_fp = self.output_fp.to(x.device)
        w, b, a_w, a_act, out_q, rep = self._engine(x.device).run(
            x, self.weight.data, self.bias.data if self.bias is not None else None, out_fp,
            self.stride, self.padding, self.qlvl_w, self.qlvl_act, self.q_act, self.mask_pyramid,
            name=self.name or "", channel_wise=self.channel_wise)
        self.weight.data = w.clone()
        self._wcodes_by_dtype = self._wcodes_cache = self._dgrad_cache = None    # keyed by address + version: drop
        if self.bias is not None:
            self.bias.data = b.clone()
        self.alpha_w.data = a_w.to(x.dtype)
        if a_act is not None:
            self.alpha_act.data = a_act.to(x.dtype)
        self.report = rep
        if self.layer_loss is not None:
            self.layer_loss.append(f"{self.name:45s}:{rep.final_loss}")
        self.output_fp = None                      # release the target
        return out_q
