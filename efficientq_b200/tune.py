"""End-to-end refinement of the activation ranges (reference src/ptqer.py:238-272).

``tune_activation_range(model, output_fp, data_batch, max_iter)`` mirrors the reference function of
the same name: every quantizer module is put in its quantized mode, Adam (lr 5e-4) runs on all
``alpha_act`` parameters, the loss is the MSE between the quantised network's output and the FP
output.  The reference defines it but never calls it from ``do_ptq``; here it is reachable through
``ptqer.do_ptq`` when the YAML/CLI sets ``tune_act_iter > 0`` (an extension key, default 0).

What runs where:
  * forward of a quantizer layer: the deployment forward (fake-quant codes + tcgen05 conv,
    ``PTQConv._quantized_forward``);
  * backward of a quantizer layer: conv dgrad -- on the tcgen05 conv kernel for stride-1 layers with >= 16
    channels (``conv_dgrad``: integer weight codes x three exact bf16 planes of the gradient), the library's
    ``torch.nn.grad.conv3d_input`` otherwise -- then ONE pass of ``effq_fakequant_ste_bwd`` producing grad_x
    and d loss / d alpha_act (csrc/tune.cu);
  * glue ops between the layers (ReLU, pooling, upsampling, residual adds): stock autograd;
  * the optimiser: ``effq_adam_step`` on the flat vector that all ``alpha_act`` parameters alias;
  * sharded calibration: each rank differentiates its own volumes, the alpha gradients (<= 28
    doubles) are all-reduced, every rank applies the same update.
Weights are frozen (the reference computes their gradients and throws them away).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import ops
from .dist import DistCtx
from .qconv import PTQConv

__all__ = ["tune_activation_range"]


def conv_dgrad(mod: PTQConv, x_shape, grad_out: torch.Tensor) -> torch.Tensor:
    """d loss / d qact of a quantizer layer.  Stride-1 3x3x3 / 1x1x1 layers whose weights lie on their L-level grid
    run on the tcgen05 conv kernel: dgrad = conv(grad_out, flipped W^T) with the integer weight codes as one
    operand and grad_out as THREE bf16 planes (hi + mid + lo = the fp32 value exactly, so every product is exact
    and the result has fp32-conv accuracy), one launch per plane.  Other shapes: the library's fp32 dgrad."""
    n, c1, d, h, w = x_shape
    c2 = mod.out_channels
    ks = mod.kernel_size
    ok = mod.q_weight and mod.qlvl_w <= 256 and tuple(mod.stride) == (1, 1, 1) and \
        tuple(mod.padding) == tuple((k - 1) // 2 for k in ks) and \
        ops.conv3d_tc_supported((n, c2, d, h, w), c1, ks, 1, mod.padding) and os.environ.get("EFFQ_DGRAD_TC", "1") != "0"
    opnd = mod._dgrad_operands() if ok else None
    if opnd is None:
        return torch.nn.grad.conv3d_input(x_shape, mod.weight.data, grad_out.contiguous(), mod.stride, mod.padding)
    wcodes_t, scale = opnd
    planes = ops.split3_ndhwc(grad_out)                              # fused layout change + exact 3-plane split
    if planes is None:                                               # odd shapes: the same with stock elementwise ops
        g = grad_out.permute(0, 2, 3, 4, 1).contiguous()
        hi = g.to(torch.bfloat16)
        r1 = g - hi.float()
        mid = r1.to(torch.bfloat16)
        planes = [hi, mid, (r1 - mid.float()).to(torch.bfloat16)]
    hi, mid, lo = planes
    out, _ = ops.conv3d_tc(hi, wcodes_t, None, scale, c1, ks, want_out=True)
    for plane in (mid, lo):
        part, _ = ops.conv3d_tc(plane, wcodes_t, None, scale, c1, ks, want_out=True)
        out += part
    return out


class _QuantLayerSTE(torch.autograd.Function):
    """y = conv3d(discretize(x/alpha)*alpha, Wq, b) with the reference's straight-through gradient.
    ``anchor`` is a dummy that requires grad, so the node is recorded (and d loss / d alpha computed)
    even when the layer input itself needs no gradient (first quantised layer)."""

    @staticmethod
    def forward(ctx, x, anchor, mod: PTQConv, alpha_view: torch.Tensor, grad_slot: torch.Tensor):
        x = x.contiguous()
        ctx.mod, ctx.alpha_view, ctx.grad_slot = mod, alpha_view, grad_slot
        ctx.save_for_backward(x)
        return mod._quantized_forward(x)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        mod = ctx.mod
        g_qact = conv_dgrad(mod, x.shape, grad_out)
        gx = ops.fakequant_ste_bwd(x, g_qact, ctx.alpha_view, mod.qlvl_act, 0.0, 1.0, ctx.grad_slot,
                                   want_grad_x=ctx.needs_input_grad[0])
        return gx, None, None, None, None


def _tuning_forward(mod: PTQConv, x):
    t = mod._tune_ctx
    if not mod.q_act:
        # no activation range on this layer (first / last layer): plain differentiable conv, frozen weights
        return F.conv3d(x, mod.weight.data, mod.bias.data if mod.bias is not None else None, mod.stride, mod.padding)
    return _QuantLayerSTE.apply(x, t["anchor"], mod, t["alpha"], t["grad"])


def tune_activation_range(model, output_fp: torch.Tensor, data_batch: torch.Tensor, max_iter: int = 1000,
                          need_init: bool = False, dist: Optional[DistCtx] = None, lr: float = 5e-4) -> List[float]:
    """Reference ptqer.py:238-272.  Returns the per-iteration losses; alpha_act is updated in place."""
    dist = dist or DistCtx()
    if not data_batch.is_cuda:
        raise ops.EffqError("tune_activation_range needs CUDA tensors: there is no CPU path")
    dev = data_batch.device
    mods = [m for m in model.modules() if isinstance(m, PTQConv)]
    if need_init:                                   # ptqer.py:250-252
        for m in mods:
            m.set_init_act()
        with torch.no_grad():
            model(data_batch)
    for m in mods:
        m.set_quantized()
    tuned = [m for m in mods if m.q_act]
    n = len(tuned)
    flat = torch.stack([m.alpha_act.data.detach().reshape(()).float() for m in tuned]).contiguous()   # master copy
    grads = torch.zeros(n, dtype=torch.float64, device=dev)
    exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
    exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
    anchor = torch.zeros((), device=dev, requires_grad=True)
    for i, m in enumerate(tuned):
        m.alpha_act.data = flat[i]                  # 0-dim view: the module reads the optimiser's vector
        m._tune_ctx = dict(forward=_tuning_forward, alpha=flat[i:i + 1], grad=grads[i:i + 1], anchor=anchor)
    for m in mods:
        if not m.q_act:
            m._tune_ctx = dict(forward=_tuning_forward)
    frozen = [(p, p.requires_grad) for p in model.parameters()]
    for p, _ in frozen:
        p.requires_grad_(False)
    x_in = data_batch.detach()
    losses = []
    world_numel = float(output_fp.numel() * dist.world)
    try:
        for it in range(max_iter):
            grads.zero_()
            with torch.enable_grad():
                out_q = model(x_in)
                # global MSE over all ranks' volumes: local SSE / global numel
                loss = ((out_q - output_fp) ** 2).sum() / world_numel
            loss.backward()
            anchor.grad = None
            if dist.world > 1:
                dist.all_reduce_sum(grads)          # the quantizer-parameter gradient all-reduce
                lt = dist.all_reduce_sum(loss.detach().double().reshape(1))
                losses.append(lt)
            else:
                losses.append(loss.detach())
            ops.adam_step(flat, grads, exp_avg, exp_avg_sq, it + 1, lr)
    finally:
        for m in mods:
            m._tune_ctx = None
        for i, m in enumerate(tuned):
            m.alpha_act.data = flat[i].clone()
        for p, rg in frozen:
            p.requires_grad_(rg)
    return [float(v) for v in torch.stack([t.reshape(()).double() for t in losses]).cpu().tolist()] if losses else []
