"""Tensor-level wrappers over the C-ABI (one Python function per entry point).

Each wrapper checks dtype/contiguity/device, passes raw device pointers and the
current CUDA stream to ``libeffq_b200.so`` and raises ``EffqError`` on failure.
Nothing here computes on the host or through PyTorch ops: if the library is
missing the first call raises (no fallback by design).
"""
from __future__ import annotations

import ctypes as C
import weakref
import os
import struct
from typing import Optional, Tuple

import torch

from . import capi
from .capi import Geom, EffqError, ptr, stream, check


class KernelTimer:
    """Optional CUDA-event timing of the major kernels on the launching stream (bench.py).
    ``work`` is the ALGORITHMIC work of the launch: {"flops": ..} or {"bytes": ..}."""

    def __init__(self):
        self.enabled = False
        self.only = None             # optional tuple of name prefixes: time just those kernels
        self.records = {}
        self.streams = {}            # name -> set of stream handles the kernel was launched on

    def run(self, name, work, fn):
        rec = capi._recorder
        if rec is not None:                  # recording a launch sequence: remember which kernel the calls belong to
            rec.tag = (name, work)
            try:
                return self._run(name, work, fn)
            finally:
                rec.tag = None
        return self._run(name, work, fn)

    def wants(self, name) -> bool:
        return self.enabled and (self.only is None or name.startswith(self.only))

    def _run(self, name, work, fn):
        if not self.enabled or (self.only is not None and not name.startswith(self.only)):
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.records.setdefault(name, []).append((e0, e1, work))
        self.streams.setdefault(name, set()).add(torch.cuda.current_stream().cuda_stream)
        return out

    def summary(self):
        """name -> dict(launches, ms, flops, bytes); call after a device synchronize."""
        res = {}
        for name, recs in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b, _ in recs)
            res[name] = dict(launches=len(recs), ms=ms, flops=sum(w.get("flops", 0) for _, _, w in recs),
                             flops_alg=sum(w.get("flops_alg", w.get("flops", 0)) for _, _, w in recs),
                             bytes=sum(w.get("bytes", 0) for _, _, w in recs),
                             pass_bytes=sum(w.get("pass_bytes", 0) for _, _, w in recs))
        return res

    def reset(self):
        self.records = {}
        self.streams = {}


timer = KernelTimer()


def replay(calls) -> None:
    """Re-issue a recorded launch sequence (capi.record) on the current stream: same functions, same
    arguments.  Kernels the timer wants are still bracketed by CUDA events."""
    for fn, args, tag, name in calls:
        if tag is not None and timer.wants(tag[0]):
            timer._run(tag[0], tag[1], lambda: check(fn(*args), name))
        else:
            rc = fn(*args)
            if rc != 0:
                check(rc, name)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise EffqError(f"{name}: expected a CUDA float32 tensor, got {t.dtype} on {t.device}")
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------------------
# device-resident state structs
# ---------------------------------------------------------------------------
class ScaleState:
    """Device copy of ``effq_scale_state``; ``read()`` is the only host sync."""
    FMT = "<ddddiiii"

    def __init__(self, device):
        self.buf = torch.zeros(capi.SCALE_STATE_BYTES, dtype=torch.uint8, device=device)

    @property
    def p(self):
        return ptr(self.buf)

    def read(self) -> dict:
        a, a_prev, s_bv, s_bb, passes, conv, failed, _ = struct.unpack(self.FMT, self.buf.cpu().numpy().tobytes())
        return dict(a=a, a_prev=a_prev, s_bv=s_bv, s_bb=s_bb, passes=passes, converged=conv, failed=failed)

    def a_f32(self) -> torch.Tensor:
        """fp32(a) as a 0-dim device tensor, without a host sync."""
        return self.buf[:8].view(torch.float64)[0].to(torch.float32)

    def set_a(self, a: float) -> None:
        self.buf[:8].view(torch.float64).fill_(a)


class ScaleStateRows:
    """Array of ``effq_scale_state``, one per output channel (per-output-channel weight scales)."""

    def __init__(self, device, rows: int):
        self.rows = rows
        self.buf = torch.zeros(rows * capi.SCALE_STATE_BYTES, dtype=torch.uint8, device=device)

    @property
    def p(self):
        return ptr(self.buf)

    def read(self):
        """[rows] dicts; a host sync."""
        raw = self.buf.cpu().numpy().tobytes()
        n = capi.SCALE_STATE_BYTES
        out = []
        for r in range(self.rows):
            a, a_prev, s_bv, s_bb, passes, conv, failed, _ = struct.unpack(ScaleState.FMT, raw[r * n:(r + 1) * n])
            out.append(dict(a=a, passes=passes, converged=conv, failed=failed))
        return out


class AdmmState:
    """Device copy of ``effq_admm_state``."""
    FMT = "<dffiiffif"

    def __init__(self, device):
        self.buf = torch.zeros(capi.ADMM_STATE_BYTES, dtype=torch.uint8, device=device)

    @property
    def p(self):
        return ptr(self.buf)

    def reset(self):
        self.buf.zero_()

    def read(self) -> dict:
        sse, best, last, best_it, it, cs, aw, take, bcs = struct.unpack(self.FMT, self.buf.cpu().numpy().tobytes())
        return dict(sse=sse, best_loss=best, last_loss=last, best_iter=best_it, iter=it, conv_scale=cs, a_w=aw,
                    best_conv_scale=bcs)

    def conv_scale_ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 24)

    def best_conv_scale_ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 36)

    def a_w_tensor(self) -> torch.Tensor:
        return self.buf[28:32].view(torch.float32)[0]


def workspace(nbytes: int, device) -> torch.Tensor:
    """Zero-initialised scratch (the kernels' completion counters start at 0)."""
    return torch.zeros(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------
# (a) fake-quant
# ---------------------------------------------------------------------------
def fakequant(x: torch.Tensor, alpha: torch.Tensor, nlvl: int, lo: float, hi: float,
              want_values: bool = True, want_codes: bool = False):
    """discretize(x/alpha, nlvl, lo, hi)*alpha (+ uint8 codes) -- PTQConv.py:110-116."""
    x = _f32c(x, "x")
    alpha = _f32c(alpha.reshape(1), "alpha")
    y = torch.empty_like(x) if want_values else None
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_codes else None
    nbytes = x.numel() * (4 + (4 if want_values else 0) + (1 if want_codes else 0))
    timer.run("fakequant_f32", {"bytes": nbytes}, lambda: check(
        capi.load().effq_fakequant_f32(ptr(x), x.numel(), ptr(alpha), float(lo), float(hi), int(nlvl),
                                       ptr(y), ptr(codes), stream()), "effq_fakequant_f32"))
    return y, codes


def fakequant_state(x: torch.Tensor, state: ScaleState, nlvl: int, lo: float, hi: float) -> torch.Tensor:
    """fp32(a)*fp32(level) with the scale-search state (EfficientQConv.py:68-70)."""
    x = _f32c(x, "x")
    y = torch.empty_like(x)
    timer.run("fakequant_state", {"bytes": 8 * x.numel()}, lambda: check(
        capi.load().effq_fakequant_state(ptr(x), x.numel(), state.p, float(lo), float(hi), int(nlvl), ptr(y),
                                         stream()), "effq_fakequant_state"))
    return y


CODE_BF16, CODE_E4M3 = 0, 1
E4M3 = torch.float8_e4m3fn


def code_dtype_of(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return CODE_BF16
    if t.dtype == E4M3:
        return CODE_E4M3
    raise EffqError(f"integer codes must be bf16 or float8_e4m3fn, got {t.dtype}")


def fp8_codes_enabled() -> bool:
    """EFFQ_FP8=0 forces bf16 codes everywhere (results are bit-identical; this only changes speed)."""
    return os.environ.get("EFFQ_FP8", "1") != "0"


def quantize_act_ndhwc(x: torch.Tensor, nlvl: int, state: Optional[ScaleState] = None,
                       alpha: Optional[torch.Tensor] = None, bf16: bool = True, e4m3: bool = False):
    """NCDHW fp32 -> (N,D,H,W,C) integer codes, the tensor-core operand: a bf16 tensor (default), or
    with e4m3=True the pair (bf16 or None, e4m3) written by the same pass over x."""
    x = _f32c(x, "x")
    n, c, d, h, w = x.shape
    out = torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=x.device) if bf16 else None
    out8 = torch.empty((n, d, h, w, c), dtype=E4M3, device=x.device) if e4m3 else None
    use64 = 1 if state is not None else 0
    a = _f32c(alpha.reshape(1), "alpha") if alpha is not None else None
    nbytes = x.numel() * (4 + 2 * bool(bf16) + bool(e4m3))
    timer.run("quantize_act_ndhwc", {"bytes": nbytes}, lambda: check(
        capi.load().effq_quantize_act_ndhwc(ptr(x), n, c, d * h * w, int(nlvl), state.p if state else None,
                                            ptr(a), use64, ptr(out), ptr(out8), stream()),
        "effq_quantize_act_ndhwc"))
    return (out, out8) if e4m3 else out


# ---------------------------------------------------------------------------
# (a3) scale search
# ---------------------------------------------------------------------------
_ss_ws = {}


def _scale_ws(device, numel: int = 0) -> torch.Tensor:
    """Scale-search workspace of the device, grown on demand (large tensors need room for the
    ambiguous-element list of the interval-stable passes)."""
    key = (device.type, device.index)
    need = capi.load().effq_scale_search_workspace(int(numel))
    if key not in _ss_ws or _ss_ws[key].numel() < need:
        _ss_ws[key] = workspace(need, device)
    return _ss_ws[key]


def scale_search_diag(device) -> dict:
    """Pass-type counts of the last streamed search on this device (SSWorkspace::diag, csrc/scale_search.cu)."""
    v = _scale_ws(device)[32:64].view(torch.float64).cpu().tolist()
    return {"classifying_passes": int(v[0]), "list_passes": int(v[1]), "reclassifying_list_passes": int(v[2]),
            "l1_passes": int(v[3])}


def _pair_view(v1: torch.Tensor, v2: Optional[torch.Tensor]):
    """(rows, cols, ld1, ld2) for v = v1 (+ v2).  v1 may be a row-strided 2-D view (w* with its
    bias column); v2, when given, is contiguous with the same number of elements."""
    for t in (v1, v2):
        if t is not None and (t.dtype != torch.float32 or not t.is_cuda):
            raise EffqError("scale search: expected CUDA float32 tensors")
    if v2 is not None and (v2.numel() != v1.numel() or not v2.is_contiguous()):
        raise EffqError("scale search: v2 must be contiguous and match v1 in size")
    if v1.is_contiguous():
        return 1, v1.numel(), v1.numel(), v1.numel()
    if v1.dim() == 2 and v1.stride(1) == 1:
        return v1.shape[0], v1.shape[1], v1.stride(0), v1.shape[1]
    raise EffqError("scale search needs a contiguous tensor or a row-strided 2-D view")


def scale_search(v1: torch.Tensor, nlvl: int, lo: float, hi: float, state: ScaleState,
                 v2: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None, comm=None) -> ScaleState:
    """project_by_iter on the device (one cooperative launch, no host sync).  With ``comm``
    (dist.PeerLink.comm_ptr) v1 is this rank's shard and every pass is all-reduced in-kernel."""
    rows, cols, ld1, ld2 = _pair_view(v1, v2)
    name = f"scale_search_w_{rows * cols}" if v2 is not None else "scale_search_act"
    if ws is None:
        ws = _scale_ws(v1.device, rows * cols)
    timer.run(name, {"pass_bytes": 4 * rows * cols * (2 if v2 is not None else 1)}, lambda: check(
        capi.load().effq_scale_search(ptr(v1), ld1, ptr(v2), ld2, rows, cols, int(nlvl), float(lo), float(hi),
                                      state.p, ptr(ws), ws.numel(), comm, stream()), "effq_scale_search"))
    return state


def scale_search_rows(v1: torch.Tensor, nlvl: int, lo: float, hi: float, states: ScaleStateRows,
                      v2: Optional[torch.Tensor] = None) -> ScaleStateRows:
    """One independent project_by_iter per ROW of the 2-D view v1 (+ v2): per-output-channel weight scales."""
    if v1.dim() != 2 or v1.stride(1) != 1 or v1.dtype != torch.float32 or not v1.is_cuda:
        raise EffqError("scale_search_rows: expected a 2-D CUDA float32 view with unit column stride")
    rows, cols = v1.shape
    if v2 is not None and (tuple(v2.shape) != (rows, cols) or not v2.is_contiguous()):
        raise EffqError("scale_search_rows: v2 must be contiguous with v1's shape")
    if states.rows != rows:
        raise EffqError("scale_search_rows: one state per row")
    timer.run(f"scale_search_rows_{rows}x{cols}", {"pass_bytes": 8 * rows * cols}, lambda: check(
        capi.load().effq_scale_search_rows(ptr(v1), v1.stride(0), ptr(v2), cols, rows, cols, int(nlvl), float(lo), float(hi),
                                           states.p, stream()), "effq_scale_search_rows"))
    return states


def scale_partial(v1: torch.Tensor, nlvl: int, lo: float, hi: float, state: ScaleState, mode: int,
                  sums: torch.Tensor, ws: torch.Tensor, v2: Optional[torch.Tensor] = None) -> None:
    rows, cols, ld1, ld2 = _pair_view(v1, v2)
    check(capi.load().effq_scale_partial(ptr(v1), ld1, ptr(v2), ld2, rows, cols, int(nlvl), float(lo), float(hi),
                                         state.p, int(mode), ptr(sums), ptr(ws), stream()), "effq_scale_partial")


def scale_step(state: ScaleState, sums: torch.Tensor, mode: int, nlvl: int) -> None:
    check(capi.load().effq_scale_step(state.p, ptr(sums), int(mode), int(nlvl), stream()), "effq_scale_step")


# ---------------------------------------------------------------------------
# (a10) conv forward + squared error
# ---------------------------------------------------------------------------
def conv3d_f32(x, w, bias, stride, padding, want_out=True, target=None, att=None, ws=None, sse=None):
    """Generic fp32 conv (+ fused sum of att*(out-target)^2).  Returns (out, sse[1] f64)."""
    x = _f32c(x, "x")
    w = _f32c(w, "w")
    g = Geom.make(x.shape, w.shape[0], w.shape[2:], stride, padding)
    od, oh, ow = g.out_spatial()
    out = torch.empty((g.n, g.c2, od, oh, ow), dtype=torch.float32, device=x.device) if want_out else None
    lib = capi.load()
    if target is not None:
        target = _f32c(target, "target")
        if ws is None:
            ws = workspace(lib.effq_conv3d_f32_workspace(C.byref(g)), x.device)
        if sse is None:
            sse = torch.zeros(1, dtype=torch.float64, device=x.device)
    if att is not None:
        att = _f32c(att, "att")
    b = _f32c(bias, "bias") if bias is not None else None
    flops = 2.0 * g.n * od * oh * ow * g.c2 * g.c1 * g.taps
    timer.run("conv3d_f32", {"flops": flops}, lambda: check(
        lib.effq_conv3d_f32(ptr(x), ptr(w), ptr(b), C.byref(g), ptr(out), ptr(target), ptr(att), ptr(sse),
                            ptr(ws), stream()), "effq_conv3d_f32"))
    return out, sse


def conv3d_tc_supported(x_shape, c2, ksize, stride, padding, code_dtype: int = CODE_BF16) -> bool:
    g = Geom.make(x_shape, c2, ksize, stride, padding)
    return bool(capi.load().effq_conv3d_tc_supported(C.byref(g), int(code_dtype)))


def conv3d_tc(xcodes: torch.Tensor, wcodes: torch.Tensor, bias, conv_scale_ptr, c2: int, ksize, want_out=True,
              target=None, att=None, ws=None, sse=None, scale_vec=None):
    """tcgen05 conv on codes.  xcodes (N,D,H,W,C1) bf16 or e4m3; wcodes of the same type in the
    tensor-core weight layout (pack_weight_codes / admm_project).  ``scale_vec`` ([C2] fp32): one scale per output
    channel instead of the scalar ``conv_scale_ptr``."""
    cdt = code_dtype_of(xcodes)
    if wcodes.dtype != xcodes.dtype:
        raise EffqError("conv3d_tc: activation and weight codes must have the same element type")
    if not xcodes.is_contiguous() or not xcodes.is_cuda:
        raise EffqError("conv3d_tc: xcodes must be contiguous CUDA NDHWC")
    n, d, h, w, c1 = xcodes.shape
    k = capi._triple(ksize)
    pad = tuple((t - 1) // 2 for t in k)
    g = Geom.make((n, c1, d, h, w), c2, k, 1, pad)
    lib = capi.load()
    out = torch.empty((n, c2, d, h, w), dtype=torch.float32, device=xcodes.device) if want_out else None
    if ws is None:
        ws = workspace(lib.effq_conv3d_tc_workspace(C.byref(g)), xcodes.device)
    if target is not None:
        target = _f32c(target, "target")
        if sse is None:
            sse = torch.zeros(1, dtype=torch.float64, device=xcodes.device)
    if att is not None:
        att = _f32c(att, "att")
    b = _f32c(bias, "bias") if bias is not None else None
    cs = conv_scale_ptr if not isinstance(conv_scale_ptr, torch.Tensor) else ptr(conv_scale_ptr)
    flops = 2.0 * n * d * h * w * c2 * c1 * g.taps
    nbytes = (xcodes.numel() + wcodes.numel()) * xcodes.element_size() + n * d * h * w * c2 * 4 * ((target is not None) + (out is not None))
    name = f"conv3d_tc_c{c1}x{c2}k{k[0]}" + ("_e4m3" if cdt == CODE_E4M3 else "")
    if scale_vec is not None:
        sv = _f32c(scale_vec, "scale_vec")
        if sv.numel() != c2:
            raise EffqError("conv3d_tc: scale_vec needs one entry per output channel")
        timer.run(name, {"flops": flops, "bytes": nbytes}, lambda: check(
            lib.effq_conv3d_tc_pc(ptr(xcodes), ptr(wcodes), cdt, ptr(b), ptr(sv), C.byref(g), ptr(out), ptr(target), ptr(att),
                                  ptr(sse), ptr(ws), stream()), "effq_conv3d_tc_pc"))
        return out, sse
    timer.run(name, {"flops": flops, "bytes": nbytes}, lambda: check(
        lib.effq_conv3d_tc(ptr(xcodes), ptr(wcodes), cdt, ptr(b), cs, C.byref(g), ptr(out), ptr(target), ptr(att),
                           ptr(sse), ptr(ws), stream()), "effq_conv3d_tc"))
    return out, sse


# digit-plane products of the fp32-accurate conv, most significant first (x digit, w digit): 2^-8(s+t) of the leading
# one.  Only (2, 2) is dropped (2^-32): the low digits of a value are as large as those of the channel maximum, so the
# s + t = 3 products are 2^-24 of the LARGEST term but ~2^-20 of a typical one (measured: 1.6e-6 of max|out| without
# them, against 1e-6 .. 2e-7 for the library's fp32 conv).
_FP_PAIRS = ((0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0), (1, 2), (2, 1))


def conv3d_fp_supported(x_shape, c2: int, ksize, stride, padding) -> bool:
    """Shapes the FP conv can run on the tcgen05 kernel: what conv3d_tc takes in bf16, with a channel count the
    digit-plane kernel handles."""
    n, c1, d, h, w = x_shape
    return bool(conv3d_tc_supported(x_shape, c2, ksize, stride, padding, CODE_BF16) and (d * h * w) % 4 == 0 and
                capi.load().effq_split3_ndhwc_supported(int(c1), int(d * h * w)))


def _digits3(t_int: torch.Tensor):
    """int32 tensor -> balanced base-256 digits (d0, d1, d2) with t = d0 * 65536 + d1 * 256 + d2."""
    d2 = ((t_int + 128) & 255) - 128
    t1 = (t_int - d2) >> 8
    d1 = ((t1 + 128) & 255) - 128
    return (t1 - d1) >> 8, d1, d2


def conv3d_fp(x: torch.Tensor, weight: torch.Tensor, bias, ksize, ws=None):
    """The un-quantised convolution of the FP pass (reference src/ptqer.py:333-335: F.conv3d in fp32) on the tcgen05
    kernel, as an EXACT integer computation (the Ozaki splitting): with power-of-two scales per input channel
    (2^ex > max|x_c|, folded into the weights) and per output channel (2^ew > max|w'_row|) both operands become 24-bit
    fixed-point integers, cut into three balanced base-256 digits each; a conv of digit planes is a sum of integers
    below 2^24 and comes out of the fp32 tensor-core accumulator without a rounding, and eight of the nine digit
    products (all but the last, 2^-32) are combined in the epilogue (first launch writes, the others accumulate) with one fp32
    rounding each.  What is left against an exact conv is the 2^-24 fixed-point resolution relative to the channel /
    row maximum -- fp32-conv class, measured against an fp64 conv in tests/test_gpu_kernels.py
    (test_conv3d_fp_matches_fp64).  Three floating-point bf16 planes per operand, the first version, were 10 x less
    accurate: unaligned products lose bits at each of the K/16 accumulation steps (truncation, so the error is a
    bias)."""
    x = _f32c(x, "x")
    n, c1, d, h, w = x.shape
    c2 = weight.shape[0]
    k = capi._triple(ksize)
    pad = tuple((t - 1) // 2 for t in k)
    g = Geom.make((n, c1, d, h, w), c2, k, 1, pad)
    lib = capi.load()
    dev = x.device
    dhw = d * h * w
    amax = torch.zeros(c1, dtype=torch.float32, device=dev)
    timer.run("channel_absmax", {"bytes": 4 * x.numel()}, lambda: check(
        lib.effq_channel_absmax(ptr(x), n, c1, dhw, ptr(amax), stream()), "effq_channel_absmax"))
    ex = torch.where(amax > 0, torch.frexp(amax).exponent, torch.zeros_like(amax, dtype=torch.int32)).int().contiguous()
    xp = torch.empty((3, n, d, h, w, c1), dtype=torch.bfloat16, device=dev)
    timer.run("fixdigits_ndhwc", {"bytes": 10 * x.numel()}, lambda: check(
        lib.effq_fixdigits_ndhwc(ptr(x), n, c1, dhw, ptr(ex), ptr(xp[0]), ptr(xp[1]), ptr(xp[2]), stream()),
        "effq_fixdigits_ndhwc"))
    # weights: fold the input-channel scales, fixed point per output channel (small tensors: host-side tensor ops)
    wf = torch.ldexp(_f32c(weight.detach().float(), "weight"), ex.view(1, c1, 1, 1, 1))
    rmax = wf.abs().amax(dim=(1, 2, 3, 4))
    ew = torch.where(rmax > 0, torch.frexp(rmax).exponent, torch.zeros_like(rmax, dtype=torch.int32)).int()
    w_int = torch.round(torch.ldexp(wf, (23 - ew).view(c2, 1, 1, 1, 1))).int()
    wplanes = [pack_weight_codes(p.float(), CODE_BF16) for p in _digits3(w_int)]
    base = torch.ldexp(torch.ones(c2, dtype=torch.float32, device=dev), ew - 46)
    scales = {st: (base * float(2 ** (8 * (4 - st)))).contiguous() for st in (0, 1, 2, 3)}
    if ws is None:
        ws = workspace(lib.effq_conv3d_tc_workspace(C.byref(g)), dev)
    out = torch.empty((n, c2, d, h, w), dtype=torch.float32, device=dev)
    b = _f32c(bias, "bias") if bias is not None else None
    flops = 2.0 * n * dhw * c2 * c1 * g.taps
    name = f"conv3d_fp_c{c1}x{c2}k{k[0]}"
    # least significant products first: the running sum stays small until the last additions, so the fp32 rounding of
    # the seven accumulating launches is that of one addition at full magnitude
    for i, (px, pw) in enumerate(reversed(_FP_PAIRS)):
        sv = scales[px + pw]
        if i == 0:
            timer.run(name, {"flops": flops}, lambda: check(
                lib.effq_conv3d_tc_pc(ptr(xp[px]), ptr(wplanes[pw]), CODE_BF16, ptr(b), ptr(sv), C.byref(g), ptr(out), None,
                                      None, None, ptr(ws), stream()), "effq_conv3d_tc_pc"))
        else:
            timer.run(name, {"flops": flops}, lambda: check(
                lib.effq_conv3d_tc_acc(ptr(xp[px]), ptr(wplanes[pw]), CODE_BF16, None, None, ptr(sv), C.byref(g), ptr(out),
                                       ptr(ws), stream()), "effq_conv3d_tc_acc"))
    return out


def pack_weight_codes(wcodes_int: torch.Tensor, code_dtype: int = CODE_BF16) -> torch.Tensor:
    """[C2][C1][kd][kh][kw] integer codes (already 2c-(L-1)) on the GPU -> bf16 / e4m3 codes in the
    tensor-core weight layout (the layout effq_admm_project emits during calibration)."""
    w = _f32c(wcodes_int.float(), "wcodes")
    c2, c1 = w.shape[:2]
    taps = w[0, 0].numel()
    out = torch.empty(c2 * c1 * taps, dtype=E4M3 if code_dtype == CODE_E4M3 else torch.bfloat16, device=w.device)
    check(capi.load().effq_pack_wcodes(ptr(w), c2, c1, taps, int(code_dtype), ptr(out), stream()),
          "effq_pack_wcodes")
    return out


# ---------------------------------------------------------------------------
# (a7+a8) normal-equation statistics
# ---------------------------------------------------------------------------
def gram(x, y, att, ksize, stride, padding, has_bias=True, x_scale=None, ws=None):
    """A0 (K'xK'), B0 (C2xK') -- solver.py:282-314 without materialising im2col."""
    x = _f32c(x, "x")
    y = _f32c(y, "y")
    g = Geom.make(x.shape, y.shape[1], ksize, stride, padding)
    k = g.c1 * g.taps
    kp = k + (1 if has_bias else 0)
    lib = capi.load()
    need = lib.effq_gram_workspace(C.byref(g), int(has_bias))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    a0 = torch.empty((kp, kp), dtype=torch.float32, device=x.device)
    b0 = torch.empty((g.c2, kp), dtype=torch.float32, device=x.device)
    if att is not None:
        att = _f32c(att, "att")
    xs = _f32c(x_scale.reshape(1), "x_scale") if x_scale is not None else None
    od, oh, ow = g.out_spatial()
    flops = 2.0 * g.n * od * oh * ow * (kp + g.c2) * kp
    timer.run("gram_f32", {"flops": flops}, lambda: check(
        lib.effq_gram_f32(ptr(x), ptr(xs), ptr(y), ptr(att), C.byref(g), int(has_bias), ptr(a0), ptr(b0),
                          ptr(ws), stream()), "effq_gram_f32"))
    return a0, b0


def gram_f64(x, y, ksize, stride, padding, has_bias=True):
    """Unweighted fp64 statistics [S ; T] ((K'+C2) x K') for conv-free scoring."""
    x = _f32c(x, "x")
    y = _f32c(y, "y")
    g = Geom.make(x.shape, y.shape[1], ksize, stride, padding)
    kp = g.c1 * g.taps + (1 if has_bias else 0)
    acc = torch.empty((kp + g.c2, kp), dtype=torch.float64, device=x.device)
    od, oh, ow = g.out_spatial()
    timer.run("gram_f64", {"flops": 2.0 * g.n * od * oh * ow * (kp + g.c2) * kp}, lambda: check(
        capi.load().effq_gram_f64(ptr(x), ptr(y), C.byref(g), int(has_bias), ptr(acc), stream()), "effq_gram_f64"))
    return acc


def quadform_sse(acc, sum_y2: float, g, bstar, sse, ws):
    """sse <- sum((conv(x,G)+b*-y)^2) from the fp64 statistics (one tiny launch per iterate)."""
    c2, k = g.shape
    check(capi.load().effq_quadform_sse(ptr(acc), float(sum_y2), ptr(g), ptr(bstar), c2, k, int(bstar is not None),
                                        ptr(sse), ptr(ws), stream()), "effq_quadform_sse")


_qf_ws = {}


def quadform_delta(acc, sum_sq: torch.Tensor, g, bstar, sse, g_ref=None, b_ref=None, st: Optional["AdmmState"] = None,
                   numel: float = 0.0, history=None):
    """sse <- sum((conv(x, G) + b* - y)^2) from fp64 statistics, tiled kernel (csrc/quadform.cu).  With
    (g_ref, b_ref) the statistics are those of the residual R = y - conv(x, g_ref) - b_ref and ``sum_sq`` (device
    fp64 scalar) is sum R^2.  ``st``: also do the best-iterate bookkeeping of the scored iterate (numel, history)."""
    c2, k = g.shape
    if sum_sq.dtype != torch.float64 or not sum_sq.is_cuda:
        raise EffqError("quadform_delta: sum_sq must be a CUDA float64 scalar")
    lib = capi.load()
    key = (g.device.type, g.device.index)
    if key not in _qf_ws:
        _qf_ws[key] = workspace(lib.effq_quadform_delta_workspace(c2, k + 1), g.device)
    kp = k + (1 if bstar is not None else 0)
    timer.run(f"quadform_k{kp}", {"flops": 2.0 * c2 * kp * kp}, lambda: check(
        lib.effq_quadform_delta(ptr(acc), ptr(sum_sq), ptr(g), ptr(bstar), ptr(g_ref), ptr(b_ref), c2, k,
                                int(bstar is not None), ptr(sse), ptr(_qf_ws[key]), st.p if st is not None else None,
                                float(numel), ptr(history), stream()), "effq_quadform_delta"))


def gram_tc_f64(xcodes, code_scale, y, att=None, has_bias=True, ws=None, att_exact=False):
    """[X^ diag(att) X^T ; Y diag(att) X^T] as one full fp64 matrix in real units (operand of quadform_delta),
    from the tcgen05 Gram kernel.  Returns (acc64, workspace, abort flag)."""
    y = _f32c(y, "y")
    n, d, h, w, c1 = xcodes.shape
    g = Geom.make((n, c1, d, h, w), y.shape[1], 3, 1, 1)
    kp = g.c1 * 27 + (1 if has_bias else 0)
    lib = capi.load()
    need = lib.effq_gram_workspace(C.byref(g), int(has_bias))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=y.device)
    acc = torch.empty((kp + g.c2, kp), dtype=torch.float64, device=y.device)
    if att is not None:
        att = _f32c(att, "att")
    cs = _f32c(code_scale.reshape(1), "code_scale")
    flops = 2.0 * g.n * d * h * w * (kp + g.c2) * kp
    timer.run("gram_tc_f64", {"flops": flops}, lambda: check(
        lib.effq_gram_tc_f64(ptr(xcodes), ptr(cs), ptr(y), ptr(att), C.byref(g), int(has_bias), int(bool(att_exact)),
                             ptr(acc), ptr(ws), stream()), "effq_gram_tc_f64"))
    flag = ws[need - 16:need - 12].view(torch.int32)
    return acc, ws, flag


def gram_tc_dual(xcodes, code_scale, y, att, ws=None, att_exact=False):
    """A0, B0 (as gram_tc, with bias) and the UNWEIGHTED statistics matrix [(K'+C2) x K'] fp64 whose first K' rows
    hold S = X^ X^T, from one pass over the codes.  Returns (a0, b0, stats64, workspace, abort flag)."""
    y = _f32c(y, "y")
    n, d, h, w, c1 = xcodes.shape
    g = Geom.make((n, c1, d, h, w), y.shape[1], 3, 1, 1)
    kp = g.c1 * 27 + 1
    lib = capi.load()
    one = (lib.effq_gram_workspace(C.byref(g), 1) + 255) // 256 * 256
    if ws is None or ws.numel() < 2 * one:
        ws = torch.empty(2 * one, dtype=torch.uint8, device=y.device)
    a0 = torch.empty((kp, kp), dtype=torch.float32, device=y.device)
    b0 = torch.empty((g.c2, kp), dtype=torch.float32, device=y.device)
    stats = torch.empty((kp + g.c2, kp), dtype=torch.float64, device=y.device)
    if att is not None:
        att = _f32c(att, "att")
    cs = _f32c(code_scale.reshape(1), "code_scale")
    vox = float(g.n) * d * h * w
    flops = 2.0 * vox * ((kp + g.c2) * kp + kp * kp)                       # full count of both Grams + B0
    flops_alg = vox * (2.0 * kp * (kp + 1) + 2.0 * g.c2 * kp)             # symmetric half per Gram, B0 in full
    timer.run("gram_tc_dual", {"flops": flops, "flops_alg": flops_alg}, lambda: check(
        lib.effq_gram_tc_dual(ptr(xcodes), ptr(cs), ptr(y), ptr(att), C.byref(g), int(bool(att_exact)), ptr(a0), ptr(b0),
                              ptr(stats), ptr(ws), stream()), "effq_gram_tc_dual"))
    off = (kp + g.c2) * kp * 8
    flag = ws[off:off + 4].view(torch.int32)
    return a0, b0, stats, ws, flag


def gram_tc_rows_f64(xcodes, code_scale, y, stats, ws=None):
    """Rows [K', K'+C2) of ``stats`` <- unweighted Y X^T for the target y (rows-only pass)."""
    y = _f32c(y, "y")
    n, d, h, w, c1 = xcodes.shape
    g = Geom.make((n, c1, d, h, w), y.shape[1], 3, 1, 1)
    kp = g.c1 * 27 + 1
    lib = capi.load()
    need = lib.effq_gram_workspace(C.byref(g), 1)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=y.device)
    cs = _f32c(code_scale.reshape(1), "code_scale")
    timer.run("gram_tc_rows", {"flops": 2.0 * g.n * d * h * w * g.c2 * kp}, lambda: check(
        lib.effq_gram_tc_rows_f64(ptr(xcodes), ptr(cs), ptr(y), C.byref(g), ptr(stats), ptr(ws), stream()),
        "effq_gram_tc_rows_f64"))
    flag = ws[need - 16:need - 12].view(torch.int32)
    return ws, flag


def gram_tc_supported(x_shape, c2, ksize, stride, padding) -> bool:
    g = Geom.make(x_shape, c2, ksize, stride, padding)
    return bool(capi.load().effq_gram_tc_supported(C.byref(g)))


_att_exact_cache = {}


def att_is_exact(att: Optional[torch.Tensor], code_max: int) -> bool:
    """True when every att * code (code <= code_max) is exactly representable in bf16: integer-valued
    weights -- the reference truncates its attention map to integers (ptqer.py:160-163) -- with
    max(att) * code_max <= 256.  One device reduction + read-back per attention map, remembered for as long
    as that tensor OBJECT lives and is not written to (a weak reference, not the address: the caching
    allocator hands the address of a freed map to the next one of the same size)."""
    if att is None:
        return True
    key = (id(att), int(code_max))
    hit = _att_exact_cache.get(key)
    if hit is not None and hit[0]() is att and hit[1] == att._version:
        return hit[2]
    for k in [k for k, v in _att_exact_cache.items() if v[0]() is None]:        # entries of dead tensors
        del _att_exact_cache[k]
    st = torch.stack([(att == att.round()).all().float(), att.max(), att.min()]).cpu().tolist()
    ok = bool(st[0] == 1.0 and st[2] >= 0.0 and st[1] * code_max <= 256.0)
    _att_exact_cache[key] = (weakref.ref(att), att._version, ok)
    return ok


def gram_tc(xcodes, code_scale, y, att, has_bias=True, ws=None, att_exact=False):
    """A0, B0 on the tensor cores from the NDHWC codes (3x3x3, stride 1, pad 1).  ``att_exact``:
    see att_is_exact (one bf16 term instead of hi + lo for the weighted codes)."""
    y = _f32c(y, "y")
    n, d, h, w, c1 = xcodes.shape
    g = Geom.make((n, c1, d, h, w), y.shape[1], 3, 1, 1)
    k = g.c1 * 27
    kp = k + (1 if has_bias else 0)
    lib = capi.load()
    need = lib.effq_gram_workspace(C.byref(g), int(has_bias))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=y.device)
    a0 = torch.empty((kp, kp), dtype=torch.float32, device=y.device)
    b0 = torch.empty((g.c2, kp), dtype=torch.float32, device=y.device)
    if att is not None:
        att = _f32c(att, "att")
    cs = _f32c(code_scale.reshape(1), "code_scale")
    od, oh, ow = g.out_spatial()
    flops = 2.0 * g.n * od * oh * ow * (kp + g.c2) * kp
    flops_alg = float(g.n) * od * oh * ow * (kp * (kp + 1) + 2.0 * g.c2 * kp)
    timer.run("gram_tc", {"flops": flops, "flops_alg": flops_alg}, lambda: check(
        lib.effq_gram_tc(ptr(xcodes), ptr(cs), ptr(y), ptr(att), C.byref(g), int(has_bias), int(bool(att_exact)),
                         ptr(a0), ptr(b0), ptr(ws), stream()), "effq_gram_tc"))
    flag = ws[need - 16:need - 12].view(torch.int32)       # non-zero if the tcgen05 kernel aborted
    return a0, b0, ws, flag


# ---------------------------------------------------------------------------
# (a9, a11) ADMM update
# ---------------------------------------------------------------------------
def admm_rhs(b0, w0p, g, dual, rho: float, eta: float, out, planes=None):
    """B = B0 + eta W0' (+ rho (G - dual) on the weight columns) as fp32 ``out`` and / or as the three
    bf16 terms ``planes`` ([3][C2][split3_ld(K')]) that solve_gemm_tc consumes."""
    c2, kp = b0.shape
    k = g.numel() // c2
    check(capi.load().effq_admm_rhs(ptr(b0), ptr(w0p), ptr(g), ptr(dual), float(rho), float(eta), c2, k,
                                    int(kp != k), ptr(out), ptr(planes), stream()), "effq_admm_rhs")
    return out


def split3_ld(cols: int) -> int:
    return int(capi.load().effq_split3_ld(int(cols)))


def split3_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows][cols] -> bf16 [3][rows][split3_ld(cols)] with x = p0 + p1 + p2 (tail columns zero)."""
    if x.dim() != 2 or x.dtype != torch.float32 or not x.is_cuda:
        raise EffqError("split3_bf16: expected a CUDA float32 matrix")
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    if out is None:
        out = torch.empty((3, rows, split3_ld(cols)), dtype=torch.bfloat16, device=x.device)
    elif tuple(out.shape) != (3, rows, split3_ld(cols)) or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise EffqError("split3_bf16: out must be a contiguous bf16 [3][rows][split3_ld(cols)] tensor")
    check(capi.load().effq_split3_bf16(ptr(x), rows, cols, x.stride(0), ptr(out), stream()), "effq_split3_bf16")
    return out


def solve_gemm_tc(a_planes: torch.Tensor, b_planes: torch.Tensor, k: int, out: Optional[torch.Tensor] = None,
                  ws: Optional[torch.Tensor] = None):
    """out[m][n] = A B^T from split planes (fp32-class accuracy on the tensor cores); A^-1 symmetric ->
    pass its planes as ``b_planes`` to get B A^-1.  Returns (out, workspace)."""
    if a_planes.dtype != torch.bfloat16 or b_planes.dtype != torch.bfloat16 or a_planes.shape[2] != b_planes.shape[2]:
        raise EffqError("solve_gemm_tc: operands must be bf16 planes with the same row pitch")
    m, n = a_planes.shape[1], b_planes.shape[1]
    lib = capi.load()
    if out is None:
        ldo = (n + 3) // 4 * 4
        out = torch.empty((m, ldo), dtype=torch.float32, device=a_planes.device)[:, :n]
    ldo = out.stride(0)
    need = lib.effq_solve_gemm_tc_workspace(m, n, int(k), ldo)
    if ws is None or ws.numel() < need:
        ws = workspace(need, a_planes.device)
    timer.run("solve_gemm_tc", {"flops": 2.0 * m * n * k}, lambda: check(
        lib.effq_solve_gemm_tc(ptr(a_planes), ptr(b_planes), m, n, int(k), ptr(out), ldo, ptr(ws), stream()),
        "effq_solve_gemm_tc"))
    return out, ws


def gemm_tc_planes(a_planes: torch.Tensor, b_planes: torch.Tensor, k: int, out: torch.Tensor, ws: Optional[torch.Tensor] = None,
                   tri: int = 0):
    """out[m][n] = A B^T from whole split-plane matrices ([3][rows][ld] bf16) through the general entry point;
    ``tri`` 1 / 2: B holds the rows of a lower triangular matrix / of its transpose (zero K range skipped)."""
    m, n = a_planes.shape[1], b_planes.shape[1]
    lib = capi.load()
    ldo = out.stride(0)
    need = lib.effq_gemm_tc_ex_workspace(m, n, int(k), ldo)
    if ws is None or ws.numel() < need:
        ws = workspace(need, a_planes.device)
    timer.run("solve_gemm_tc", {"flops": 2.0 * m * n * k * (0.5 if tri else 1.0)}, lambda: check(
        lib.effq_gemm_tc_ex(ptr(a_planes), a_planes.shape[2], m * a_planes.shape[2], ptr(b_planes), b_planes.shape[2],
                            n * b_planes.shape[2], m, n, int(k), 1.0, 0.0, None, 0, ptr(out), ldo, int(tri) << 1, ptr(ws),
                            stream()), "effq_gemm_tc_ex"))
    return out, ws


def admm_lhs(a0, rho: float, eta: float, has_bias: bool, out):
    kp = a0.shape[0]
    check(capi.load().effq_admm_lhs(ptr(a0), float(rho), float(eta), kp, int(has_bias), ptr(out), stream()),
          "effq_admm_lhs")
    return out


def admm_project(wstar, dual, wstate: ScaleState, xstate: Optional[ScaleState], nlvl_w: int, nlvl_a: int,
                 c2: int, c1: int, taps: int, has_bias: bool, dual_div: float, g_out, bstar_out, wcodes_out,
                 st: AdmmState, next_rhs=None, keep=None, per_channel_out=None):
    """``next_rhs`` = (b0, w0p, rho_next, eta, planes): also emit the next iteration's right-hand side
    as split planes (replaces the next admm_rhs launch).  ``keep`` = (best_g, best_b, best_wcodes): save the
    previous iterate (still in g_out / bstar_out / wcodes_out) first if the step that scored it made it the best."""
    kp_ = None
    if keep is not None:
        bg, bb, bw = keep[:3]
        bpc = keep[3] if len(keep) > 3 else None
        kp_ = C.byref(capi.AdmmKeep(bg.data_ptr(), bb.data_ptr() if bb is not None else None,
                                    bw.data_ptr() if bw is not None else None,
                                    bpc.data_ptr() if bpc is not None else None))
    ldw = wstar.stride(0)
    nx = None
    if next_rhs is not None:
        b0, w0p, rho_n, eta, planes = next_rhs
        nx = C.byref(capi.NextRhs(b0.data_ptr(), w0p.data_ptr(), float(rho_n), float(eta), planes.data_ptr()))
    check(capi.load().effq_admm_project(ptr(wstar), ldw, ptr(dual), wstate.p, xstate.p if xstate else None,
                                        int(nlvl_w), int(nlvl_a), c2, c1, taps, int(has_bias), float(dual_div),
                                        ptr(g_out), ptr(bstar_out), ptr(wcodes_out),
                                        code_dtype_of(wcodes_out) if wcodes_out is not None else 0, st.p, nx, kp_,
                                        ptr(per_channel_out), stream()),
          "effq_admm_project")


def admm_decide(st: AdmmState, sse, numel: float, history, comm=None):
    """Best-iterate bookkeeping of the iterate whose squared error is in ``sse`` (all-reduced in-kernel with
    ``comm``); the copy of the best iterate happens in the next admm_project (``keep``) / admm_keep."""
    check(capi.load().effq_admm_decide(st.p, ptr(sse), float(numel), ptr(history), comm, stream()), "effq_admm_decide")


def admm_keep(st: AdmmState, g, bstar, best_g, best_b, aux_src=None, aux_dst=None, pc_src=None, pc_dst=None):
    nb = aux_src.numel() * aux_src.element_size() if aux_src is not None else 0
    check(capi.load().effq_admm_keep(st.p, ptr(g), ptr(bstar), g.numel(), g.shape[0], ptr(best_g), ptr(best_b),
                                     ptr(aux_src), ptr(aux_dst), nb, ptr(pc_src), ptr(pc_dst), stream()), "effq_admm_keep")


def admm_track(st: AdmmState, sse, numel: float, g, bstar, best_g, best_b, history, aux_src=None, aux_dst=None,
               comm=None):
    """``comm`` (dist.PeerLink.comm_ptr) makes the kernel all-reduce this rank's SSE share over NVLink."""
    nb = aux_src.numel() * aux_src.element_size() if aux_src is not None else 0
    check(capi.load().effq_admm_track(st.p, ptr(sse), float(numel), ptr(g), ptr(bstar), g.numel(),
                                      g.shape[0], ptr(best_g), ptr(best_b), ptr(history), ptr(aux_src),
                                      ptr(aux_dst), nb, comm, stream()), "effq_admm_track")


# ---------------------------------------------------------------------------
# (a15) end-to-end activation-range refinement: STE backward + Adam
# ---------------------------------------------------------------------------
_ste_ws = {}


def fakequant_ste_bwd(x: torch.Tensor, grad_out: torch.Tensor, alpha: torch.Tensor, nlvl: int, lo: float, hi: float,
                      grad_alpha_acc: torch.Tensor, want_grad_x: bool = True) -> Optional[torch.Tensor]:
    """Backward of discretize(x/alpha)*alpha under the reference's STE (layer_helper.py:13-37):
    returns grad_x (or None) and ADDS d loss / d alpha into the 0-dim/1-element fp64 ``grad_alpha_acc``."""
    x = _f32c(x, "x")
    grad_out = _f32c(grad_out, "grad_out")
    alpha = _f32c(alpha.reshape(1), "alpha")
    if grad_alpha_acc.dtype != torch.float64 or not grad_alpha_acc.is_cuda:
        raise EffqError("grad_alpha_acc must be a CUDA float64 tensor")
    lib = capi.load()
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    if key not in _ste_ws:
        _ste_ws[key] = workspace(lib.effq_ste_bwd_workspace(), x.device)
    gx = torch.empty_like(x) if want_grad_x else None
    timer.run("fakequant_ste_bwd", {"bytes": x.numel() * (8 + (4 if want_grad_x else 0))}, lambda: check(
        lib.effq_fakequant_ste_bwd(ptr(x), ptr(grad_out), x.numel(), ptr(alpha), float(lo), float(hi), int(nlvl),
                                   ptr(gx), ptr(grad_alpha_acc), ptr(_ste_ws[key]), stream()),
        "effq_fakequant_ste_bwd"))
    return gx


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    """torch.optim.Adam's update of a flat fp32 parameter vector from fp64 gradients, one launch."""
    if params.dtype != torch.float32 or grads.dtype != torch.float64 or not params.is_contiguous():
        raise EffqError("adam_step: params fp32 contiguous, grads fp64")
    check(capi.load().effq_adam_step(ptr(params), ptr(grads), float(grad_scale), ptr(exp_avg), ptr(exp_avg_sq),
                                     params.numel(), float(lr), float(beta1), float(beta2), float(eps), int(step),
                                     stream()), "effq_adam_step")


def split3_ndhwc(x: torch.Tensor):
    """NCDHW fp32 -> (hi, mid, lo) NDHWC bf16 planes with hi + mid + lo == x exactly (operand of the tensor-core
    dgrad), one fused pass; None when the shape is not handled (caller falls back to elementwise ops)."""
    x = _f32c(x, "x")
    n, c, d, h, w = x.shape
    lib = capi.load()
    if not lib.effq_split3_ndhwc_supported(c, d * h * w):
        return None
    planes = [torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=x.device) for _ in range(3)]
    timer.run("split3_ndhwc", {"bytes": 10 * x.numel()}, lambda: check(
        lib.effq_split3_ndhwc(ptr(x), n, c, d * h * w, ptr(planes[0]), ptr(planes[1]), ptr(planes[2]), stream()),
        "effq_split3_ndhwc"))
    return planes


# ---------------------------------------------------------------------------
# (f.4) glue ops between the quantizer layers: one HBM pass each, the neighbouring elementwise op fused in
# ---------------------------------------------------------------------------
def relu(x: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    """nn.ReLU of a unit (factoryQ.py:66-81)."""
    x = _f32c(x, "x")
    y = x if inplace else torch.empty_like(x)
    if x.numel() == 0:
        return y
    timer.run("glue_relu", {"bytes": 8 * x.numel()}, lambda: check(
        capi.load().effq_glue_elementwise(ptr(x), None, x.numel(), 1, ptr(y), stream()), "effq_glue_elementwise"))
    return y


def add(a: torch.Tensor, b: torch.Tensor, relu_after: bool = False) -> torch.Tensor:
    """Residual add (factory_blk.py:147-166), optionally followed by a ReLU in the same pass."""
    a, b = _f32c(a, "a"), _f32c(b, "b")
    if a.shape != b.shape:
        raise EffqError(f"add: shapes differ ({tuple(a.shape)} vs {tuple(b.shape)})")
    y = torch.empty_like(a)
    if a.numel() == 0:
        return y
    timer.run("glue_add", {"bytes": 12 * a.numel()}, lambda: check(
        capi.load().effq_glue_elementwise(ptr(a), ptr(b), a.numel(), 1 if relu_after else 0, ptr(y), stream()),
        "effq_glue_elementwise"))
    return y


def maxpool3d(x: torch.Tensor, kernel, relu_after: bool = False) -> torch.Tensor:
    """nn.MaxPool3d(kernel, kernel) (factory_blk.py:18-42) [+ the ReLU of the unit that follows]."""
    x = _f32c(x, "x")
    n, c, d, h, w = x.shape
    kd, kh, kw = capi._triple(kernel)
    y = torch.empty((n, c, d // kd, h // kh, w // kw), dtype=torch.float32, device=x.device)
    if y.numel() == 0:
        return y
    timer.run("glue_maxpool", {"bytes": 4 * (x.numel() + y.numel())}, lambda: check(
        capi.load().effq_glue_maxpool3d(ptr(x), n * c, d, h, w, kd, kh, kw, 1 if relu_after else 0, ptr(y), stream()),
        "effq_glue_maxpool3d"))
    return y


def upsample_trilinear(x: torch.Tensor, factor, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Upsample(scale_factor=factor, mode="trilinear") [+ skip] (factory_blk.py:45-93), integer factors."""
    x = _f32c(x, "x")
    n, c, d, h, w = x.shape
    fd, fh, fw = capi._triple(factor)
    y = torch.empty((n, c, d * fd, h * fh, w * fw), dtype=torch.float32, device=x.device)
    if skip is not None:
        skip = _f32c(skip, "skip")
        if skip.shape != y.shape:
            raise EffqError(f"upsample_trilinear: skip has shape {tuple(skip.shape)}, output {tuple(y.shape)}")
    if y.numel() == 0:
        return y
    timer.run("glue_upsample", {"bytes": 4 * (x.numel() + y.numel() * (2 if skip is not None else 1))}, lambda: check(
        capi.load().effq_glue_upsample_trilinear(ptr(x), ptr(skip), n * c, d, h, w, fd, fh, fw, ptr(y), stream()),
        "effq_glue_upsample_trilinear"))
    return y
