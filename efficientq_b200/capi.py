"""ctypes binding of ``libeffq_b200.so`` (the C-ABI declared in ``include/effq_b200.h``).

There is NO fallback: if the shared library is missing or an entry point returns an
error the call raises.  ``torch`` is used only to own device memory and streams; every
entry point receives raw device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeffq_b200.so")


class EffqError(RuntimeError):
    pass


class PeerComm(C.Structure):
    """``effq_peer_comm`` (include/effq_b200.h)."""
    _fields_ = [("slots", C.c_void_p * 8), ("rank", C.c_int32), ("world", C.c_int32)]


class NextRhs(C.Structure):
    """``effq_next_rhs`` (include/effq_b200.h)."""
    _fields_ = [("b0", C.c_void_p), ("w0p", C.c_void_p), ("rho", C.c_float), ("eta", C.c_float), ("planes", C.c_void_p)]


class AdmmKeep(C.Structure):
    """``effq_admm_keep_bufs`` (include/effq_b200.h)."""
    _fields_ = [("best_g", C.c_void_p), ("best_b", C.c_void_p), ("best_wcodes", C.c_void_p), ("best_pc", C.c_void_p)]


class Geom(C.Structure):
    """``effq_geom`` (include/effq_b200.h)."""
    _fields_ = [(n, C.c_int32) for n in
                ("n", "c1", "d", "h", "w", "c2", "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph", "pw")]

    @classmethod
    def make(cls, x_shape, c2, ksize, stride, padding) -> "Geom":
        n, c1, d, h, w = [int(t) for t in x_shape]
        k = _triple(ksize)
        s = _triple(stride)
        p = _triple(padding)
        return cls(n, c1, d, h, w, int(c2), k[0], k[1], k[2], s[0], s[1], s[2], p[0], p[1], p[2])

    def out_spatial(self):
        return ((self.d + 2 * self.pd - self.kd) // self.sd + 1,
                (self.h + 2 * self.ph - self.kh) // self.sh + 1,
                (self.w + 2 * self.pw - self.kw) // self.sw + 1)

    @property
    def taps(self):
        return self.kd * self.kh * self.kw


SCALE_STATE_BYTES = 48     # sizeof(effq_scale_state)
ADMM_STATE_BYTES = 40      # sizeof(effq_admm_state)


def _triple(v):
    return (int(v),) * 3 if isinstance(v, int) else tuple(int(t) for t in v)


_SIGS = {
    "effq_abi_version": (C.c_int, []),
    "effq_last_error": (C.c_char_p, []),
    "effq_launch_count": (C.c_uint64, []),
    "effq_reset_launch_count": (None, []),
    "effq_fakequant_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_fakequant_state": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                       C.c_void_p, C.c_void_p]),
    "effq_quantize_act_ndhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_scale_search_workspace": (C.c_int64, [C.c_int64]),
    "effq_scale_search": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                    C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_void_p]),
    "effq_scale_partial": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                     C.c_float, C.c_float, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "effq_scale_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "effq_conv3d_f32_workspace": (C.c_int64, [C.POINTER(Geom)]),
    "effq_conv3d_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_conv3d_tc_supported": (C.c_int, [C.POINTER(Geom), C.c_int32]),
    "effq_conv3d_tc_workspace": (C.c_int64, [C.POINTER(Geom)]),
    "effq_conv3d_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_conv3d_tc_acc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom),
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_channel_absmax": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "effq_fixdigits_ndhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "effq_conv3d_tc_pc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_pack_wcodes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_gram_workspace": (C.c_int64, [C.POINTER(Geom), C.c_int32]),
    "effq_gram_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_int32,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_gram_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_quadform_sse": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_quadform_delta_workspace": (C.c_int64, [C.c_int32, C.c_int32]),
    "effq_quadform_delta": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                      C.c_void_p]),
    "effq_gram_tc_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_gram_tc_accumulate2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "effq_gram_tc_dual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_gram_tc_rows_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "effq_gram_tc_supported": (C.c_int, [C.POINTER(Geom)]),
    "effq_gram_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom),
                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_gram_tc_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Geom), C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_admm_rhs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_split3_ld": (C.c_int64, [C.c_int64]),
    "effq_split3_bf16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "effq_solve_gemm_tc_workspace": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "effq_solve_gemm_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p]),
    "effq_gemm_tc_ex_workspace": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int64]),
    "effq_gemm_tc_ex": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_void_p]),
    "effq_potrf_tile": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_void_p]),
    "effq_split3_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_int32, C.c_int32, C.c_void_p]),
    "effq_admm_lhs": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_admm_project": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_scale_search_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                         C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "effq_admm_decide": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_admm_keep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_admm_track": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p]),
    "effq_ste_bwd_workspace": (C.c_int64, []),
    "effq_fakequant_ste_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "effq_split3_ndhwc_supported": (C.c_int, [C.c_int32, C.c_int64]),
    "effq_split3_ndhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "effq_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_float,
                                 C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "effq_glue_elementwise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_glue_maxpool3d": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_glue_upsample_trilinear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "effq_peer_bytes": (C.c_int64, []),
    "effq_peer_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p]),
    "effq_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "effq_peer_close": (C.c_int, [C.c_void_p]),
    "effq_peer_free": (C.c_int, [C.c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib: Optional[C.CDLL] = None


class _Recorder:
    """Stands in for the library while a launch sequence is being recorded: every entry point still runs,
    and (function, converted arguments, tag) is appended to ``calls`` so that the same sequence can be
    re-issued later without the Python wrappers around it (layer_engine: the ADMM iterations of one rho block
    launch the same kernels on the same buffers; re-issuing them costs ~2 us per launch instead of ~17 us)."""

    def __init__(self, lib):
        self._lib = lib
        self.calls = []
        self.tag = None

    def __getattr__(self, name):
        fn = getattr(self._lib, name)

        def rec(*args):
            if fn.restype is C.c_int and fn.argtypes and fn.argtypes[-1] is C.c_void_p and not name.endswith("_supported"):
                self.calls.append((fn, args, self.tag, name))      # launches only, not size / support queries
            return fn(*args)
        return rec


_recorder: Optional[_Recorder] = None


class record:
    """``with capi.record() as rec: ...`` -- see _Recorder."""

    def __enter__(self):
        global _recorder
        load()
        _recorder = _Recorder(_lib)
        return _recorder

    def __exit__(self, *exc):
        global _recorder
        _recorder = None
        return False


def load(path: Optional[str] = None) -> C.CDLL:
    """dlopen the kernel library and type every entry point.  Raises if it is absent."""
    global _lib
    if _recorder is not None and path is None:
        return _recorder
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EffqError(
            f"{path} not found: build it with `python -m efficientq_b200.build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.effq_abi_version() != 1:
        raise EffqError("libeffq_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().effq_last_error()
        raise EffqError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def ptr(t: Optional[torch.Tensor]):
    """Raw device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise EffqError("effq_b200 kernels take CUDA tensors only (no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().effq_launch_count())


def reset_launch_count() -> None:
    load().effq_reset_launch_count()
