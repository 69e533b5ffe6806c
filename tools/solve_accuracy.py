"""Accuracy of the proximal step  w* = B A^-1  on real layer systems (golden fixtures), every chain against an fp64
solve of the SAME fp32 A and B:
  lu32    torch.linalg.solve in fp32 -- what the reference does every iteration (solver.py:331)
  lib     library Cholesky + inverse, then the tensor-core split GEMM (the round-1 chain)
  own     blocked Cholesky + block triangular inverse + W^T W on the repo's kernels, then the same GEMM
    python tools/solve_accuracy.py > profiles/r02_solve_accuracy.txt"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientq_b200 import ops  # noqa: E402
from efficientq_b200.spd_inverse import SpdInverter  # noqa: E402

DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
inv_own = SpdInverter(torch.device(DEV))
print(f"{'system':28s} {'rho':>6s} {'lu32 (reference)':>18s} {'lib inverse + GEMM':>20s} {'own inverse + GEMM':>20s}   (max |err| / max |w*|)")
for fname, name in (("layers.npz", "w4a4_k3"), ("layers.npz", "w2a2_k3"), ("layers_wide.npz", "w4a4_k3_c32"),
                    ("layers_wide.npz", "w2a4_k3_c64")):
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", fname))
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    x, w, b, y, att = [torch.from_numpy(g[f"{name}_{t}"]).to(DEV) for t in ("x", "w", "b", "y", "att")]
    c2, c1 = w.shape[:2]
    kk = c1 * k ** 3
    kp = kk + 1
    st = ops.ScaleState(torch.device(DEV))
    ops.scale_search(x, la, 0.0, 1.0, st)
    qx = ops.fakequant_state(x, st, la, 0.0, 1.0)
    a0, b0 = ops.gram(qx, y, att, (k, k, k), s, p, has_bias=True)
    rs = max(y.numel() * y.std().item() / (w.numel() * w.std().item()), 1.0) * att.mean().item()
    for mult in (10.0, 160.0):
        rho, eta = mult * rs, 1.0 * rs
        a = torch.empty_like(a0)
        ops.admm_lhs(a0, rho, eta, True, a)
        w0p = torch.cat([w.reshape(c2, kk), b.reshape(c2, 1)], 1).contiguous()
        gq = w.reshape(c2, kk).clone()
        dual = 0.01 * torch.randn_like(gq)
        bmat = torch.empty(c2, kp, device=DEV)
        planes = torch.empty((3, c2, ops.split3_ld(kp)), dtype=torch.bfloat16, device=DEV)
        ops.admm_rhs(b0, w0p, gq, dual, rho, eta, bmat, planes=planes)
        ref = torch.linalg.solve(a.double(), bmat.double().T).T
        lu32 = torch.linalg.solve(a, bmat.T).T
        ainv_lib = torch.cholesky_inverse(torch.linalg.cholesky(a))
        tc_lib, _ = ops.solve_gemm_tc(planes, ops.split3_bf16(ainv_lib.contiguous()), kp)
        ainv_own, info = inv_own.invert(a)
        tc_own, _ = ops.solve_gemm_tc(planes, ops.split3_bf16(ainv_own.contiguous()), kp)
        torch.cuda.synchronize()
        assert int(info.item()) == 0
        errs = [((v.double() - ref).abs().max() / ref.abs().max()).item() for v in (lu32, tc_lib, tc_own)]
        e_inv = ((ainv_own.double() - torch.linalg.inv(a.double())).abs().max() / ainv_lib.abs().max()).item()
        e_inv_lib = ((ainv_lib.double() - torch.linalg.inv(a.double())).abs().max() / ainv_lib.abs().max()).item()
        print(f"{name + ' K=' + str(kp):28s} {mult:6.0f} {errs[0]:18.2e} {errs[1]:20.2e} {errs[2]:20.2e}   "
              f"| A^-1 itself: lib {e_inv_lib:.1e} own {e_inv:.1e} | cond(A) {torch.linalg.cond(a.double()).item():.1e}")
