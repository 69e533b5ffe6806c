"""Accuracy of the proximal-step product B A^-1 on a real layer system (golden fixture): library fp32
SGEMM vs the tensor-core split GEMM, both against fp64 on the SAME fp32 inputs.
    python tools/solve_accuracy.py [fixture-name]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientq_b200 import ops  # noqa: E402

DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "w2a4_k3_c64"
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "layers_wide.npz"))
k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
x, w, b, y, att = [torch.from_numpy(g[f"{name}_{t}"]).to(DEV) for t in ("x", "w", "b", "y", "att")]
c2, c1 = w.shape[:2]
kk = c1 * k ** 3
kp = kk + 1
st = ops.ScaleState(torch.device(DEV))
ops.scale_search(x, la, 0.0, 1.0, st)
qx = ops.fakequant_state(x, st, la, 0.0, 1.0)
a0, b0 = ops.gram(qx, y, att, (k, k, k), s, p, has_bias=True)
torch.backends.cuda.matmul.allow_tf32 = False
rs = max(y.numel() * y.std().item() / (w.numel() * w.std().item()), 1.0) * att.mean().item()
for mult in (10.0, 160.0):
    rho, eta = mult * rs, 1.0 * rs
    a = torch.empty_like(a0)
    ops.admm_lhs(a0, rho, eta, True, a)
    ainv = torch.cholesky_inverse(torch.linalg.cholesky(a))
    ainv_rm = ainv if ainv.stride(1) == 1 else ainv.T
    w0p = torch.cat([w.reshape(c2, kk), b.reshape(c2, 1)], 1).contiguous()
    gq = w.reshape(c2, kk).clone()
    dual = 0.01 * torch.randn_like(gq)
    bmat = torch.empty(c2, kp, device=DEV)
    planes = torch.empty((3, c2, ops.split3_ld(kp)), dtype=torch.bfloat16, device=DEV)
    ops.admm_rhs(b0, w0p, gq, dual, rho, eta, bmat, planes=planes)
    ref = bmat.double() @ ainv_rm.double()
    lib = bmat @ ainv_rm
    tc, _ = ops.solve_gemm_tc(planes, ops.split3_bf16(ainv_rm), kp)
    torch.cuda.synchronize()
    for nm, v in (("sgemm", lib), ("tensor-core split", tc)):
        e = (v.double() - ref).abs()
        print(f"{name} rho={mult}*rs  {nm:18s} max|err|/max|ref| {e.max().item() / ref.abs().max().item():.2e}   "
              f"rms err / rms ref {e.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item():.2e}   "
              f"max |err| / row rms {(e.max(1).values / ref.pow(2).mean(1).sqrt()).max().item():.2e}")
    print("   cond-ish: max|Ainv| %.3e  max|B| %.3e  max|ref| %.3e  sum|B||Ainv| / |ref| (median) %.1f" % (
        ainv.abs().max().item(), bmat.abs().max().item(), ref.abs().max().item(),
        ((bmat.abs().double() @ ainv_rm.abs().double()) / ref.abs().clamp_min(1e-30)).median().item()))
