"""Bring-up: run one layer through LayerCalibrator with every ops.* call checked for non-finite outputs.
Usage: python tools/bringup/nan_trace.py N C1 C2 D H W K LW LA [n_iter]"""
import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from efficientq_b200 import layer_engine, ops
torch.backends.cuda.matmul.allow_tf32 = False
n, c1, c2, d, h, w, k, lw, la = [int(v) for v in sys.argv[1:10]]
n_iter = int(sys.argv[10]) if len(sys.argv) > 10 else 3
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
x = torch.relu(torch.randn(n, c1, d, h, w, generator=g)).to(dev)
wt = (torch.randn(c2, c1, k, k, k, generator=g) * (2.0 / (c1 * k ** 3)) ** 0.5).to(dev)
b = (torch.randn(c2, generator=g) * 0.05).to(dev)
y = F.conv3d(x, wt, b, 1, (k - 1) // 2)
att = ((torch.rand(n, d, h, w, generator=g) * 3).floor() + 1.0).to(dev)


def finite(v):
    if isinstance(v, torch.Tensor):
        if v.dtype in (torch.float32, torch.float64, torch.bfloat16):
            return bool(torch.isfinite(v.float() if v.dtype == torch.bfloat16 else v).all())
        return True
    if isinstance(v, (tuple, list)):
        return all(finite(t) for t in v)
    return True


seen = {}
def wrap(name, fn):
    def inner(*a, **kw):
        out = fn(*a, **kw)
        torch.cuda.synchronize()
        ok = finite(out) and finite([t for t in a if isinstance(t, torch.Tensor)]) and finite(kw.get("out")) and finite(kw.get("sse")) and finite(kw.get("planes"))
        seen[name] = seen.get(name, 0) + 1
        if not ok and name not in bad:
            bad.append(name)
            print(f"NON-FINITE after {name} (call #{seen[name]})", [(tuple(t.shape), bool(torch.isfinite(t.float()).all())) for t in list(a) + [kw.get('out'), kw.get('sse')] if isinstance(t, torch.Tensor) and t.is_floating_point()], flush=True)
        return out
    return inner
bad = []
for name in ["fakequant_state", "quantize_act_ndhwc", "scale_search", "conv3d_f32", "conv3d_tc", "gram", "gram_tc", "gram_f64",
             "admm_rhs", "split3_bf16", "solve_gemm_tc", "admm_lhs", "admm_project", "admm_track"]:
    setattr(ops, name, wrap(name, getattr(ops, name)))
for name in ["cholesky_ex", "solve_triangular"]:
    setattr(torch.linalg, name, wrap("linalg." + name, getattr(torch.linalg, name)))
torch.cholesky_inverse = wrap("cholesky_inverse", torch.cholesky_inverse)
eng = layer_engine.LayerCalibrator(dev, n_iter=n_iter, keep_history=True)
try:
    wq, bq, a_w, a_act, out_q, rep = eng.run(x, wt, b, y, 1, (k - 1) // 2, lw, la, True, [att], name="dbg")
    print("ok: tc", rep.used_tc, "hist", rep.history, "final", rep.final_loss, "a_w", rep.alpha_w, "a_act", rep.alpha_act)
except Exception as e:  # noqa: BLE001
    print("raised:", repr(e))
print("calls", seen, "first non-finite:", bad[:3])
