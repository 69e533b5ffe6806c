"""Bring-up: LiTS-config calibration with per-layer input statistics (find the first layer that goes non-finite)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from efficientq_b200 import layer_engine, ops, ptqer, synth
wl = dict(bench.WORKLOADS["lits_w2a2_4x160"])
if len(sys.argv) > 1:
    wl["n"] = int(sys.argv[1])
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model, margs = bench.build_model(wl)
model.to(dev)
x = synth.batch(wl["n"], 0, 1, wl["size"], "lits").to(dev)
orig = layer_engine.LayerCalibrator.run
def run(self, x, weight, bias, out_fp, stride, padding, qlvl_w, qlvl_act, q_act, mask_pyramid=None, name=""):
    print(f"{name:45s} x {tuple(x.shape)} finite {bool(torch.isfinite(x).all())} max {float(x.max()):.4g} mean {float(x.mean()):.4g} "
          f"nonzero {float((x != 0).float().mean()):.3f} | y std {float(out_fp.std()):.4g} | L {qlvl_w}/{qlvl_act} q_act {q_act}", flush=True)
    try:
        res = orig(self, x, weight, bias, out_fp, stride, padding, qlvl_w, qlvl_act, q_act, mask_pyramid, name)
    except Exception as e:  # noqa: BLE001
        print("   FAILED:", repr(e), "| xstate", self.xstate.read(), "| wstate", self.wstate.read(), "| admm", self.st.read(), flush=True)
        raise
    rep = res[-1]
    print(f"   -> tc {rep.used_tc} final {rep.final_loss:.5g} best_it {rep.best_iter} a_w {rep.alpha_w:.5g} a_act {rep.alpha_act} passes {rep.act_passes} rho_scale {rep.rho_scale:.4g} fp64 {rep.fp64_factor}", flush=True)
    return res
layer_engine.LayerCalibrator.run = run

# non-finite tracer on every ops.* / library call of the layers whose name contains argv[2]
trace_name = sys.argv[2] if len(sys.argv) > 2 else None
tracing = [False]
def finite(v):
    if isinstance(v, torch.Tensor):
        return bool(torch.isfinite(v.float()).all()) if v.is_floating_point() and v.dtype != ops.E4M3 else True
    if isinstance(v, (tuple, list)):
        return all(finite(t) for t in v)
    return True
bad = []
def wrap(name, fn):
    def inner(*a, **kw):
        out = fn(*a, **kw)
        if tracing[0] and not bad:
            torch.cuda.synchronize()
            parts = {"out": out, "args": [t for t in a if isinstance(t, torch.Tensor)], **{k: v for k, v in kw.items() if isinstance(v, torch.Tensor)}}
            nf = [k for k, v in parts.items() if not finite(v)]
            if nf:
                bad.append(name)
                print(f"   NON-FINITE after {name}: {nf}", [(tuple(t.shape), finite(t)) for t in parts["args"]], flush=True)
                if name == "linalg.cholesky_ex":
                    print("   cholesky info", int(out[1].item()), "A diag min/max", float(a[0].diag().min()), float(a[0].diag().max()), "A finite", finite(a[0]), flush=True)
        return out
    return inner
if trace_name:
    for nm in ["fakequant_state", "quantize_act_ndhwc", "scale_search", "conv3d_f32", "conv3d_tc", "gram", "gram_tc",
               "admm_rhs", "split3_bf16", "solve_gemm_tc", "admm_lhs", "admm_project", "admm_track"]:
        setattr(ops, nm, wrap(nm, getattr(ops, nm)))
    for nm in ["cholesky_ex", "solve_triangular"]:
        setattr(torch.linalg, nm, wrap("linalg." + nm, getattr(torch.linalg, nm)))
    torch.cholesky_inverse = wrap("cholesky_inverse", torch.cholesky_inverse)
    traced = layer_engine.LayerCalibrator.run
    def run2(self, *a, name="", **kw):
        tracing[0] = trace_name in name
        return traced(self, *a, name=name, **kw)
    layer_engine.LayerCalibrator.run = run2
res = ptqer.calibrate(model, x, "lits", margs.init_stride)
print("done t_fp", res["t_fp"], "t_ptq", res["t_ptq"])
