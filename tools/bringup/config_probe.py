"""Layer losses of a bench workload under different factorisation precisions (quality check of the fp32 explicit
inverse on ill-conditioned deep layers).  python tools/bringup/config_probe.py [workload]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from efficientq_b200 import ptqer, synth, layer_engine
name = sys.argv[1] if len(sys.argv) > 1 else "brats_w4a4_32x128"
wl = bench.WORKLOADS[name]
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
out = {}
for tag, f64 in (("fp32", False), ("fp64", True)):
    layer_engine.LayerCalibrator.force_fp64_factor = f64
    model, margs = bench.build_model(wl); model.to("cuda:0")
    data = synth.batch(wl["n"], 0, bench.N_MOD[wl["task"]], wl["size"], wl["task"]).to("cuda:0")
    res = ptqer.calibrate(model, data, wl["task"], margs.init_stride)
    out[tag] = np.array([float(ln.rsplit(":", 1)[1]) for ln in res["layer_loss"]])
    print(tag, "t_ptq", round(res["t_ptq"], 3))
    del model, data, res
    torch.cuda.empty_cache()
rel = (out["fp32"] - out["fp64"]) / out["fp64"]
print("fp32 vs fp64 factorisation, relative difference of the layer losses:")
print(" ".join(f"{r:+.3f}" for r in rel))
