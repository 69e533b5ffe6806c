"""BASELINE configs[0] on the UNMODIFIED reference, in full: BraTS-config 3D U-Net, W4A4, 8 synthetic 4x64^3 volumes,
200 ADMM iterations, CPU.  Writes tests/golden/config0.npz (per-layer losses and scales: the fixture
tests/test_gpu_config0.py compares the GPU run of the same job with) and prints the measured wall-clock next to what
bench.py's bounded-sample model predicts for the same job (profiles/r02_cpu_model_check.txt).
    python tools/ref_config0.py        (build container only: needs baseline/_ref, ~10-15 min on 8 cores)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from baseline import ref_harness as H  # noqa: E402
from efficientq_b200 import synth  # noqa: E402

wl = bench.WORKLOADS["brats_w4a4_8x64"]
assert H.ensure_ref()
R = H.import_reference()
torch.set_num_threads(os.cpu_count())
margs = bench.task_args(wl)
model = H.build_reference_model(R, margs, bench.seeded_state)
data = synth.batch(wl["n"], 0, 4, wl["size"], "brats")
t0 = time.perf_counter()
timers, t_fp, t_ptq, losses = H.timed_ptq(R, model, data, "brats", margs.init_stride, 200)
wall = time.perf_counter() - t0
names = [ln.rsplit(":", 1)[0].strip() for ln in losses]
vals = np.array([float(ln.rsplit(":", 1)[1]) for ln in losses])
res = {"layer_names": np.array(names), "layer_losses": vals, "wall_s": np.float64(wall), "t_fp": np.float64(t_fp),
       "t_ptq": np.float64(t_ptq), "threads": np.int64(torch.get_num_threads()),
       "phases": np.array(json.dumps({k: round(v, 3) for k, v in timers.items()}))}
for name, m in model.named_modules():
    if isinstance(m, R["ptqconv"].PTQConv):
        res[f"alpha_w::{name}"] = np.float32(m.alpha_w.item())
        res[f"alpha_act::{name}"] = np.float32(m.alpha_act.item())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "config0.npz"), **res)
print(f"unmodified reference, configs[0] in full: {wall:.1f} s on {torch.get_num_threads()} threads "
      f"(FP pass {t_fp:.2f} s, quantizing pass {t_ptq:.1f} s); phases {json.dumps({k: round(v, 1) for k, v in timers.items()})}")
v, desc, spent, full, kind = bench.cpu_sample(wl)
print(f"bounded-sample model ({kind}): {full:.0f} s predicted from a {spent:.0f} s sample -> ratio measured / predicted {wall / full:.2f}")
print(desc)
for nm, lv in zip(names, vals):
    print(f"{nm:45s} {lv:.6e}")
