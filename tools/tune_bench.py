"""Time of one Adam iteration of the end-to-end alpha_act refinement (row a15) on the BraTS config, 128^3 volumes:
tensor-core dgrad (tcgen05 conv on bf16 split planes) vs the library's fp32 dgrad.  Usage: python tools/tune_bench.py [N]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from efficientq_b200 import ops, ptqer, synth, tune
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = dict(bench.WORKLOADS["brats_w4a4_32x128"]); wl["n"] = n
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
model, margs = bench.build_model(wl); model.to(dev)
x = synth.batch(n, 0, 4, wl["size"], "brats").to(dev)
res = ptqer.calibrate(model, x, "brats", margs.init_stride, n_iter=10)      # short calibration: weights on their grids
out_fp = res["output_fp"]
del res
for flag, name in (("1", "tcgen05 dgrad"), ("0", "library fp32 dgrad")):
    os.environ["EFFQ_DGRAD_TC"] = flag
    tune.tune_activation_range(model, out_fp, x, max_iter=1)                # warm-up
    torch.cuda.synchronize(); t = time.time()
    losses = tune.tune_activation_range(model, out_fp, x, max_iter=3)
    torch.cuda.synchronize()
    print(f"{name:22s}: {(time.time() - t) / 3 * 1e3:8.1f} ms per Adam iteration ({n} volumes of 128^3), losses {losses}")
