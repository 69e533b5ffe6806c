import os, sys, time
sys.path.insert(0, os.getcwd())
os.environ["EFFQ_LOOP_PROF"] = "1"
import torch, bench
from efficientq_b200 import ptqer, synth, capi
capi.load()
wl = bench.WORKLOADS["brats_w4a4_32x128"]
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
model, margs = bench.build_model(wl); model.to("cuda:0")
fp_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
data = synth.batch(32, 0, 4, wl["size"], "brats").to("cuda:0")
os.environ["EFFQ_LOOP_PROF"] = "0"
for i in range(2):
    model.load_state_dict(fp_state, strict=False)
    ptqer.calibrate(model, data, "brats", margs.init_stride); torch.cuda.synchronize()
os.environ["EFFQ_LOOP_PROF"] = "1"
model.load_state_dict(fp_state, strict=False)
t=time.time(); res = ptqer.calibrate(model, data, "brats", margs.init_stride); torch.cuda.synchronize(); print("step with loop prof", time.time()-t, res["t_fp"], res["t_ptq"])
# per-layer wall time without the profiler's syncs: events around each module's ptq
os.environ["EFFQ_LOOP_PROF"] = "0"
from efficientq_b200.qconv import EfficientQConv
evs = []
orig = EfficientQConv.ptq
def timed(self, x):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = orig(self, x); b.record(); evs.append((self.name, a, b)); return out
EfficientQConv.ptq = timed
model.load_state_dict(fp_state, strict=False)
res = ptqer.calibrate(model, data, "brats", margs.init_stride); torch.cuda.synchronize()
tot = 0
for n, a, b in evs:
    ms = a.elapsed_time(b); tot += ms
    print(f"[layer] {n:45s} {ms:8.2f} ms")
print("sum of layers", tot, "t_fp", res["t_fp"], "t_ptq", res["t_ptq"])
