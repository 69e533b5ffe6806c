"""Sharded (torchrun, NCCL) calibration of the toy U-Net vs the reference fixture computed
UNSHARDED on the same 2 volumes.  torchrun --nproc-per-node 2 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficientq_b200 import fold_bn, ptqer, synth  # noqa: E402
from efficientq_b200.dist import init_from_env  # noqa: E402
from tests.test_gpu_layer import build_toy  # noqa: E402

dist = init_from_env("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
g = np.load(os.path.join(ROOT, "tests", "golden", "toy_net.npz"))
model, cfg = build_toy()
model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
model.eval()
fold_bn.search_fold_and_remove_bn(model)
model.to(dev)
lo, hi = dist.shard(cfg["n"])
data = synth.batch(hi - lo, lo, cfg["num_mod"], (cfg["size"],) * 3, cfg["task"]).to(dev)
res = ptqer.calibrate(model, data, "brats", "2,2,2", dist)
losses = np.array([float(ln.rsplit(":", 1)[1]) for ln in res["layer_loss"]])
ref = g["layer_losses"]
if dist.rank == 0:
    assert res["class_nums"] == [int(v) for v in g["class_nums"]], (res["class_nums"], g["class_nums"])
    for nm, a, b in zip(g["layer_names"], losses, ref):
        print(f"{str(nm):45s} sharded {a:.6e} ref {b:.6e} rel {abs(a - b) / b:.2e}")
    assert abs(losses[0] - ref[0]) <= 1e-3 * ref[0]
    assert np.allclose(losses, ref, rtol=5e-2)
    print(f"DIST OK world={dist.world} t_fp {res['t_fp']:.2f}s t_ptq {res['t_ptq']:.2f}s")
# every rank must hold identical quantised weights (replicated solve / projection)
w = torch.cat([p.detach().flatten() for p in model.parameters()])
wsum = w.double().sum().reshape(1)
mx = wsum.clone()
dist.all_reduce_max(mx)
mn = -wsum.clone()
dist.all_reduce_max(mn)
assert mx.item() == -mn.item(), "ranks diverged"
if dist.rank == 0:
    print("ranks hold identical weights")
torch.distributed.destroy_process_group() if dist.world > 1 else None
