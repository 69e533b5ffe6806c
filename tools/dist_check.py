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
    # the same job UNSHARDED on this GPU: the first two layers must agree to 1e-6 (from the third on the fp32 partial
    # sums of A0, added in a different order, have flipped a code somewhere and the trajectories part ways like
    # any two correct implementations do), and all layers must lie in the reference's own ensemble range
    # (tests/golden/toy_net.npz::ensemble_losses) widened by one range on either side
    model1, _ = build_toy()
    model1.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    model1.eval()
    fold_bn.search_fold_and_remove_bn(model1)
    model1.to(dev)
    full = synth.batch(cfg["n"], 0, cfg["num_mod"], (cfg["size"],) * 3, cfg["task"]).to(dev)
    from efficientq_b200.dist import DistCtx
    res1 = ptqer.calibrate(model1, full, "brats", "2,2,2", DistCtx(single=True))     # no collectives: rank 1 is not here
    one = np.array([float(ln.rsplit(":", 1)[1]) for ln in res1["layer_loss"]])
    ens = np.concatenate([g["ensemble_losses"], ref[None]], 0)
    lo, hi = ens.min(0), ens.max(0)
    for nm, a, b, c in zip(g["layer_names"], losses, one, ref):
        print(f"{str(nm):45s} sharded {a:.6e} unsharded {b:.6e} (rel {abs(a - b) / b:.1e}) reference {c:.6e}")
    assert np.allclose(losses[:2], one[:2], rtol=1e-6), (losses[:2], one[:2])
    assert abs(losses[0] - ref[0]) <= 1e-3 * ref[0]
    width = hi - lo
    assert ((losses >= lo - width - 1e-3 * ref) & (losses <= hi + width + 1e-3 * ref)).all()
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

# a15: sharded alpha_act refinement against the reference's UNSHARDED tune_activation_range fixture.
# Every rank differentiates its own volume; the alpha gradients are all-reduced (tune.py).
from efficientq_b200 import ops, tune  # noqa: E402
from efficientq_b200.qconv import PTQConv  # noqa: E402
gt = np.load(os.path.join(ROOT, "tests", "golden", "toy_tune.npz"))
model2, _ = build_toy()
model2.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
model2.eval()
fold_bn.search_fold_and_remove_bn(model2)
model2.to(dev)
ptqer.set_name(model2)
ptqer.set_fp(model2)
with torch.no_grad():
    out_fp = model2(data).detach()
mods = {n: m for n, m in model2.named_modules() if isinstance(m, PTQConv)}
for name, m in mods.items():
    m.weight.data = torch.from_numpy(gt[f"pre::{name}.weight"]).to(dev)
    m.bias.data = torch.from_numpy(gt[f"pre::{name}.bias"]).to(dev)
    m.alpha_w.data = torch.tensor(float(gt[f"pre::{name}.alpha_w"]), device=dev)
    m.alpha_act.data = torch.tensor(float(gt[f"pre::{name}.alpha_act"]), device=dev)
losses_t = tune.tune_activation_range(model2, out_fp, data, max_iter=3, dist=dist)
post = {n: float(m.alpha_act.detach()) for n, m in mods.items() if m.q_act}
if dist.rank == 0:
    print("tune losses sharded", losses_t, "ref (unsharded)", gt["tune_losses"].tolist())
    # same bars as the unsharded test (tests/test_gpu_tune.py): 1e-5 on step 0; after the first Adam step a few codes sit
    # next to a rounding boundary, where the reference's fp32 MKL conv and the tensor-core conv decide differently
    # (2.5e-4 on the loss of step 1 with this fixture, 1e-7 on steps 0 and 2)
    assert np.allclose(losses_t[0], gt["tune_losses"][0], rtol=1e-5)
    assert np.allclose(losses_t, gt["tune_losses"], rtol=5e-4)
    for n, v in post.items():
        r = float(gt[f"post::{n}.alpha_act"])
        assert abs(v - r) <= 2e-5 * r, (n, v, r)
    print(f"TUNE DIST OK world={dist.world}: refined alpha_act equal the unsharded reference's")
torch.distributed.destroy_process_group() if dist.world > 1 else None
