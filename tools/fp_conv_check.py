"""FP targets of the trained miniature under the library conv and under the repo's own FP conv: per layer max
difference relative to max|target|, against an fp64 forward of the same layer input."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import torch.nn.functional as F
from tests.test_gpu_layer import build_toy
from efficientq_b200 import fold_bn, ptqer, synth, ops
from efficientq_b200.qconv import PTQConv
DEV = "cuda:0"
g = np.load("tests/golden/toy_dice.npz")
model, cfg = build_toy("brats")
model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
model.eval(); fold_bn.search_fold_and_remove_bn(model); model.to(DEV)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
data = synth.batch(cfg["n"], 0, cfg["num_mod"], (cfg["size"],) * 3, cfg["task"]).to(DEV)
ptqer.set_name(model); ptqer.set_fp(model)
rec = {}
def hook(tag):
    def h(m, i, o):
        x = i[0].detach()
        ref = F.conv3d(x.double(), m.weight.double(), m.bias.double() if m.bias is not None else None, m.stride, m.padding)
        rec.setdefault(m.name, {})[tag] = (o.detach().clone(), ref, tuple(x.shape), tuple(m.weight.shape), m.stride)
    return h
for tag in ("lib", "own"):
    os.environ["EFFQ_FP_CONV"] = tag
    hs = [m.register_forward_hook(hook(tag)) for _, m in model.named_modules() if isinstance(m, PTQConv)]
    with torch.no_grad():
        model(data)
    for h_ in hs: h_.remove()
for name, r in rec.items():
    ol, rl, xs, ws, st = r["lib"]; oo, ro, _, _, _ = r["own"]
    sc = float(rl.abs().max())
    sup = ops.conv3d_fp_supported(xs, ws[0], ws[2:], st, tuple((k - 1) // 2 for k in ws[2:]))
    print(f"{name:45s} x{xs} w{ws} s{st} tc={int(sup)} | lib-fp64 {float((ol.double()-rl).abs().max())/sc:.2e} own-fp64 {float((oo.double()-ro).abs().max())/sc:.2e} own-lib {float((oo-ol).abs().max())/sc:.2e} input diff {float((ro-rl).abs().max())/sc:.2e}")
