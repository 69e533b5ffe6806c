"""Bring-up: first-iterate loss of the smoke layer under different kernel-path switches."""
import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientq_b200 import layer_engine, ops
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(7)
n, c1, c2, sp = 2, 32, 32, (8, 16, 16)
x = torch.relu(torch.randn(n, c1, *sp, generator=g))
w = torch.randn(c2, c1, 3, 3, 3, generator=g) * (2.0 / (c1 * 27)) ** 0.5
b = torch.randn(c2, generator=g) * 0.05
y = F.conv3d(x, w, b, 1, 1)
att = (torch.rand(n, *sp, generator=g) * 3).floor() + 1.0
dev = torch.device("cuda:0")
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
generic = tag == "generic"
eng = layer_engine.LayerCalibrator(dev, n_iter=int(os.environ.get("NITER", "20")), keep_history=True, force_generic=generic)
pyr = [att.to(dev)] if "pyr2" not in tag else [torch.ones(n, 3, 3, 3, device=dev), att.to(dev)]
wq, bq, a_w, a_act, out_q, rep = eng.run(x.to(dev), w.to(dev), b.to(dev), y.to(dev), 1, 1, 16, 16, True, pyr, name="smoke")
print(f"{tag:12s} env={ {k: v for k, v in os.environ.items() if k.startswith('EFFQ_')} } hist0 {rep.history[0]:.10f} final {rep.final_loss:.8f} "
      f"alpha_w {rep.alpha_w:.9f} alpha_act {rep.alpha_act:.9f} rho_scale {rep.rho_scale:.6f} (oracle hist0 0.0119172, final 0.0209342, alpha_act 2.923534583)")
