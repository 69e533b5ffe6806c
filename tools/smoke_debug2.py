"""Bring-up: iterate-0 intermediates of the smoke layer against oracle values (tools/_dbg/smoke_it0.npz)."""
import os, sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientq_b200 import layer_engine, ops
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_dbg", "smoke_it0.npz"))
dev = torch.device("cuda:0")
T = lambda k: torch.from_numpy(d[k]).to(dev)
x, w, b, y, att = T("x"), T("w"), T("b"), T("y"), T("att")
def rel(a, r): return float((a.double() - r.double()).abs().max() / r.double().abs().max())
for generic in (False, True):
    eng = layer_engine.LayerCalibrator(dev, n_iter=1, keep_history=True, force_generic=generic)
    wq, bq, a_w, a_act, out_q, rep = eng.run(x, w, b, y, 1, 1, 16, 16, True, [att], name="smoke")
    G, bo = T("G"), T("bq")
    print("generic" if generic else "tc", "hist0", rep.history[0], "oracle", float(d["h0"]), "final", rep.final_loss, "oracle", float(d["final"]))
    print("  a_w", float(a_w), float(d["a_w"]), "a_act", rep.alpha_act, float(d["a_act"]))
    print("  G rel", rel(wq, G), "n codes differ", int(((wq - G).abs() > 1e-4).sum()), "of", G.numel(), " bias rel", rel(bq, bo), bq[:4].tolist(), bo[:4].tolist())
    qact = T("qact")
    o_ref = F.conv3d(qact.double(), G.double(), bo.double(), 1, 1)
    o_gpuw = F.conv3d(qact.double(), wq.double(), bq.double(), 1, 1)
    print("  out_q vs fp64 conv(gpu w)", rel(out_q, o_gpuw), " vs conv(oracle w)", rel(out_q, o_ref))
    print("  mse fp64: gpu w", float(((o_gpuw - y) ** 2).mean()), "oracle w", float(((o_ref - y) ** 2).mean()),
          "att-weighted gpu w", float((att.unsqueeze(1) * (o_gpuw - y) ** 2).mean()), "oracle w", float((att.unsqueeze(1) * (o_ref - y) ** 2).mean()))
    # normal equations and the first proximal step, piece by piece
    qx = ops.fakequant_state(x, eng.xstate, 16, 0.0, 1.0)
    print("  qact rel", rel(qx, qact), "n differ", int(((qx - qact).abs() > 1e-6).sum()))
    a0, b0 = ops.gram(qx, y, att, (3, 3, 3), 1, 1, has_bias=True)
    print("  gram_f32 A0 rel", rel(a0, T("a0")), "B0 rel", rel(b0, T("b0")))
    rs = rep.rho_scale
    qe = torch.eye(865, dtype=torch.float64, device=dev); qe[-1, -1] = 0
    A = T("a0").double() + 10 * rs * qe + rs * torch.eye(865, dtype=torch.float64, device=dev)
    w0p = torch.cat([w.reshape(32, -1), b.reshape(32, 1)], 1).double()
    B = T("b0").double() + rs * w0p; B[:, :864] += 10 * rs * w.reshape(32, -1).double()
    sol = torch.linalg.solve(A, B.T).T
    print("  fp64 solve on GPU vs oracle wstar", rel(sol[:, :864].reshape(w.shape), T("wstar")), "bstar", rel(sol[:, 864], T("bstar")))
