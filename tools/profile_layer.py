"""One level-1 layer's kernels at the BASELINE config-2 shape (32 volumes, 32 ch, 64^3 voxels),
each launched a few times: the target of the `ncu --set full` captures under profiles/.
Usage: python tools/profile_layer.py [n_volumes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c, sp = 32, (64, 64, 64)
torch.manual_seed(0)
x = torch.relu(torch.randn(n, c, *sp, device=dev))
y = torch.randn(n, c, *sp, device=dev)
att = (torch.rand(n, *sp, device=dev) * 3).floor() + 1
st = ops.ScaleState(dev)
ops.scale_search(x, 16, 0.0, 1.0, st)
print("act scale", st.read())
qx = ops.fakequant_state(x, st, 16, 0.0, 1.0)
alpha = st.a_f32().reshape(1)
yq, codes = ops.fakequant(x, alpha, 16, 0.0, 1.0, want_codes=True)
xq, xq8 = ops.quantize_act_ndhwc(x, 16, state=st, e4m3=True)
# glue ops between the layers (csrc/glue.cu)
r_ = ops.relu(x)
ad_ = ops.add(x, y)
mp_ = ops.maxpool3d(x, 2, relu_after=True)
up_ = ops.upsample_trilinear(mp_, 2, y)
del r_, ad_, mp_, up_
w = (2 * torch.randint(0, 16, (c, c, 3, 3, 3), device=dev) - 15).float()
wq = ops.pack_weight_codes(w)
wq8 = ops.pack_weight_codes(w, ops.CODE_E4M3)
cs = torch.full((1,), 1e-3, device=dev)
ws = ops.workspace(16 + 8 * 1024, dev)
sse = torch.zeros(1, dtype=torch.float64, device=dev)
for _ in range(4):
    ops.conv3d_tc(xq8, wq8, None, cs, c, 3, want_out=False, target=y, ws=ws, sse=sse)      # the step's conv (e4m3 codes)
for _ in range(2):
    ops.conv3d_tc(xq, wq, None, cs, c, 3, want_out=False, target=y, ws=ws, sse=sse)        # bf16 codes, for comparison
code_scale = (st.a_f32() / 15.0).reshape(1)
a0, b0, _, flag = ops.gram_tc(xq, code_scale, y, att, True, att_exact=True)
# round 2: weighted + unweighted Gram in one pass, rows-only pass, conv-free scoring, own SPD inverse
a0d, b0d, stats, gws, flagd = ops.gram_tc_dual(xq, code_scale, y, att, att_exact=True)
ops.gram_tc_rows_f64(xq, code_scale, y, stats, ws=gws)
yy = torch.ones(1, dtype=torch.float64, device=dev)
g0 = torch.randn(c, c * 27, device=dev) * 0.05
b_0 = torch.randn(c, device=dev) * 0.05
g1, b_1 = g0 + 1e-3 * torch.randn_like(g0), b_0 + 1e-3
for _ in range(3):
    ops.quadform_delta(stats, yy, g1, b_1, sse, g0, b_0)
from efficientq_b200.spd_inverse import SpdInverter  # noqa: E402
inv = SpdInverter(dev)
a_spd = a0d + 50.0 * torch.eye(a0d.shape[0], device=dev)
for _ in range(2):
    inv.invert(a_spd, want_inverse=False)
rows = ops.ScaleStateRows(dev, c)
sol = torch.randn(c, c * 27 + 1, device=dev) * 0.05
dual = torch.zeros(c, c * 27, device=dev)
wst = ops.ScaleState(dev)
for _ in range(3):
    ops.scale_search(sol[:, : c * 27], 16, -1.0, 1.0, wst, v2=dual)
ops.scale_search_rows(sol[:, : c * 27], 16, -1.0, 1.0, rows, v2=dual)
# proximal-step GEMM at the deepest layer's shape (C2 = 256, K' = 6913)
kp = 6913
bm = torch.randn(256, kp, device=dev)
ai = torch.randn(kp, kp, device=dev) * 1e-3
ap, bp = ops.split3_bf16(bm), ops.split3_bf16(ai)
for _ in range(3):
    ops.solve_gemm_tc(ap, bp, kp)
torch.cuda.synchronize()
print("done", sse.item(), int(flag.item()), wst.read()["passes"])
