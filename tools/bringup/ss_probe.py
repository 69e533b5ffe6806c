"""Weight scale search timing: us per search and per pass for the tensor sizes of the BraTS net.
   python tools/bringup/ss_probe.py [reps]      (EFFQ_SS_BUCKET=0 selects the plain cluster / grid kernels)"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from efficientq_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
cases = [(96, 256), (3456, 256), (2048, 16), (8192, 16), (27648, 16), (32768, 16), (110592, 16), (442368, 16), (1769472, 16)]
st = ops.ScaleState(dev)
print(f"EFFQ_SS_BUCKET={os.environ.get('EFFQ_SS_BUCKET', '1')}")
for n, L in cases:
    g = torch.Generator().manual_seed(n)
    v = (torch.randn(n, generator=g) * 0.03).to(dev)
    d = (torch.randn(n, generator=g) * 0.003).to(dev)
    for _ in range(3):
        ops.scale_search(v, L, -1.0, 1.0, st, v2=d)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ops.scale_search(v, L, -1.0, 1.0, st, v2=d)
    b.record()
    torch.cuda.synchronize()
    s = st.read()
    us = a.elapsed_time(b) * 1e3 / reps
    print(f"numel {n:8d} L {L:3d}: {us:8.1f} us / search, {s['passes']:4d} passes, {us / max(s['passes'], 1):6.2f} us / pass, a {s['a']:.9f}")
if os.environ.get("EFFQ_SS_DEBUG") == "1":
    # clock stamps of the first passes (warp 0: publish, sums in, divisions done; warp 1: woke up, cells done, pushed)
    base = ops.capi.load().effq_scale_search_workspace(0)
    for n, L in [(96, 256), (3456, 256), (27648, 16)]:
        v = (torch.randn(n) * 0.03).to(dev)
        ops.scale_search(v, L, -1.0, 1.0, st)
        torch.cuda.synchronize()
        ws = ops._scale_ws(dev, n)
        tl = ws[base:base + 16 * 8 * 8].view(torch.int64).cpu().view(16, 8)
        t0 = int(tl[0, 0])
        print(f"timeline numel {n} L {L}: per pass [w0 publish | w1 awake | w1 cells done | w1 pushed | w0 sums in | w0 div done] (cycles from first publish)")
        for p in range(1, 8):
            r = tl[p]
            print("   pass", p, [int(r[0]) - t0, int(r[3]) - t0, int(r[4]) - t0, int(r[5]) - t0, int(r[1]) - t0, int(r[2]) - t0])
