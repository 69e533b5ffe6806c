"""Probe: activation scale search on a level-1 sized tensor: plain passes vs interval-stable passes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from efficientq_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
n = int(os.environ.get("N", 268435456))
x = torch.relu(torch.randn(n, device=dev) * 1.1 + 0.2)
print({k: v for k, v in os.environ.items() if k.startswith("EFFQ_")}, "numel", n)
small = ops.workspace(ops.capi.load().effq_scale_search_workspace(0), dev)
for L in (16, 4):
    for tag, ws in (("interval", None), ("plain", small)):
        st = ops.ScaleState(dev)
        ops.scale_search(x, L, 0.0, 1.0, st, ws=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.scale_search(x, L, 0.0, 1.0, st, ws=ws)
        e1.record()
        torch.cuda.synchronize()
        s = st.read()
        d = ops.scale_search_diag(dev) if ws is None else {}
        ms = e0.elapsed_time(e1)
        print(f"L={L} {tag:9s} {ms:8.2f} ms  passes {s['passes']}  a {s['a']:.12f}  {d}  "
              f"plain-equivalent GB/s {4e-9 * n * (s['passes'] + 1) / (ms * 1e-3):.0f}")
