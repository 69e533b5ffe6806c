"""HBM throughput of the STE-backward kernel (effq_fakequant_ste_bwd) at a level-1 activation tensor of
BASELINE config 2 (32 x 32ch x 64^3 = 268 M elements), timed alone with CUDA events, inputs larger than L2."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from efficientq_b200 import ops
dev = torch.device("cuda:0")
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
n = 32 * 32 * 64 ** 3
x = torch.relu(torch.randn(n, device=dev)) * 1.5
g = torch.randn(n, device=dev)
alpha = torch.tensor([2.9], device=dev)
acc = torch.zeros(1, dtype=torch.float64, device=dev)
for want in (True, False):
    for _ in range(3):
        ops.fakequant_ste_bwd(x, g, alpha, 16, 0.0, 1.0, acc, want_grad_x=want)
    ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.fakequant_ste_bwd(x, g, alpha, 16, 0.0, 1.0, acc, want_grad_x=want); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    ms.sort(); t = ms[len(ms) // 2]
    b = n * (12 if want else 8)
    print(f"fakequant_ste_bwd grad_x={want}: {t:.3f} ms, {b / t / 1e6:.0f} GB/s = {100 * b / t / 1e6 / pk:.0f} % of the measured HBM peak ({pk} GB/s)")
