"""HBM-bound kernels of the path, each timed alone on tensors far larger than L2 (CUDA events around 8 back-to-back
launches, best of 5 after 2 warm-ups), against the measured copy bandwidth in MEASURED_PEAKS.json:
    python tools/hbm_bench.py [n_volumes]          (EFFQ_QA_V3=0 / EFFQ_FQ_STATE_F64=1 select the older kernels)
fake-quant (values / values+codes), Qact from the scale-search state, NDHWC code kernel (bf16+e4m3) at the three
channel widths of the BraTS net, and the glue ops next to the library ops they replace."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficientq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
nvol = int(sys.argv[1]) if len(sys.argv) > 1 else 32
peak = 6549.1
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]


def best_ms(fn, reps=5, warm=2, inner=8):
    """Best of `reps` brackets of `inner` back-to-back launches each (per-launch time): with ONE launch per bracket the
    host-side cost of the call (ctypes + wrapper, ~30 us against ~5 us for a library op) sits between the start event
    and the kernel and is charged to kernels that only run for 0.3 ms."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / inner)
    return best


rows = []


def line(name, nbytes, fn):
    ms = best_ms(fn)
    gbs = nbytes / ms / 1e6
    rows.append((name, nbytes / 1e6, ms, gbs, gbs / peak))
    print(f"| {name} | {nbytes / 1e6:.0f} | {ms:.3f} | {gbs:.0f} | {gbs / peak:.2f} |", flush=True)


print(f"HBM copy peak {peak:.0f} GB/s; EFFQ_QA_V3={os.environ.get('EFFQ_QA_V3', '1')} "
      f"EFFQ_FQ_STATE_F64={os.environ.get('EFFQ_FQ_STATE_F64', '0')}; {nvol} volumes")
print("| kernel | algorithmic MB | ms | GB/s | of the copy peak |")
print("|---|---:|---:|---:|---:|")
torch.manual_seed(0)
for c, sp in ((32, 64), (64, 32), (128, 16)):
    x = torch.relu(torch.randn(nvol, c, sp, sp, sp, device=dev))
    n = x.numel()
    st = ops.ScaleState(dev)
    st.set_a(float(x.mean().item()) * 2.5)
    alpha = st.a_f32().reshape(1)
    if c == 32:
        line("fakequant_f32 (values)", 8 * n, lambda: ops.fakequant(x, alpha, 16, 0.0, 1.0))
        line("fakequant_f32 (values + codes)", 9 * n, lambda: ops.fakequant(x, alpha, 16, 0.0, 1.0, want_codes=True))
        line("fakequant_state (Qact)", 8 * n, lambda: ops.fakequant_state(x, st, 16, 0.0, 1.0))
    line(f"quantize_act_ndhwc C={c} (bf16 + e4m3, fp64 scale)", 7 * n,
         lambda: ops.quantize_act_ndhwc(x, 16, state=st, e4m3=True))
    line(f"quantize_act_ndhwc C={c} (e4m3 only, fp32 scale)", 5 * n,
         lambda: ops.quantize_act_ndhwc(x, 16, alpha=alpha, bf16=False, e4m3=True))
    if c == 32:
        y = torch.randn_like(x)
        line("glue relu (own)", 8 * n, lambda: ops.relu(x))
        line("relu (library)", 8 * n, lambda: F.relu(x))
        line("glue add (own)", 12 * n, lambda: ops.add(x, y))
        line("add (library)", 12 * n, lambda: x + y)
        line("glue maxpool2 + relu (own, one pass)", 4 * n + n // 2, lambda: ops.maxpool3d(x, 2, relu_after=True))
        line("max_pool3d then relu_ (library, two passes)", 4 * n + n // 2, lambda: F.relu_(F.max_pool3d(x, 2, 2)))
        del y
        xs = torch.randn(nvol, c, sp // 2, sp // 2, sp // 2, device=dev)
        skip = torch.randn(nvol, c, sp, sp, sp, device=dev)
        line("glue upsample x2 + skip (own, one pass)", 8 * n + n // 2, lambda: ops.upsample_trilinear(xs, 2, skip))
        line("interpolate then add (library, two passes)", 8 * n + n // 2,
             lambda: F.interpolate(xs, scale_factor=2, mode="trilinear") + skip)
        del xs, skip
    del x
    torch.cuda.empty_cache()
