"""BASELINE.json configs[4]: fake-quant 3D-conv micro-sweep.

channels 16-256, levels 4/16/256 (2/4/8 bits), volumes 32^3 - 128^3, k=3 s=1 p=1, x = relu(randn) fp32,
w = randn*(2/(27C))^0.5.  For every point, timed alone with CUDA events on
the launching stream after warm-up, with a 256 MB write between timed launches (L2 flush):

  fq      effq_fakequant_f32            fp32 -> fp32 fake-quant values         numel*8 B        vs HBM peak
  codes   effq_quantize_act_ndhwc       fp32 NCDHW -> NDHWC codes (operand)    numel*(4+b) B    vs HBM peak
  conv    effq_conv3d_tc (+fused SSE)   codes x codes conv + sum (out-y)^2     2*V*C*C*27 flop  vs bf16 tensor peak
                                                                               and bytes vs HBM peak
The batch N is chosen so the fp32 activation tensor is ~256 MB (larger than the 126 MB L2) but at least 1.
Writes a markdown table to stdout (and gpurun_out/conv_sweep.md when that directory exists).
Usage: python tools/conv_sweep.py [--quick]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficientq_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
quick = "--quick" in sys.argv
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TF = pk["hbm_gbs"], pk["bf16_tflops"]       # kernels timed alone -> burst figures
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(reps):
        flush_buf.fill_(1)                                  # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


rows = []
chans = [16, 32, 64, 128, 256]
levels = [4, 16, 256]
edges = [32, 64, 128]
for c in chans:
    for edge in edges:
        vox = edge ** 3
        n = max(1, min(64, (256 << 20) // (4 * c * vox)))
        if quick and edge != 64:
            continue
        torch.manual_seed(c * 1000 + edge)
        x = torch.relu(torch.randn(n, c, edge, edge, edge, device=dev))
        y = torch.randn(n, c, edge, edge, edge, device=dev)
        numel = x.numel()
        for lv in levels:
            st = ops.ScaleState(dev)
            ops.scale_search(x, lv, 0.0, 1.0, st)
            alpha = st.a_f32().reshape(1)
            t_fq = timed(lambda: ops.fakequant(x, alpha, lv, 0.0, 1.0))
            fp8 = lv <= 16 and ops.conv3d_tc_supported(x.shape, c, 3, 1, 1, ops.CODE_E4M3)
            tc_ok = ops.conv3d_tc_supported(x.shape, c, 3, 1, 1)
            t_codes = timed(lambda: ops.quantize_act_ndhwc(x, lv, state=st, bf16=not fp8, e4m3=fp8))
            cb = 1 if fp8 else 2
            row = dict(c=c, edge=edge, n=n, lv=lv, kind="e4m3" if fp8 else "bf16",
                       fq_ms=t_fq, fq_gbs=numel * 8 / t_fq / 1e6, codes_ms=t_codes,
                       codes_gbs=numel * (4 + cb) / t_codes / 1e6)
            if tc_ok:
                res = ops.quantize_act_ndhwc(x, lv, state=st, bf16=not fp8, e4m3=fp8)
                xc = res[1] if fp8 else res
                wint = (2 * torch.randint(0, lv, (c, c, 3, 3, 3), device=dev) - (lv - 1)).float()
                wc = ops.pack_weight_codes(wint, ops.CODE_E4M3 if fp8 else ops.CODE_BF16)
                cs = torch.full((1,), 1e-3, device=dev)
                ws = ops.workspace(16 + 8 * 1024, dev)
                sse = torch.zeros(1, dtype=torch.float64, device=dev)
                t_conv = timed(lambda: ops.conv3d_tc(xc, wc, None, cs, c, 3, want_out=False, target=y, ws=ws, sse=sse))
                if sse.item() != sse.item():
                    raise SystemExit(f"conv aborted at c={c} edge={edge} lv={lv}")
                flops = 2.0 * n * vox * c * c * 27
                nbytes = numel * cb + numel * 4 + wc.numel() * cb
                row.update(conv_ms=t_conv, conv_tf=flops / t_conv / 1e9, conv_gbs=nbytes / t_conv / 1e6,
                           mvox_s=n * vox / t_conv / 1e3)
            rows.append(row)
            print(row, file=sys.stderr, flush=True)
        del x, y
        torch.cuda.empty_cache()

out = ["| C | volume | N | levels | operand | fake-quant GB/s (% HBM) | NDHWC codes GB/s (% HBM) | conv+SSE ms | Mvoxel/s | TFLOP/s (% bf16 peak) | conv GB/s (% HBM) |",
       "|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    conv = (f"{r['conv_ms']:.3f} | {r['mvox_s']:.0f} | {r['conv_tf']:.0f} ({100 * r['conv_tf'] / TF:.0f} %) | "
            f"{r['conv_gbs']:.0f} ({100 * r['conv_gbs'] / HBM:.0f} %)") if "conv_ms" in r else "generic fp32 path | | |"
    out.append(f"| {r['c']} | {r['edge']}^3 | {r['n']} | {r['lv']} | {r['kind']} | {r['fq_gbs']:.0f} ({100 * r['fq_gbs'] / HBM:.0f} %) | "
               f"{r['codes_gbs']:.0f} ({100 * r['codes_gbs'] / HBM:.0f} %) | {conv} |")
text = f"peaks: HBM {HBM} GB/s, bf16 {TF} TFLOP/s (MEASURED_PEAKS.json, burst)\n\n" + "\n".join(out) + "\n"
print(text)
if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
    open(os.path.join(ROOT, "gpurun_out", "conv_sweep.md"), "w").write(text)
