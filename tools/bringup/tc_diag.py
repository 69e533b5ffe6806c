"""Bring-up diagnostics for the tcgen05 conv kernel: structured inputs, simplest case first.
Usage (GPU box): python tools/bringup/tc_diag.py > gpurun_out/tc_diag.log 2>&1"""
import sys
import os
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from efficientq_b200 import ops  # noqa: E402

DEV = "cuda:0"


def run(tag, xc, wc, k, scale=1.0, bias=None):
    n, c1 = xc.shape[:2]
    c2 = wc.shape[0]
    want = F.conv3d(xc.double(), wc.double(), None, 1, (k - 1) // 2).float() * scale
    if bias is not None:
        want = want + bias.view(1, -1, 1, 1, 1)
    xq = xc.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16).to(DEV)
    wq = ops.pack_weight_codes(wc.to(DEV))
    cs = torch.tensor([scale], dtype=torch.float32, device=DEV)
    ws = ops.workspace(16 + 8 * 1024, torch.device(DEV))
    t0 = time.time()
    try:
        out, sse = ops.conv3d_tc(xq, wq, bias.to(DEV) if bias is not None else None, cs, c2, k, want_out=True,
                                 target=want.to(DEV), ws=ws)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[{tag}] EXCEPTION {e!r}")
        return False
    dt = time.time() - t0
    flags = ws[:8].view(torch.int32).cpu().tolist()
    diff = (out.cpu() - want).abs()
    ok = diff.max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    print(f"[{tag}] ok={ok} maxdiff={diff.max().item():.4g} wantmax={want.abs().max().item():.4g} "
          f"sse={sse.item():.4g} ws(done,abort)={flags} {dt * 1e3:.1f} ms")
    if not ok:
        idx = torch.nonzero(diff > 1e-5 * max(1.0, want.abs().max().item()))
        print("   mismatches:", idx.shape[0], "of", diff.numel(), "first:", idx[:6].tolist())
        for i in idx[:6].tolist():
            print("     at", i, "got", out.cpu()[tuple(i)].item(), "want", want[tuple(i)].item())
        ch = (diff > 0).any(dim=0).any(dim=-1).any(dim=-1).any(dim=-1)
        print("   bad channels:", torch.nonzero(ch).flatten().tolist()[:40])
    return ok


def main():
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0), {k: v for k, v in os.environ.items() if k.startswith("EFFQ_")})
    ok = True
    for c1, c2 in [(32, 32), (64, 64), (16, 16)]:
        sp = (2, 16, 8)
        ones_x = torch.ones(1, c1, *sp)
        ones_w = torch.ones(c2, c1, 1, 1, 1)
        ok &= run(f"k1 ones c{c1}->{c2}", ones_x, ones_w, 1)
        x = torch.randint(0, 16, (1, c1, *sp)).float()
        w1 = torch.zeros(c2, c1, 1, 1, 1)
        for j in range(c2):
            w1[j, j % c1] = 1.0
        ok &= run(f"k1 perm c{c1}->{c2}", x, w1, 1)
        w = (2 * torch.randint(0, 16, (c2, c1, 1, 1, 1)) - 15).float()
        ok &= run(f"k1 rand c{c1}->{c2}", x, w, 1, 0.01, torch.randn(c2))
    for c1, c2 in [(32, 32), (64, 64)]:
        sp = (3, 16, 8)
        x = torch.randint(0, 16, (1, c1, *sp)).float()
        for tap in [(1, 1, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0), (2, 2, 2)]:
            w = torch.zeros(c2, c1, 3, 3, 3)
            for j in range(c2):
                w[j, j % c1, tap[0], tap[1], tap[2]] = 1.0
            ok &= run(f"k3 tap{tap} c{c1}->{c2}", x, w, 3)
        w = (2 * torch.randint(0, 16, (c2, c1, 3, 3, 3)) - 15).float()
        ok &= run(f"k3 rand c{c1}->{c2}", x, w, 3, 0.01, torch.randn(c2))
    x = torch.randint(0, 16, (2, 128, 4, 20, 12)).float()
    w = (2 * torch.randint(0, 16, (128, 128, 3, 3, 3)) - 15).float()
    ok &= run("k3 rand c128->128 ragged", x, w, 3, 0.01, torch.randn(128))
    x = torch.randint(0, 16, (1, 256, 4, 8, 8)).float()
    w = (2 * torch.randint(0, 16, (256, 256, 3, 3, 3)) - 15).float()
    ok &= run("k3 rand c256->256", x, w, 3, 0.01, torch.randn(256))
    print("ALL OK" if ok else "SOME FAILED")
    # throughput of the scoring call (no output written) at level-1 / level-2 shapes
    fp8 = os.environ.get("EFFQ_DIAG_FP8", "0") == "1"
    for c, sp, n in [(32, (64, 64, 64), 8), (64, (32, 32, 32), 32), (128, (16, 16, 16), 32)]:
        xq = torch.randint(0, 16, (n, *sp, c), device=DEV).to(torch.float8_e4m3fn if fp8 else torch.bfloat16)
        w = (2 * torch.randint(0, 16, (c, c, 3, 3, 3), device=DEV) - 15).float()
        wq = ops.pack_weight_codes(w, ops.CODE_E4M3 if fp8 else ops.CODE_BF16)
        tgt = torch.randn(n, c, *sp, device=DEV)
        cs = torch.ones(1, device=DEV)
        ws = ops.workspace(16 + 8 * 1024 + 16 * 1024, torch.device(DEV))
        sse = torch.zeros(1, dtype=torch.float64, device=DEV)
        for _ in range(3):
            ops.conv3d_tc(xq, wq, None, cs, c, 3, want_out=False, target=tgt, ws=ws, sse=sse)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.conv3d_tc(xq, wq, None, cs, c, 3, want_out=False, target=tgt, ws=ws, sse=sse)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * n * sp[0] * sp[1] * sp[2] * c * c * 27
        print(f"[perf {'e4m3' if fp8 else 'bf16'} c{c} {n}x{sp}] {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
        if int(os.environ.get("EFFQ_TC_DEBUG", "0")) & 8:
            tl = ws[16 + 8 * 1024:16 + 8 * 1024 + 16 * 1024].view(torch.int64).cpu().reshape(-1, 8)
            t0 = int(tl[2, 0])
            names = ["mma:acc_free", "mma:halo_ready", "mma:issued", "epi:acc_full", "epi:done", "prod:slot_free",
                     "prod:copies_issued", "prod:signalled"]
            print("   timeline (cycles rel. to tile 2, CTA 0):", names)
            for i in range(2, 12):
                print("   tile", i, [int(v) - t0 for v in tl[i]])


if __name__ == "__main__":
    main()
