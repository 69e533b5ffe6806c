"""Probe: time the per-iteration solve GEMM  sol = B @ Ainv  (fp32) for the BraTS layer shapes."""
import os, sys, torch
dev = "cuda:0"
print("cublas env:", {k: v for k, v in os.environ.items() if "CUBLAS" in k}, torch.version.cuda)
torch.backends.cuda.matmul.allow_tf32 = False
for c2, kp in [(32, 865), (64, 1729), (128, 3457), (256, 6913)]:
    b = torch.randn(c2, kp, device=dev)
    a = torch.randn(kp, kp, device=dev)
    a = a + a.T
    ref = (b.double() @ a.double())
    for _ in range(3):
        out = b @ a
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = b @ a
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"C2={c2} K'={kp}: {ms*1e3:.1f} us  {2*c2*kp*kp/ms/1e9:.1f} TFLOP/s  max rel err {err:.2e}")
