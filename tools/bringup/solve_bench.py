"""Time the library linear algebra of the proximal step at the BraTS layer sizes."""
import torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for kp, c2 in [(865, 32), (1729, 64), (3457, 128), (6913, 256)]:
    x = torch.randn(kp, kp + 64, device=dev)
    a = x @ x.T + kp * torch.eye(kp, device=dev)
    b = torch.randn(c2, kp, device=dev)
    chol = torch.linalg.cholesky(a)
    ainv = torch.cholesky_inverse(chol)
    print(f"K'={kp:5d} C2={c2:3d}: cholesky_ex {t(lambda: torch.linalg.cholesky_ex(a)):8.3f} ms  "
          f"cholesky_inverse {t(lambda: torch.cholesky_inverse(chol)):8.3f} ms  inv {t(lambda: torch.linalg.inv(a), 2):8.3f} ms  "
          f"B@Ainv {t(lambda: b @ ainv, 20):7.3f} ms  cholesky_solve {t(lambda: torch.cholesky_solve(b.T.contiguous(), chol), 5):7.3f} ms")
