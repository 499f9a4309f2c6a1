"""PCIe probe: pinned H2D / D2H bandwidth alone and concurrently (what bounds the e2e number)."""
import torch, time
n = 1 << 27   # 1 GiB of f64
h1 = torch.empty(n, dtype=torch.float64).pin_memory(); h2 = torch.empty(n // 2, dtype=torch.float64).pin_memory()
d1 = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.empty(n // 2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
d2h = t(lambda: h1.copy_(d1, non_blocking=True)); h2d = t(lambda: d2.copy_(h2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
b = t(both)
print(f"D2H 1 GiB: {d2h*1e3:.1f} ms = {n*8/d2h/1e9:.1f} GB/s; H2D 0.5 GiB: {h2d*1e3:.1f} ms = {n*4/h2d/1e9:.1f} GB/s; both concurrently: {b*1e3:.1f} ms")
