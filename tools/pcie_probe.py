"""Raw PCIe bandwidth of the box (pinned memory), to put the e2e figure in context."""
import torch, time
n = 1 << 27  # 1 GiB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory(); h.fill_(1.0)
h2 = torch.empty(n // 2, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device="cuda"); d2 = torch.ones(n // 2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True))
b = t(lambda: h2.copy_(d2, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print("H2D 1 GiB alone: %.1f ms (%.1f GB/s) | D2H 0.5 GiB alone: %.1f ms (%.1f GB/s) | both concurrently: %.1f ms (H2D-equivalent %.1f GB/s)"
      % (a * 1e3, 2**30 / a / 1e9, b * 1e3, 2**29 / b / 1e9, c * 1e3, 2**30 / c / 1e9))
