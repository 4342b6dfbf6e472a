"""Pinned host <-> device copy bandwidth of the box (what bounds the e2e figure of bench.py)."""
import torch, time
n = 78643200
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("%s: %.1f GB/s" % (name, 10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
# both directions at once on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("both directions at once: %.1f GB/s each" % (10 * n / dt / 1e9))
