"""Raw pinned H2D / D2H bandwidth of the box (context for the e2e number)."""
import torch, time
n = 64 * 1024 * 1024
h = torch.empty(n // 4, dtype=torch.float32).pin_memory(); d = torch.empty(n // 4, dtype=torch.float32, device="cuda")
for name, f in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print(name, "%.1f GB/s" % (n / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n // 4, dtype=torch.float32).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("duplex %.1f GB/s each way" % (n / dt / 1e9))
