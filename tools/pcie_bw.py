#!/usr/bin/env python
"""Host<->device copy bandwidth from pinned memory (context for the e2e number of bench.py)."""
import time, torch
n = 256 * 1024 * 1024
h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print("%s 1 GiB: %.1f ms  %.1f GB/s" % (name, dt * 1e3, n * 4 / dt / 1e9))
# both directions at once
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("both 1 GiB each: %.1f ms  %.1f GB/s per direction" % (dt * 1e3, n * 4 / dt / 1e9))
# many 6.7 MB pieces
m = 6_700_000 // 4
torch.cuda.synchronize(); t = time.perf_counter()
for k in range(150): d[k * m:(k + 1) * m].copy_(h[k * m:(k + 1) * m], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("H2D 150 x 6.7 MB: %.1f ms  %.1f GB/s" % (dt * 1e3, 150 * m * 4 / dt / 1e9))
