"""Pinned host <-> device copy bandwidth of this box (the ceiling of bench.py's e2e)."""
import torch, time
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, a, b in (('D2H', h, d), ('H2D', d, h)):
  for _ in range(2):
    a.copy_(b, non_blocking=True)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(5):
    a.copy_(b, non_blocking=True)
  torch.cuda.synchronize()
  dt = (time.perf_counter() - t0) / 5
  print('%s pinned, 4 GiB: %.1f GB/s' % (name, n / dt / 1e9))
# two copies in flight on two streams (what the engine does with file 1 / file 2)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
  with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
  with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print('D2H two streams: %.1f GB/s' % (6 * n / (time.perf_counter() - t0) / 1e9))
