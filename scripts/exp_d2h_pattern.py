"""How much of the raw D2H rate do different copy patterns of one host-path step keep?  (experiment, round 2)"""
import time, torch
dev = torch.device('cuda', 0)
N = 65536
def bufs(n, sizes):
  return [(torch.empty(s * n, dtype=torch.uint8, device=dev), torch.empty(s * n, dtype=torch.uint8, pin_memory=True)) for s in sizes]
def run(groups, sizes, iters=300, label=''):
  n = N // groups
  gb = [bufs(n, sizes) for _ in range(groups)]
  streams = [torch.cuda.Stream(dev) for _ in range(groups)]
  evs = [torch.cuda.Event() for _ in range(groups)]
  def submit(g):
    with torch.cuda.stream(streams[g]):
      for d, h in gb[g]: h.copy_(d, non_blocking=True)
      evs[g].record(streams[g])
  for g in range(groups): submit(g)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for g in range(groups): submit(g)
  for i in range(iters):
    for g in range(groups):
      evs[g].synchronize(); submit(g)
  torch.cuda.synchronize()
  dt = time.perf_counter() - t0
  total = sum(sizes) * N * (iters + 1)
  print(f"{label:48s} groups={groups}: {total / dt / 1e9:6.1f} GB/s  {dt / (iters + 1) * 1e6:7.1f} us/step")
for g in (1, 2, 3, 4):
  run(g, [256], label='obs only (256 B/env)')
  run(g, [256, 8, 1, 1], label='obs + reward + terminated + truncated (4 copies)')
  run(g, [256, 10], label='obs + one packed 10 B/env record (2 copies)')
  run(g, [266], label='everything in one copy')
