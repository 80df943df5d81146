"""Aggregate D2H bandwidth of N GPUs copying into pinned host memory at the same time (plain cudaMemcpyAsync, no
kernels): the platform ceiling the e2e host path runs into.  torchrun --nproc-per-node N scripts/host_d2h_ceiling.py"""
import json, os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
src = torch.empty(64 << 20, dtype=torch.uint8, device=dev)          # 64 MiB per copy
dst = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
up = torch.empty(4 << 20, dtype=torch.uint8, pin_memory=True); upd = torch.empty(4 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < 1.0:
  for _ in range(8): dst.copy_(src, non_blocking=True)
  torch.cuda.synchronize(); n += 8
dt = time.perf_counter() - t0
gbs = torch.tensor([n * (64 << 20) / dt / 1e9], dtype=torch.float64, device=dev)
alls = [torch.zeros_like(gbs) for _ in range(world)]
if world > 1: dist.all_gather(alls, gbs)
else: alls = [gbs]
if rank == 0:
  v = [round(float(x.item()), 1) for x in alls]
  print(json.dumps({"n_gpus": world, "d2h_gbs_per_gpu": v, "aggregate_gbs": round(sum(v), 1)}))
if world > 1: dist.destroy_process_group()
