"""Aggregate host->device bandwidth of N ranks copying from pinned memory at the same time (torchrun, one rank per GPU):
the ceiling of bench.py's e2e leg, which uploads 1.126 GB of feature maps per rank and step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bench.py [--no-numa]
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
where = "not bound (--no-numa)" if "--no-numa" in sys.argv else bench.bind_to_gpu_numa_node(torch, local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 1 << 30
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.fill_(rank)
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(2):
    dst.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dst.copy_(host, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = 10 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
t = torch.tensor([gbs], device=dev, dtype=torch.float64)
allg = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allg, t)
else:
    allg = [t]
print(f"rank {rank}: {gbs:6.1f} GB/s  ({where})", flush=True)
if world > 1:
    dist.barrier()
if rank == 0:
    vals = [float(x) for x in allg]
    print(f"N={world}: per rank min {min(vals):.1f} / max {max(vals):.1f} GB/s, aggregate {sum(vals):.1f} GB/s (all ranks copying concurrently)", flush=True)
if world > 1:
    dist.destroy_process_group()
