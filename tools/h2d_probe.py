#!/usr/bin/env python
"""Pinned host->device copy rates per rank in the patterns bench.py's e2e leg uses (torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N tools/h2d_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def report(tag, gbs):
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        vals = [float(x) for x in out]
    else:
        vals = [gbs]
    if rank == 0:
        print(f"{tag:58s} per rank GB/s: " + " ".join(f"{v:6.1f}" for v in vals) + f"   sum {sum(vals):7.1f}", flush=True)


nbytes = 184_550_400
host = [torch.empty(nbytes // 4, dtype=torch.float32).pin_memory() for _ in range(3)]
for h in host:
    h.normal_()
dst = torch.empty(nbytes // 4, device=dev)
work = torch.randn(8192, 8192, device=dev)


def copy_rate(n=5, stream=None, busy=False):
    s = stream or torch.cuda.current_stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record()
        for i in range(n):
            dst.copy_(host[i % 3], non_blocking=True)
            if busy:
                with torch.cuda.stream(torch.cuda.default_stream()):
                    torch.mm(work, work)
        e1.record()
    torch.cuda.synchronize()
    return nbytes * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


if rank == 0:
    os.system("nvidia-smi topo -m 2>/dev/null | head -14; nproc; numactl -H 2>/dev/null | head -6")
barrier()
for rep in range(2):
    for r in range(world):  # one rank at a time
        barrier()
        g = copy_rate() if r == rank else 0.0
        barrier()
        report(f"[{rep}] only rank {r} copies", g)
    barrier()
    report(f"[{rep}] all ranks copy at once", copy_rate())
    barrier()
    report(f"[{rep}] all ranks, copy stream, GEMMs on the default stream", copy_rate(stream=torch.cuda.Stream(), busy=True))
    barrier()
    # fresh device allocations per copy, as tensor.to(device, non_blocking=True) does in bench.py
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keep = [host[i % 3].to(dev, non_blocking=True) for i in range(5)]
    e1.record()
    torch.cuda.synchronize()
    report(f"[{rep}] all ranks, .to(device) with fresh allocations", nbytes * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del keep
    # many small tensors (9 per batch) instead of one large one
    parts = [h.view(9, -1) for h in host]
    torch.cuda.synchronize()
    e0.record()
    for i in range(5):
        for j in range(9):
            dst.view(9, -1)[j].copy_(parts[i % 3][j], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    report(f"[{rep}] all ranks, 9 copies of 20 MB per step", nbytes * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
if world > 1:
    dist.destroy_process_group()
