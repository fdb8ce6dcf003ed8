"""What the box gives N concurrent pinned host->device streams (the ceiling of the end-to-end
path at N GPUs): every rank copies from its own pinned buffer to its own GPU for ~2 s, all ranks at
once, in the copy sizes the RDF feed uses; with and without binding the rank to its GPU's NUMA node
before the pinned allocation.

    python -m torch.distributed.run --nproc-per-node N scripts/h2d_ceiling.py > out.json
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200.utils import affinity  # noqa: E402

rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def measure(chunk_mb: float, seconds: float = 1.5):
    n = int(chunk_mb * (1 << 20) // 4)
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in host:
        h.fill_(1.0)  # first touch here, on this rank's CPUs
    dev = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(2)]
    stream = torch.cuda.Stream()
    for k in range(2):
        with torch.cuda.stream(stream):
            dev[k].copy_(host[k], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    copies = 0
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record()
        while time.perf_counter() - t0 < seconds:
            for k in range(2):
                dev[k].copy_(host[k], non_blocking=True)
            copies += 2
            if copies % 16 == 0:
                stream.synchronize()  # keep the queue short so the time check means something
        ev1.record()
    torch.cuda.synchronize()
    gbs = copies * n * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    mn, sm = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return {"chunk_mb": chunk_mb, "gbs_rank0": gbs, "gbs_min_rank": float(mn), "gbs_sum": float(sm)}


out = {"n_gpus": world, "unbound": [measure(mb) for mb in (4.0, 47.0, 190.0)]}
bind = affinity.bind_to_gpu(local)
out["binding_rank0"] = bind
out["bound_to_gpu_numa_node"] = [measure(mb) for mb in (4.0, 47.0, 190.0)]
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
