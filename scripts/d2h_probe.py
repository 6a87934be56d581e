"""Platform probe under torchrun: what does ONE rank's device->host copy get when N ranks copy at the same time?
Each rank copies a 2 GiB device buffer into page-locked host memory `reps` times, all ranks between barriers, first with
the process left wherever the launcher put it, then after numa.bind_to_gpu + a fresh page-locked allocation.
    python -m torch.distributed.run --nproc-per-node N scripts/d2h_probe.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pyradiomics_b200 import numa

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 2 << 30
src = torch.empty(n, dtype=torch.uint8, device="cuda")
src.fill_(1)
out = {}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(tag, reps=4):
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.copy_(src)                       # faults the block in
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        host.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([reps * n / dt / 1e9], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        vals = [float(x.item()) for x in lst]
    else:
        vals = [float(t.item())]
    # the same copy with only THIS rank active (the others wait)
    solo = []
    for r in range(world):
        barrier()
        if r == rank:
            t0 = time.perf_counter()
            for _ in range(reps):
                host.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            solo.append(reps * n / (time.perf_counter() - t0) / 1e9)
    barrier()
    s = torch.tensor([solo[0]], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(s) for _ in range(world)]
        dist.all_gather(lst, s)
        solo_all = [float(x.item()) for x in lst]
    else:
        solo_all = [float(s.item())]
    out[tag] = {"concurrent_GBps_per_rank": [round(v, 1) for v in vals], "concurrent_sum_GBps": round(sum(vals), 1),
                "solo_GBps_per_rank": [round(v, 1) for v in solo_all]}
    del host


run("unbound")
cpus = numa.bind_to_gpu(local)
run("bound_to_gpu_numa_node")
out["cpus_after_bind"] = len(cpus)
if rank == 0:
    print(json.dumps({"d2h_probe": out, "world": world, "bytes_per_copy": n}))
if world > 1:
    dist.destroy_process_group()
