"""Concurrent pinned host->device (and device->host) copy bandwidth with NO kernels running, one process per GPU:
what the host side of the box can feed when N ranks push their 4.32 GB step input at the same time (the ceiling of
bench.py's e2e figure at N GPUs).

    python tools/h2d_concurrent.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29655 \
        tools/h2d_concurrent.py                                      # N GPUs at once

Rank 0 prints one JSON line: per-rank GB/s (H2D alone, D2H alone, both directions at once), the aggregate, and where
each rank's pinned buffer lives (NUMA node of the process, CPU affinity)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

GB = float(os.environ.get("H2D_GB", "4.32"))
n = int(GB * 1e9) // 4
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
host.fill_(1.0)
back = torch.empty(n // 6, dtype=torch.float32, pin_memory=True)   # the step's results are ~1/6 of its input
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.ones(n // 6, dtype=torch.float32, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s_in):
                d_in.copy_(host, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                back.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return (reps * host.numel() * 4 / dt / 1e9 if h2d else 0.0), (reps * back.numel() * 4 / dt / 1e9 if d2h else 0.0)


run(True, True, 1)
res = {"h2d_alone": run(True, False)[0], "d2h_alone": run(False, True)[1]}
both = run(True, True)
res["h2d_with_d2h"], res["d2h_with_h2d"] = both
try:
    res["cpus"] = sorted(os.sched_getaffinity(0))[:4] + ["..."] + [len(os.sched_getaffinity(0))]
except Exception:
    pass
try:
    bus = torch.cuda.get_device_properties(dev).pci_bus_id
    res["numa_node"] = open(f"/sys/bus/pci/devices/0000:{bus:02x}:00.0/numa_node").read().strip()
except Exception:
    pass
t = torch.tensor([res["h2d_alone"], res["d2h_alone"], res["h2d_with_d2h"], res["d2h_with_h2d"]], dtype=torch.float64, device=dev)
if world > 1:
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
else:
    allr = [t]
if rank == 0:
    rows = [[round(float(x), 2) for x in r.tolist()] for r in allr]
    print(json.dumps({"n_gpus": world, "gb_per_copy": GB, "columns": ["h2d_alone", "d2h_alone", "h2d_with_d2h", "d2h_with_h2d"],
                      "per_rank_gbs": rows, "sum_gbs": [round(sum(r[i] for r in rows), 2) for i in range(4)],
                      "rank0": {k: v for k, v in res.items() if k in ("cpus", "numa_node")},
                      "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
