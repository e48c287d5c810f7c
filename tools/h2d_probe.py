#!/usr/bin/env python
"""Pinned host -> device copy bandwidth on this box (the ceiling of bench.py's end-to-end number).

    python tools/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
                                                                # all GPUs copying AT THE SAME TIME: the host's aggregate ceiling
Under torchrun every rank copies from its own pinned buffer to its own GPU between two barriers; rank 0 prints the per-GPU and
the aggregate rate plus the NUMA node / CPU affinity nvidia-smi reports for each GPU.
"""
import json
import os
import subprocess

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

out = {"world": world, "sizes": {}}
for mb in (1, 16, 64, 256):
    n = mb * 2**20 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = mb * 2**20 / (e0.elapsed_time(e1) / reps) / 1e6
    t = torch.tensor([gbs], device="cuda", dtype=torch.float64)
    if world > 1:
        allg = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allg, t)
        rates = [float(x.item()) for x in allg]
    else:
        rates = [gbs]
    out["sizes"][f"{mb} MiB"] = {"per_gpu_GBs": [round(r, 1) for r in rates], "aggregate_GBs": round(sum(rates), 1)}
if rank == 0:
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        out["topo"] = [ln for ln in topo.split("\n") if ln.startswith("GPU")][: world + 1]
    except Exception:
        pass
    out["cpus"] = os.cpu_count()
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
