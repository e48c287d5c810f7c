"""Pinned host -> device copy bandwidth on this box (informs the e2e ceiling of bench.py)."""
import torch, time
for mb in (1, 4, 16, 64, 256):
    n = mb * 2**20 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D {mb} MiB: {ms*1e3:.1f} us  {mb*2**20/ms/1e6:.1f} GB/s")
    e0.record()
    for _ in range(20):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"D2H {mb} MiB: {ms*1e3:.1f} us  {mb*2**20/ms/1e6:.1f} GB/s")
import os
print("cpus", os.cpu_count())
