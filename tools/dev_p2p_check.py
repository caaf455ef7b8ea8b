#!/usr/bin/env python3
"""Development check (2+ GPUs, one process): can the GPUs reach each other directly, and at what rate?"""
import torch, time
n = torch.cuda.device_count()
print("devices", n, "peer access 0->1:", torch.cuda.can_device_access_peer(0, 1) if n > 1 else None)
if n > 1:
    a = torch.empty(96 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda:0")
    b = torch.empty_like(a, device="cuda:1")
    for _ in range(3):
        b.copy_(a, non_blocking=True)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t0 = time.perf_counter()
    for _ in range(10):
        b.copy_(a, non_blocking=True)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    dt = (time.perf_counter() - t0) / 10
    print("cuda:0 -> cuda:1 copy of 96 MB: %.3f ms, %.1f GB/s" % (dt * 1e3, a.numel() * 8 / dt / 1e9))
