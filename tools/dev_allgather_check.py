#!/usr/bin/env python3
"""Development check (torchrun, N ranks): NCCL all_gather_into_tensor of the C5 LUT block sizes, CUDA-event timed."""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
M, K = 131072, 184
loc = torch.randn((M // world, K), dtype=torch.float64, device="cuda")
out = torch.empty((M, K), dtype=torch.float64, device="cuda")
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); dist.all_gather_into_tensor(out, loc); b.record(); b.synchronize()
    if rank == 0:
        ms = a.elapsed_time(b)
        print("all_gather %d ranks, %d MB total: %.3f ms, %.1f GB/s received per rank" % (world, M * K * 8 // 1000000, ms, M * K * 8 * (world - 1) / world / ms / 1e6))
dist.destroy_process_group()
