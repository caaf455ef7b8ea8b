#!/usr/bin/env python3
"""Development check (torchrun, N ranks): pinned D2H rate of 196 MB against the CPU the pinned buffer was allocated
from (the box hides its NUMA layout from sysfs), all ranks copying at once."""
import os, torch, torch.distributed as dist, subprocess
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
allc = sorted(os.sched_getaffinity(0))
if rank == 0:
    print("cpus", len(allc), subprocess.run("lscpu | grep -i -E 'numa|socket|model name'", shell=True, capture_output=True, text=True).stdout)
src = torch.empty(196048512 // 8, dtype=torch.float64, device=dev)
n = len(allc)
cands = {"all": allc, "first": allc[:1], "q0": allc[: n // 4], "q1": allc[n // 4: n // 2], "q2": allc[n // 2: 3 * n // 4], "q3": allc[3 * n // 4:],
         "own": allc[rank * n // world: (rank + 1) * n // world]}
for name, cpus in cands.items():
    os.sched_setaffinity(0, cpus)
    h = torch.empty(src.numel(), dtype=torch.float64, pin_memory=True)
    h.zero_()
    best = 1e9
    for k in range(4):
        dist.barrier(); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); h.copy_(src, non_blocking=True); b.record(); b.synchronize()
        if k: best = min(best, a.elapsed_time(b))
    print("rank %d pinned buffer allocated on cpus %-5s (%2d): %.2f ms  %.1f GB/s" % (rank, name, len(cpus), best, 196.048512 / best), flush=True)
    del h
    os.sched_setaffinity(0, allc)
dist.destroy_process_group()
