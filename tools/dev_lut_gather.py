#!/usr/bin/env python3
"""Development check (torchrun): per-rank times of the sharded C5 LUT kernels and the all-gather that follows."""
import os, sys
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gort_b200
from gort_b200 import workloads as wk
from gort_b200.api import LUT_STRIDE
from gort_b200.parallel import shard_range
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g = gort_b200.Gort(lr)
st = wk.c5_lut_grid()["structure"]; M = st.shape[1]
lo, hi = shard_range(M, rank, world)
d_blk = torch.from_numpy(np.ascontiguousarray(st[:, lo:hi])).to(dev)
d_loc = torch.empty((hi - lo, LUT_STRIDE), dtype=torch.float64, device=dev)
d_all = torch.empty((M, LUT_STRIDE), dtype=torch.float64, device=dev)
ts = torch.cuda.Stream(device=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "ts"
for it in range(5):
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    if mode == "ts":
        with torch.cuda.stream(ts):
            ev[0].record(ts); g.lut_dev(d_blk, d_loc, stream=ts.cuda_stream); ev[1].record(ts)
            dist.all_gather_into_tensor(d_all, d_loc); ev[2].record(ts)
    else:
        cur = torch.cuda.current_stream()
        ev[0].record(cur); g.lut_dev(d_blk, d_loc, stream=cur.cuda_stream); ev[1].record(cur)
        dist.all_gather_into_tensor(d_all, d_loc); ev[2].record(cur)
    torch.cuda.synchronize()
    print("mode %s it %d rank %d: kernels %.3f ms, gather %.3f ms" % (mode, it, rank, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])), flush=True)
g.close(); dist.destroy_process_group()
