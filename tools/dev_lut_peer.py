#!/usr/bin/env python3
"""Development check (torchrun): the C5 LUT grid assembled on every rank by the kernels' own peer stores
(gort_lut_batch_scatter_dev over symmetric memory; per-peer addresses and, when offered, the multicast address)
against kernels + one NCCL all-gather.  Bits must be equal; times are CUDA events on the launching stream, max over
ranks."""
import os, sys, json
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gort_b200
from gort_b200 import workloads as wk
from gort_b200.api import LUT_STRIDE
from gort_b200.parallel import shard_range, PeerLutTable, lut_generate_peer, allgather_rows
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g = gort_b200.Gort(lr)
st = wk.c5_lut_grid()["structure"]
if len(sys.argv) > 1:
    st = np.ascontiguousarray(st[:, :int(sys.argv[1])])
M = st.shape[1]
lo, hi = shard_range(M, rank, world)
d_blk = torch.from_numpy(np.ascontiguousarray(st[:, lo:hi])).to(dev)
d_loc = torch.empty((hi - lo, LUT_STRIDE), dtype=torch.float64, device=dev)
cur = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(cur)


def timed(fn, reps=4):
    best = None
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur); fn(); e1.record(cur); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = t.item() if best is None else min(best, t.item())
    return best


out = {"world": world, "luts": M}
ref = {}
def nccl():
    g.lut_dev(d_blk, d_loc, stream=cur.cuda_stream)
    ref["t"] = allgather_rows(d_loc, M, rank, world)
out["kernels_ms"] = timed(lambda: g.lut_dev(d_blk, d_loc, stream=cur.cuda_stream))
out["nccl_total_ms"] = timed(nccl)
try:
    tab = PeerLutTable(M, dev)
    out["multicast_ptr"] = bool(tab.multicast_ptr)
    tab.table.zero_(); torch.cuda.synchronize(); dist.barrier()
    out["peer_total_ms"] = timed(lambda: lut_generate_peer(d_blk, tab, g, cur))
    torch.cuda.synchronize(); dist.barrier()
    out["peer_equal_bits"] = bool(torch.equal(tab.table.view(torch.int64), ref["t"].view(torch.int64)))
    out["barrier_pair_ms"] = timed(lambda: (tab.barrier(), tab.barrier()))
    if tab.multicast_ptr:
        dist.barrier(); tab.table.zero_(); torch.cuda.synchronize(); dist.barrier()
        out["mc_total_ms"] = timed(lambda: lut_generate_peer(d_blk, tab, g, cur, multicast=True))
        torch.cuda.synchronize(); dist.barrier()
        out["mc_equal_bits"] = bool(torch.equal(tab.table.view(torch.int64), ref["t"].view(torch.int64)))
except Exception as e:                      # development script: report what the platform refused
    import traceback
    out["peer_error"] = "%s: %s" % (type(e).__name__, e)
    traceback.print_exc()
flags = torch.tensor([int(out.get("peer_equal_bits", False)), int(out.get("mc_equal_bits", True))], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
out["equal_on_every_rank"] = [bool(x) for x in flags.tolist()]
if rank == 0:
    print(json.dumps(out), flush=True)
g.close(); dist.destroy_process_group()
