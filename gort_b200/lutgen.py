"""LUT generation over a structural-parameter grid on 1-8 GPUs (BASELINE.json config 5).

    python -m gort_b200.lutgen --grid 8,8,4,8,8,8 --out luts/            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m gort_b200.lutgen --grid 8,8,4,8,8,8 --out luts/                # eight GPUs

Each rank computes the KOpen / P(n) LUTs of its block of parameter sets with the CUDA kernels.  On several GPUs the
kernels store every row they produce into every rank's table over NVLink (gort_lut_batch_scatter_dev on symmetric
memory, `--assemble peer`, the default; `multicast` uses the NVSwitch multicast address); `--assemble allgather` runs
the kernels and then ONE NCCL all-gather (also what is used when the platform refuses symmetric memory).  Rank 0 writes
the records in the reference's "-W" text layout so that `gortt -P luts/lut_000123.txt ...` can consume them.
"""
import argparse
import json

import numpy as np
import torch

from . import workloads as wk
from .api import Gort, LUT_FULL, LUT_Q08, LUT_STRIDE
from .parallel import init_distributed, lut_generate_sharded, write_lut_directory


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="8,8,4,8,8,8", help="points along r, b/r, h1, h2-h1, cover, favd")
    ap.add_argument("--out", default=None, help="directory for the -W layout text files (rank 0)")
    ap.add_argument("--limit", type=int, default=0, help="only the first N grid points")
    ap.add_argument("--q08", action="store_true")
    ap.add_argument("--pipelined", action="store_true",
                    help="grid in 4 super-blocks, the all-gather of one under the kernels of the next (pays only when a rank's "
                         "share is several tens of thousands of sets: measured slower than kernels + one all-gather for C5 on 8 GPUs)")
    ap.add_argument("--assemble", choices=("peer", "multicast", "allgather"), default="peer",
                    help="how the ranks' blocks reach every rank (world > 1): stores from the producing kernels into the peers' "
                         "tables, the same through the NVSwitch multicast address, or one NCCL all-gather after the kernels")
    ap.add_argument("--write-max", type=int, default=4096, help="cap on the number of text files written")
    args = ap.parse_args()

    rank, local_rank, world = init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    g = Gort(local_rank)
    st = wk.c5_lut_grid(tuple(int(x) for x in args.grid.split(",")))["structure"]
    if args.limit:
        st = np.ascontiguousarray(st[:, :args.limit])
    M = st.shape[1]
    from .parallel import lut_generate_pipelined, pipelined_blocks, shard_range, PeerLutTable, lut_generate_peer
    ts = torch.cuda.Stream(device=dev)                   # created once, outside the timed region
    method = LUT_Q08 if args.q08 else LUT_FULL
    n_sub = 4 if args.pipelined and pipelined_blocks(M, rank, world, 4) is not None else 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    assemble = "allgather" if world == 1 else ("pipelined" if n_sub else args.assemble)
    tab = None
    if assemble in ("peer", "multicast"):
        try:
            tab = PeerLutTable(M, dev)                   # set-up: symmetric allocation + handle exchange, once
            if assemble == "multicast" and not tab.multicast_ptr:
                assemble = "peer"
        except Exception as e:
            if rank == 0:
                print("lutgen: symmetric memory unavailable (%s: %s); assembling with an all-gather" % (type(e).__name__, e), flush=True)
            assemble = "allgather"
    if tab is not None:
        lo, hi = shard_range(M, rank, world)
        d_blk = torch.from_numpy(np.ascontiguousarray(st[:, lo:hi])).to(dev)
        mc = assemble == "multicast"
        lut_generate_peer(d_blk, tab, g, ts, method=method, multicast=mc)     # warm-up
        torch.cuda.synchronize()
        torch.distributed.barrier()
        with torch.cuda.stream(ts):
            e0.record(ts)
            luts = lut_generate_peer(d_blk, tab, g, ts, method=method, multicast=mc)
            e1.record(ts)
    elif n_sub:
        # the grid in 4 super-blocks, each split over the ranks: the all-gather of one runs under the kernels of the next
        with torch.cuda.stream(ts):
            luts, d_blocks = lut_generate_pipelined(st, g, rank, world, dev, n_sub=n_sub, method=method, compute_stream=ts)   # warm-up
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            e0.record(ts)
            luts, d_blocks = lut_generate_pipelined(st, g, rank, world, dev, n_sub=n_sub, method=method, compute_stream=ts, d_blocks=d_blocks)
            e1.record(ts)
    else:
        lo, hi = shard_range(M, rank, world)
        d_blk = torch.from_numpy(np.ascontiguousarray(st[:, lo:hi])).to(dev)      # this rank's structure block: resident before timing
        d_loc = torch.empty((hi - lo, LUT_STRIDE), dtype=torch.float64, device=dev)

        def compute_local(block):
            with torch.cuda.stream(ts):
                g.lut_dev(d_blk, d_loc, method, stream=ts.cuda_stream)
            return d_loc

        with torch.cuda.stream(ts):
            lut_generate_sharded(st, compute_local, rank, world)              # warm-up: module load, NCCL communicator
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            e0.record(ts)
            luts = lut_generate_sharded(st, compute_local, rank, world)       # kernels, then ONE all_gather_into_tensor
            e1.record(ts)
    torch.cuda.synchronize()
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
    dt = float(dt.item())
    if rank == 0:
        n_written = 0
        if args.out:
            n_written = write_lut_directory(luts[:args.write_max], args.out)
        print(json.dumps({"luts": M, "gpus": world, "seconds": dt, "luts_per_s": M / dt, "assemble": assemble,
                          "table_bytes": M * LUT_STRIDE * 8, "files_written": n_written,
                          "nan_luts": int(torch.isnan(luts).any(dim=1).sum())}))
    g.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
