"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on GPUs, gloo on CPU for
the host-logic tests).

The GORT path shards embarrassingly: every (parameter set, input line) unit is
independent, so ranks own contiguous blocks of parameter sets (or of lines) and never exchange data on
the compute path.  The ONE collective is the all-gather that assembles gap-probability LUTs on every
rank (BASELINE.json config 5), after which rank 0 writes the reference's "-W" text layout.

Nothing here does model arithmetic: `compute_local` is the GPU call (Gort.lut_dev) in production and
the oracle in the CPU tests.
"""
import os
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

from .api import LUT_STRIDE, lut_write_text


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n units owned by `rank`: sizes differ by at most one, larger blocks first."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_counts(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def init_distributed(backend=None):
    """Read RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun); returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def allgather_rows(local, n_total, rank, world):
    """All-gather row blocks of a [n_local, K] tensor into [n_total, K] on every rank.

    Blocks follow shard_range(); ranks with one row fewer pad to the common block size so that a single
    all_gather_into_tensor (one NCCL all-gather over NVLink / NVSwitch) moves everything."""
    if world == 1:
        assert local.shape[0] == n_total
        return local
    counts = shard_counts(n_total, world)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    cmax = max(counts)
    K = local.shape[1]
    if local.shape[0] < cmax:
        pad = torch.zeros((cmax - local.shape[0], K), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * cmax, K), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    if all(c == cmax for c in counts):
        return out
    keep = torch.cat([torch.arange(r * cmax, r * cmax + counts[r], device=local.device) for r in range(world)])
    return out.index_select(0, keep)


def lut_generate_sharded(structure, compute_local, rank, world, device=None):
    """LUTs for all M parameter sets of `structure` ([6][M] numpy), sharded by parameter set.

    compute_local(structure_block [6][m]) -> [m][184] tensor on `device`.
    Returns the assembled [M][184] tensor, identical on every rank."""
    M = structure.shape[1]
    lo, hi = shard_range(M, rank, world)
    block = np.ascontiguousarray(structure[:, lo:hi])
    local = compute_local(block)
    assert local.shape == (hi - lo, LUT_STRIDE)
    return allgather_rows(local, M, rank, world)


def pipelined_blocks(n_total, rank, world, n_sub):
    """Block layout for LUT generation with the all-gather hidden under the kernels: the n_total parameter sets are cut
    into n_sub super-blocks, each super-block is split over the ranks in equal pieces.  Gathering piece j of every rank
    fills super-block j of the assembled array in natural order, so the gather of super-block j can run while the
    kernels of super-block j+1 do.  Returns [(lo, hi)] of this rank's piece in every super-block, or None if the sizes
    do not divide evenly."""
    if n_sub < 1 or n_total % (n_sub * world) != 0:
        return None
    piece = n_total // (n_sub * world)
    return [(j * piece * world + rank * piece, j * piece * world + (rank + 1) * piece) for j in range(n_sub)]


def lut_generate_pipelined(structure, gort, rank, world, device, n_sub=4, method=0, compute_stream=None, d_blocks=None):
    """LUTs of all M sets on every rank (GPU only): kernels of super-block j+1 on `compute_stream` while one NCCL
    all-gather assembles super-block j on a second stream.  Returns (assembled [M][184] tensor, d_blocks) -- pass
    d_blocks back in to reuse the device copies of this rank's structure pieces."""
    M = structure.shape[1]
    blocks = pipelined_blocks(M, rank, world, n_sub)
    assert blocks is not None, "M must divide by n_sub * world"
    cs = compute_stream or torch.cuda.current_stream()
    if d_blocks is None:
        d_blocks = [torch.from_numpy(np.ascontiguousarray(structure[:, lo:hi])).to(device) for lo, hi in blocks]
    piece = blocks[0][1] - blocks[0][0]
    out = torch.empty((M, LUT_STRIDE), dtype=torch.float64, device=device)
    loc = [torch.empty((piece, LUT_STRIDE), dtype=torch.float64, device=device) for _ in range(n_sub)]
    gs = getattr(lut_generate_pipelined, "_gather_stream", None)
    if gs is None or gs.device != torch.device(device):
        gs = torch.cuda.Stream(device=device)
        lut_generate_pipelined._gather_stream = gs
    gs.wait_stream(cs)                                  # `out` and `loc` were allocated on the compute stream
    for j in range(n_sub):
        gort.lut_dev(d_blocks[j], loc[j], method, stream=cs.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(cs)
        with torch.cuda.stream(gs):
            gs.wait_event(ev)
            if world > 1:
                dist.all_gather_into_tensor(out[j * piece * world:(j + 1) * piece * world], loc[j])
            else:
                out[j * piece:(j + 1) * piece].copy_(loc[j], non_blocking=True)
    cs.wait_stream(gs)
    for t in loc:
        t.record_stream(gs)
    return out, d_blocks


class PeerLutTable:
    """A [rows][184] LUT table that exists on every rank of the group, with every rank's copy mapped into every
    process (torch symmetric memory: cuMem allocations whose handles are exchanged once, NVLink P2P), plus the NVSwitch
    multicast address of the table when the fabric offers one.  Plumbing only: the kernels that store into the peers'
    copies are the library's own (gort_lut_batch_scatter_dev)."""

    def __init__(self, rows, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.rows = rows
        self.group = group if group is not None else dist.group.WORLD
        self.table = symm.empty((rows, LUT_STRIDE), dtype=torch.float64, device=device)
        self.handle = symm.rendezvous(self.table, self.group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        try:
            self.multicast_ptr = int(self.handle.multicast_ptr or 0)
        except Exception:
            self.multicast_ptr = 0

    def peer_addresses(self, row):
        """Addresses of `row` inside the other ranks' tables."""
        off = row * LUT_STRIDE * 8
        return [p + off for r, p in enumerate(self.ptrs) if r != self.rank]

    def multicast_address(self, row):
        return self.multicast_ptr + row * LUT_STRIDE * 8 if self.multicast_ptr else 0

    def barrier(self):
        """Device-side barrier of the group on the current stream (signal pads in symmetric memory)."""
        self.handle.barrier()


def lut_generate_peer(d_block, table, gort, stream, method=0, multicast=False, lo=None):
    """This rank's block of LUT records computed into `table` on EVERY rank, no collective: the kernels that produce a
    row also store it into the peers' tables (or once to the multicast address).  d_block: [6][m] device tensor, the
    rank's contiguous block shard_range(table.rows, rank, world) (or starting at `lo`).  Everything -- a barrier of the
    ranks (nobody still reads the old table), the kernels, a second barrier (all rows have landed everywhere) -- is
    enqueued on `stream`, a torch.cuda.Stream other than the default one (the library maps a null stream to the
    context's own, which the barriers would not be ordered with)."""
    assert stream.cuda_stream != 0, "lut_generate_peer needs an explicit (non-default) stream"
    if lo is None:
        lo = shard_range(table.rows, table.rank, table.world)[0]
    m = d_block.shape[1]
    if multicast:
        assert table.multicast_ptr, "no multicast address for this table"
        dst = [table.multicast_address(lo)]
    else:
        dst = table.peer_addresses(lo)
    with torch.cuda.stream(stream):
        table.barrier()
        gort.lut_scatter_dev(d_block, table.table[lo:lo + m], dst, method, multicast=multicast, stream=stream.cuda_stream)
        table.barrier()
    return table.table


def write_lut_directory(luts, out_dir, names=None):
    """Rank 0: one "-W"-layout text file per parameter set (gortt.c:123-128), readable by `gortt -P`."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    luts = luts.detach().cpu().numpy() if isinstance(luts, torch.Tensor) else np.asarray(luts)
    for k in range(luts.shape[0]):
        name = names[k] if names is not None else "lut_%06d.txt" % k
        lut_write_text(luts[k], str(out_dir / name))
    return luts.shape[0]
