"""CPU tests of the N > 1 host logic with the gloo backend (world_size 2 and 3): shard ranges, the padded
row all-gather that assembles LUTs, and shard invariance (N ranks assemble the same bits one rank
computes).  The per-shard compute is the oracle here (test infrastructure); on GPUs it is
gort_lut_batch_dev and the collective is the same call over NCCL."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from gort_b200.parallel import allgather_rows, lut_generate_sharded, pipelined_blocks, shard_counts, shard_range  # noqa: E402
from gort_b200 import workloads as wk  # noqa: E402


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 9, 131072, 100000):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == shard_counts(n, world)


def test_pipelined_blocks_tile_the_grid_in_gather_order():
    """super-block j = the pieces of ranks 0..world-1 in order: gathering piece j of every rank fills a contiguous range"""
    for M, world, n_sub in ((131072, 2, 4), (131072, 8, 4), (96, 3, 2), (64, 1, 4)):
        order = [pipelined_blocks(M, r, world, n_sub)[j] for j in range(n_sub) for r in range(world)]
        assert order[0][0] == 0 and order[-1][1] == M
        assert all(order[i][1] == order[i + 1][0] for i in range(len(order) - 1))
        assert len({b[1] - b[0] for b in order}) == 1
    assert pipelined_blocks(100, 0, 3, 4) is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from checkers import oracle
        o = oracle()
        rng = np.random.Generator(np.random.PCG64(5))
        st = wk.random_structures(rng, n_sets)

        def compute_local(block):
            return torch.from_numpy(np.stack([o.lut(block[:, k], 1) for k in range(block.shape[1])]).reshape(-1, 184))

        luts = lut_generate_sharded(st, compute_local, rank, world)
        # generic padded all-gather with an uneven split
        lo, hi = shard_range(n_sets, rank, world)
        rows = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)
        gathered = allgather_rows(rows, n_sets, rank, world)
        q.put((rank, luts.numpy(), gathered.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_sets", [(2, 7), (2, 8), (3, 10)])
def test_lut_allgather_shard_invariance(world, n_sets):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sets, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # serial answer
    from checkers import oracle
    o = oracle()
    rng = np.random.Generator(np.random.PCG64(5))
    st = wk.random_structures(rng, n_sets)
    serial = np.stack([o.lut(st[:, k], 1) for k in range(n_sets)])
    for rank, luts, gathered in results:
        assert luts.shape == (n_sets, 184)
        assert np.array_equal(luts, serial), "rank %d assembled different bits" % rank
        assert np.array_equal(gathered[:, 0], np.arange(n_sets, dtype=np.float64))


def test_peer_table_addresses_skip_the_own_rank_and_step_by_whole_records():
    """Address arithmetic of PeerLutTable (the symmetric allocation itself needs GPUs; bench.py / tools/dev_lut_peer.py
    check the stores under torchrun)."""
    from gort_b200.parallel import PeerLutTable
    from gort_b200.api import LUT_STRIDE
    t = PeerLutTable.__new__(PeerLutTable)
    t.rank, t.world, t.rows = 2, 4, 1000
    t.ptrs = [0x1000000, 0x2000000, 0x3000000, 0x4000000]
    t.multicast_ptr = 0x9000000
    lo = shard_range(t.rows, t.rank, t.world)[0]
    want = [p + lo * LUT_STRIDE * 8 for r, p in enumerate(t.ptrs) if r != 2]
    assert t.peer_addresses(lo) == want and len(want) == 3
    assert t.multicast_address(lo) == 0x9000000 + lo * 184 * 8
    t.multicast_ptr = 0
    assert t.multicast_address(lo) == 0
