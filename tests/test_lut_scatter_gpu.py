"""gort_lut_batch_scatter_dev: the LUT kernels store every row they produce into further tables as well (on several GPUs
those are the peers' tables, tools / bench.py check that path under torchrun; here the further tables are plain buffers of
the same GPU, which exercises the same stores).  Every copy must hold the bits of gort_lut_batch_dev."""
import numpy as np
import pytest
import torch

import gort_b200
from gort_b200 import workloads as wk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    h = gort_b200.Gort(0)
    yield h
    h.close()


@pytest.mark.parametrize("method", [gort_b200.LUT_FULL, gort_b200.LUT_Q08])
@pytest.mark.parametrize("grouped", [False, True])
def test_every_destination_holds_the_bits_of_the_plain_call(g, method, grouped):
    dev = torch.device("cuda:0")
    if grouped:
        st = np.ascontiguousarray(wk.c5_lut_grid((3, 3, 2, 3, 5, 7))["structure"])     # shape groups with sub-groups of 8
    else:
        st = wk.random_structures(np.random.Generator(np.random.PCG64(11)), 333)         # every set its own crown shape
    M = st.shape[1]
    d_st = torch.from_numpy(st).to(dev)
    want = torch.empty((M, gort_b200.LUT_STRIDE), dtype=torch.float64, device=dev)
    g.lut_dev(d_st, want, method)
    g.synchronize()
    # the block lands at row 5 of three larger tables: the local one and two further ones
    tabs = [torch.full((M + 9, gort_b200.LUT_STRIDE), -3.0, dtype=torch.float64, device=dev) for _ in range(3)]
    row = 5
    dst = [t.data_ptr() + row * gort_b200.LUT_STRIDE * 8 for t in tabs[1:]]
    g.lut_scatter_dev(d_st, tabs[0][row:row + M], dst, method)
    g.synchronize()
    for t in tabs:
        assert torch.equal(t[row:row + M].view(torch.int64), want.view(torch.int64))
        assert bool((t[:row] == -3.0).all()) and bool((t[row + M:] == -3.0).all())       # nothing outside the block


def test_no_further_destination_is_the_plain_call(g):
    dev = torch.device("cuda:0")
    st = wk.random_structures(np.random.Generator(np.random.PCG64(12)), 40)
    d_st = torch.from_numpy(st).to(dev)
    a = torch.empty((40, gort_b200.LUT_STRIDE), dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    g.lut_dev(d_st, a)
    g.lut_scatter_dev(d_st, b, [])
    g.synchronize()
    assert torch.equal(a.view(torch.int64), b.view(torch.int64))


def test_bad_destinations_are_refused(g):
    dev = torch.device("cuda:0")
    st = wk.random_structures(np.random.Generator(np.random.PCG64(13)), 8)
    d_st = torch.from_numpy(st).to(dev)
    out = torch.empty((8, gort_b200.LUT_STRIDE), dtype=torch.float64, device=dev)
    with pytest.raises(gort_b200.GortError):
        g.lut_scatter_dev(d_st, out, [0])                                   # null destination
    with pytest.raises(gort_b200.GortError):
        g.lut_scatter_dev(d_st, out, [out.data_ptr() + 4])                  # not 8-byte aligned
    with pytest.raises(gort_b200.GortError):
        g.lut_scatter_dev(d_st, out, [out.data_ptr()] * 17)                 # more than GORT_LUT_MAX_DST
