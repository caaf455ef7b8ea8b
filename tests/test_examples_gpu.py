"""The C example over the C ABI (examples/enkf_forward.c) gives the same bits as the same calls through the Python
binding: host code in C needs nothing but include/gort_b200.h and the shared library."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from gort_b200 import workloads as wk

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_enkf_forward_example(gort, tmp_path):
    exe = ROOT / "gort_b200" / "bin" / "enkf_forward"
    assert exe.exists(), "build with `make -C gort_b200/csrc`"
    w = wk.c4_enkf(n_members=300, seed=5)
    st, leaf, soil, wl, ang = w["structure"], w["leaf"], w["soil"], w["wavelength"], w["angles"]
    M, G, W = st.shape[1], ang.shape[2], wl.size
    src, dst = tmp_path / "members.bin", tmp_path / "out.bin"
    with open(src, "wb") as f:
        np.array([M, G, W], dtype=np.int32).tofile(f)
        for a in (st, leaf, soil, wl, ang):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
    p = subprocess.run([str(exe), str(src), str(dst)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    out = np.fromfile(dst, dtype=np.float64).reshape(M, G, W)
    lut = gort.lut(st)
    rl, tl, rs = gort.spectra(leaf, soil, wl)
    ref = gort.brdf(st, lut, ang, rl, tl, rs)
    assert np.array_equal(out, ref, equal_nan=True)


def test_lut_grid_multigpu_example(gort, tmp_path):
    """examples/lut_grid_multigpu.cu: one C++ host process, every visible GPU computing its block of a LUT grid and
    storing it into every GPU's table (gort_lut_batch_scatter_dev over peer mappings; with one GPU the degenerate case).
    The assembled table must hold the bits one GPU computes, on every GPU."""
    exe = ROOT / "gort_b200" / "bin" / "lut_grid_multigpu"
    assert exe.exists(), "build with `make -C gort_b200/csrc`"
    st = np.ascontiguousarray(wk.c5_lut_grid((3, 3, 2, 3, 5, 7))["structure"])
    M = st.shape[1]
    src, dst = tmp_path / "structure.bin", tmp_path / "lut.bin"
    with open(src, "wb") as f:
        np.array([M], dtype=np.int32).tofile(f)
        st.tofile(f)
    p = subprocess.run([str(exe), str(src), str(dst)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "tables identical: yes" in p.stdout
    out = np.fromfile(dst, dtype=np.float64).reshape(M, -1)
    assert np.array_equal(out, gort.lut(st), equal_nan=True)
