"""GPU test of gort_jacobian_batch: d rsurf / d LAI, d / d favd and d / d lambda by central
differences through the whole chain on the GPU, against the same difference quotient formed with the CPU checker's
chain (LUT -> spectra -> BRDF at the perturbed parameters).  The reference has no derivative code; its LAI -> favd map
(gortt.c:1127-1131) defines the LAI derivative.  Tolerance: 1e-7 of the largest Jacobian entry of the member (the
quotient amplifies the 1e-13 agreement of the two chains by 1 / (2 h) = 5000)."""
import numpy as np
import pytest

import gort_b200
from gort_b200 import workloads as wk

pytestmark = pytest.mark.gpu
H = 1e-4


def oracle_chain(oracle, st6, leaf, soil, wl, ang):
    lut = oracle.lut(st6)
    rl, tl, rs = oracle.spectra(leaf, soil, wl)
    return oracle.brdf(st6, lut, ang, rl, tl, rs, want_scomp=False)[0]


def oracle_jacobian(oracle, st6, leaf, soil, wl, ang, row, lai):
    sp, sm = st6.copy(), st6.copy()
    sp[row] *= 1.0 + H; sm[row] *= 1.0 - H
    p = st6[row]
    if lai:
        p = st6[5] * (st6[0] * st6[1] * st6[1] * np.pi * st6[2] * 4.0) / 3.0
    return (oracle_chain(oracle, sp, leaf, soil, wl, ang) - oracle_chain(oracle, sm, leaf, soil, wl, ang)) / (2.0 * H * p)


@pytest.mark.parametrize("param,row,lai", [(gort_b200.JAC_LAI, 5, True), (gort_b200.JAC_FAVD, 5, False), (gort_b200.JAC_LAMBDA, 0, False)])
def test_jacobian_against_checker_difference_quotient(gort, oracle, param, row, lai):
    w = wk.c4_enkf(n_members=24, seed=31)
    st, ang, wl = w["structure"], w["angles"], w["wavelength"]
    jac, rsurf = gort.jacobian(st, w["leaf"], w["soil"], wl, ang, param=param, rel_step=H, want_rsurf=True)
    assert np.array_equal(rsurf, gort.forward(st, w["leaf"], w["soil"], wl, ang), equal_nan=True)
    worst = 0.0
    for m in range(st.shape[1]):
        if not np.isfinite(rsurf[m]).all():
            continue
        want = oracle_jacobian(oracle, st[:, m].copy(), w["leaf"][:, m], w["soil"][:, m], wl, np.ascontiguousarray(ang[:, m, :].T), row, lai)
        scale = np.max(np.abs(want))
        err = np.max(np.abs(jac[m] - want)) / scale
        worst = max(worst, err)
        assert err <= 1e-7, "member %d: Jacobian differs by %.3e of its largest entry" % (m, err)
    print("param %d: worst |dJ| / max|J| = %.3e" % (param, worst))


def test_jacobian_step_consistency_and_lai_scaling(gort):
    w = wk.c4_enkf(n_members=300, seed=32)
    st, ang, wl = w["structure"], w["angles"], w["wavelength"]
    j1 = gort.jacobian(st, w["leaf"], w["soil"], wl, ang, param=gort_b200.JAC_LAI, rel_step=1e-4)
    j2 = gort.jacobian(st, w["leaf"], w["soil"], wl, ang, param=gort_b200.JAC_LAI, rel_step=4e-4)
    ok = np.isfinite(j1).all(axis=(1, 2)) & np.isfinite(j2).all(axis=(1, 2))
    assert ok.mean() > 0.98
    scale = np.max(np.abs(j1[ok]), axis=(1, 2), keepdims=True)
    assert np.max(np.abs(j1[ok] - j2[ok]) / scale) < 1e-5        # O(h^2) truncation: the chain is smooth in LAI
    jf = gort.jacobian(st, w["leaf"], w["soil"], wl, ang, param=gort_b200.JAC_FAVD, rel_step=1e-4)
    lai = st[5] * (st[0] * st[1] ** 2 * np.pi * st[2] * 4.0) / 3.0
    assert np.allclose(j1[ok], (jf * (st[5] / lai)[:, None, None])[ok], rtol=1e-12, atol=0)
    # more leaf area darkens the red band (645 nm) for a typical member and view
    assert np.median(j1[ok][:, :, 0]) < 0
    with pytest.raises(gort_b200.GortError):
        gort.jacobian(st, w["leaf"], w["soil"], wl, ang, param=9)
