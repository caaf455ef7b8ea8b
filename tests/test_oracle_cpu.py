"""CPU tests (no GPU): pin the oracle.

  * the plain-C restatement (oracle/libgort_oracle.so) against the committed golden vectors that were
    generated from the unmodified reference (tests/golden/make_golden.py) -- BIT-EXACT;
  * against the live reference library oracle/_ref when it is present (the build container) on fresh
    seeded inputs -- BIT-EXACT;
  * SURVEY.md App. E golden numbers;
  * an independent cross-check of the PROSPECT-D exponential-integral polynomial against scipy.
"""
from pathlib import Path

import numpy as np
import pytest

import checkers
from gort_b200 import workloads as wk

GOLD = Path(__file__).resolve().parent / "golden" / "ref_vectors.npz"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_lut_full_bit_exact_vs_golden(oracle, gold):
    for k, st in enumerate(gold["structure"]):
        assert same(oracle.lut(st, 0), gold["lut_full"][k]), "full LUT, structure %d" % k


def test_lut_q08_bit_exact_vs_golden(oracle, gold):
    for k, st in enumerate(gold["structure"]):
        assert same(oracle.lut(st, 1), gold["lut_q08"][k]), "Q08 LUT, structure %d" % k


def test_spectra_bit_exact_vs_golden(oracle, gold):
    for k in range(8):
        rl, tl, rs = oracle.spectra(gold["leaf"][k], gold["soil"][k], gold["wavelength"])
        assert same(rl, gold["rleaf"][k]) and same(tl, gold["tleaf"][k]) and same(rs, gold["rsoil"][k])


def test_brdf_bit_exact_vs_golden(oracle, gold):
    for k in range(8):
        r, s, kp = oracle.brdf(gold["structure"][k], gold["lut_full"][k], gold["angles"], gold["rleaf"][k],
                               gold["tleaf"][k], gold["rsoil"][k])
        assert same(r, gold["rsurf"][k]) and same(s, gold["scomp"][k]) and same(kp, gold["kprop"][k])
    r, _, _ = oracle.brdf(gold["structure"][0], gold["lut_full"][0], gold["angles"], gold["rleaf"][0],
                          gold["tleaf"][0], gold["rsoil"][0], beta=0.3, fd=0.8)
    assert same(r, gold["rsurf_beta_fd"])


def test_grazing_angles_in_golden_set(oracle, gold):
    """Lines 2 and 3 of the golden geometry set have a zenith > 89 deg, where the interpolation touches
    LUT row 90 (epgap = 0 there, gortt_pn_kopen.c:1099).  With a computed LUT the lerp still gives a
    positive probability, so the reference stays finite; with rows that underflow to 0 the hotspot's
    -log(0) gives NaN.  Either way the oracle must reproduce the reference."""
    lut0 = gold["lut_full"][0].copy()
    assert np.isfinite(gold["rsurf"][0][2]).all() and np.isfinite(gold["rsurf"][0][3]).all()
    lut0[91 + 89] = 0.0                                   # force the degenerate case
    r, _, _ = oracle.brdf(gold["structure"][0], lut0, gold["angles"][2:4], gold["rleaf"][0], gold["tleaf"][0], gold["rsoil"][0])
    assert np.isnan(r).all()


def test_energy_bit_exact_vs_golden(oracle, gold):
    for k in range(3):
        a, v, s = oracle.energy(gold["structure"][k], gold["lut_full"][k], gold["angles"][:3], gold["rleaf"][k],
                                gold["tleaf"][k], gold["rsoil"][k])
        assert same(a, gold["albedo"][k]) and same(v, gold["favegt"][k]) and same(s, gold["fasoil"][k])


def test_gauleg_bit_exact_vs_golden(oracle, gold):
    x, w = oracle.gauleg(32)
    assert same(x, gold["gauleg_x"]) and same(w, gold["gauleg_w"])
    assert abs(w.sum() - 2.0) < 1e-10


def test_survey_appendix_e_vectors(oracle):
    """SURVEY.md App. E1/E2/E4 (captured from the reference at -O0 during the survey)."""
    st = wk.structure_from_options(lai=4.0)
    assert st[5] == 1.5119088288082658 and st[2] == 2.6999988000000004
    lut = oracle.lut(st)
    assert [lut[0], lut[91], lut[30], lut[121], lut[60], lut[151], lut[182], lut[183]] == [
        0.4827649653877204, 0.049784714005155141, 0.18980914239271468, 0.054402229555638691,
        0.010675513076152983, 0.017277579752321828, 0.11690037154524389, 0.0348709390952388]
    wl = np.array([450.0, 600.0, 800.0, 1000.0])
    rl, tl, rs = oracle.spectra(wk.DEFAULT_LEAF, wk.DEFAULT_SOIL, wl, user_leaf=0.5)
    r, s, k = oracle.brdf(st, lut, np.array([[10.0, 0.0, 30.0, 20.0]]), rl, tl, rs)
    assert r[0].tolist() == [0.050996023726925049, 0.056904932488151413, 0.063793005449302428, 0.068538876965779413]
    assert k[0].tolist() == [0.24500453130170924, 0.078537903246602814, 0.33486586946440577, 0.34159169598728212]
    assert s[0, 0].tolist() == [0.1841728181840469, 0.050410010000000005, 0.0011235478145234198, 0.0045010298653438939]
    a, v, f = oracle.energy(st, lut, np.array([[10.0, 0.0, 30.0, 20.0]]), rl, tl, rs)
    assert [a[0, 0], v[0, 0], f[0, 0]] == [0.063777031489123184, 0.68728814305278241, 0.24893482545809434]
    r2, _, _ = oracle.brdf(st, oracle.lut(st, 1), np.array([[10.0, 0.0, 30.0, 20.0]]), rl, tl, rs)
    assert r2[0].tolist() == [0.050742404003082275, 0.056485203146758124, 0.063178656471069655, 0.067789806911394482]


def test_survey_appendix_e3_e5(oracle):
    st = wk.structure_from_options(hb=2, br=1.5, pcc=0.6, lai=3.7)
    lut = oracle.lut(st)
    rl, tl, rs = oracle.spectra(wk.DEFAULT_LEAF, wk.DEFAULT_SOIL, np.array([650.0, 850.0]), user_leaf=0.5, user_soil=0.2)
    r, _, k = oracle.brdf(st, lut, np.array([[0.0, 0, 45, 0], [30.0, 0, 45, 180], [-30.0, 0, 45, 0]]), rl, tl, rs)
    assert r[0, 0] == 0.072842386767957165 and r[1, 0] == 0.056055941717577183 and r[2, 0] == 0.056055941717577183
    assert k[0].tolist() == [0.15624258875274599, 0.18606383793253523, 0.29494578823791651, 0.36274778507680228]
    st = wk.structure_from_options(lai=4.0)
    lut = oracle.lut(st)
    rl, tl, rs = oracle.spectra(wk.DEFAULT_LEAF, wk.DEFAULT_SOIL, np.array([469, 555, 645, 858.5, 1240, 1640, 2130.0]), user_leaf=0.5)
    r, _, k = oracle.brdf(st, lut, np.array([[0.0, 0, 30, 0]]), rl, tl, rs)
    assert r[0].tolist() == [0.038250941196766607, 0.042141216030473889, 0.045173279536650658, 0.052774846747902872,
                             0.065029544987653745, 0.073598099690711433, 0.064762248401535577]


def test_oracle_vs_live_reference_on_fresh_seeds(oracle, ref):
    """Only where oracle/_ref exists (it is compiled from /root/reference in the build container)."""
    rng = np.random.Generator(np.random.PCG64(99))
    st = wk.random_structures(rng, 4)
    leaf = wk.random_leaves(rng, 4)
    wl = np.sort(rng.uniform(400, 2500, 12))
    ang = np.stack([rng.uniform(-85, 89, 30), rng.uniform(-720, 720, 30), rng.uniform(-85, 89, 30), rng.uniform(-720, 720, 30)], axis=1)
    for m in range(4):
        lo, lr = oracle.lut(st[:, m]), ref.lut(st[:, m])
        assert same(lo, lr)
        assert same(oracle.lut(st[:, m], 1), ref.lut(st[:, m], 1))
        so, sr = oracle.spectra(leaf[:, m], wk.DEFAULT_SOIL, wl), ref.spectra(leaf[:, m], wk.DEFAULT_SOIL, wl)
        assert all(same(a, b) for a, b in zip(so, sr))
        bo, br = oracle.brdf(st[:, m], lo, ang, *so), ref.brdf(st[:, m], lr, ang, *sr)
        assert all(same(a, b) for a, b in zip(bo, br))
    eo, er = oracle.energy(st[:, 0], lo, ang[:2], *so), ref.energy(st[:, 0], lo, ang[:2], *so)
    assert all(same(a, b) for a, b in zip(eo, er))


def test_makefile_flag_build_of_the_reference_gives_the_same_bits(ref):
    """oracle/_ref/libgortt_ref_g.so: the reference compiled with its own makefile's CFLAGS (-Wall -g).  bench.py times it
    next to the canonical -O2 -ffp-contract=off build (SURVEY.md 8c / 8d say the two are bit-identical: checked here)."""
    slow = checkers.ref_makefile_flags()
    if slow is None:
        pytest.skip("oracle/_ref/libgortt_ref_g.so not built (no /root/reference here)")
    rng = np.random.Generator(np.random.PCG64(7))
    st = wk.random_structures(rng, 2)
    leaf = wk.random_leaves(rng, 2)
    wl = np.sort(rng.uniform(400, 2500, 9))
    ang = np.stack([rng.uniform(-85, 89, 20), rng.uniform(-720, 720, 20), rng.uniform(-85, 89, 20), rng.uniform(-720, 720, 20)], axis=1)
    for m in range(2):
        la, lb = ref.lut(st[:, m]), slow.lut(st[:, m])
        assert same(la, lb)
        sa, sb = ref.spectra(leaf[:, m], wk.DEFAULT_SOIL, wl), slow.spectra(leaf[:, m], wk.DEFAULT_SOIL, wl)
        assert all(same(a, b) for a, b in zip(sa, sb))
        assert all(same(a, b) for a, b in zip(ref.brdf(st[:, m], la, ang, *sa), slow.brdf(st[:, m], lb, ang, *sb)))
    assert all(same(a, b) for a, b in zip(ref.energy(st[:, 0], la, ang[:2], *sa), slow.energy(st[:, 0], la, ang[:2], *sa)))


def test_prospect_exponential_integral_matches_scipy(oracle):
    """Independent check of the transcription of the NAG S13AAF polynomials (prospect_DB.f90:107-138):
    tau(k) = (1-k) exp(-k) + k^2 E1(k)."""
    import ctypes as C
    from scipy.special import exp1
    f = oracle.lib.gort_oracle_plate_tau
    f.argtypes = [C.c_double]; f.restype = C.c_double
    for k in np.concatenate([np.geomspace(1e-6, 4.0, 60), np.linspace(4.0001, 84.9, 60)]):
        want = (1 - k) * np.exp(-k) + k * k * exp1(k)
        # for large k tau is a small difference of large terms: the algorithm itself loses digits there
        assert abs(f(float(k)) - want) <= (2e-12 if k <= 4 else 5e-10) * want, k
    assert f(0.0) == 1.0 and f(-1.0) == 1.0 and f(86.0) == 0.0


def test_prospect_output_is_physical(oracle):
    rl, tl, _ = oracle.spectra(wk.DEFAULT_LEAF, wk.DEFAULT_SOIL, np.arange(400.0, 2501.0))
    assert np.all(rl > 0) and np.all(tl >= 0) and np.all(rl + tl < 1)
    # green peak and NIR plateau of a healthy leaf
    assert rl[150] > rl[50] and rl[150] > rl[270] and rl[380:600].min() > 0.35


def test_ulp_sensitivity_harness(oracle):
    """The 1-ULP-libm variant of the oracle moves well-conditioned outputs by ~1e-15 only."""
    from checkers import sensitivity
    st = wk.structure_from_options(lai=4.0)
    lut = oracle.lut(st)
    s = sensitivity(lambda c: c.lut(st), lut)
    rel = s / np.maximum(np.abs(lut), 1e-12)
    assert 0 < rel.max() < 1e-9


def test_dead_intermediates_restatement_vs_live_reference(oracle, ref):
    """vb / fb / t_open / dt_open / dk_open / k_open[h] (gortt_pn_kopen.c:925-1078, :351-375): the restatement equals the
    compiled reference bit for bit."""
    rng = np.random.Generator(np.random.PCG64(808))
    st = wk.random_structures(rng, 4)
    for m in range(4):
        a, b = oracle.lut_dead(st[:, m]), ref.lut_dead(st[:, m])
        for k in a:
            assert same(a[k], b[k]), "%s, structure %d" % (k, m)
