"""CPU self-test of the parity-audit machinery (tests/parity_audit.py): the summaries, the merge, and one whole
audit pass in which the "GPU" is played by the plain-C restatement on shrunken workloads (every error is then 0 and
every stage must pass); a perturbed result must be caught."""
import numpy as np

import parity_audit as pa
from gort_b200 import workloads as wk


class OracleAsGort:
    """Stands in for gort_b200.Gort in this self-test only: same method signatures, results from the restatement."""

    def __init__(self, oracle, poke=None):
        self.o, self.poke = oracle, poke

    def lut(self, st, method=0):
        return np.stack([self.o.lut(np.ascontiguousarray(st[:, m]), method) for m in range(st.shape[1])])

    def spectra(self, leaf, soil, wl):
        sp = [self.o.spectra(np.ascontiguousarray(leaf[:, m]), np.ascontiguousarray(soil[:, m]), wl) for m in range(leaf.shape[1])]
        return tuple(np.stack([s[k] for s in sp]) for k in range(3))

    def _each(self, st, lut, ang, rl, tl, rs, fn):
        M = st.shape[1]
        res = []
        for m in range(M):
            a = ang[:, m, :] if ang.ndim == 3 else ang
            sp = [x[m] if x.ndim == 2 else x for x in (rl, tl, rs)]
            res.append(fn(np.ascontiguousarray(st[:, m]), lut[m], np.ascontiguousarray(a.T), *sp))
        return [np.stack([r[k] for r in res]) for k in range(3)]

    def brdf(self, st, lut, ang, rl, tl, rs, want_scomp=False, want_kprop=False):
        r, s, k = self._each(st, lut, ang, rl, tl, rs, self.o.brdf)
        if self.poke:
            r = r.copy(); r.flat[self.poke] *= 1.0 + 1e-7
        out = (r,) + ((s,) if want_scomp else ()) + ((k,) if want_kprop else ())
        return out if len(out) > 1 else r

    def energy(self, st, lut, ang, rl, tl, rs):
        return tuple(self._each(st, lut, ang, rl, tl, rs, self.o.energy))


def test_compare_counts_and_excuses():
    ref = np.array([[1.0, 2.0, np.nan], [1e-15, 4.0, np.inf]])
    x = ref.copy()
    x[0, 0] += 5e-9            # beyond tolerance, unexcused
    x[0, 1] += 3e-8            # beyond tolerance, excused by sensitivity
    x[1, 0] += 1e-22           # below the floor: fine
    sens = np.zeros_like(ref); sens[0, 1] = 1e-9
    s = pa.compare(x, ref, sens, band_axis=-1)
    assert s["n"] == 6 and s["n_beyond_tol"] == 2 and s["n_excused"] == 1 and s["n_unexplained"] == 1
    assert s["nonfinite_ref"] == 2 and s["nonfinite_mismatch"] == 0
    assert abs(s["worst_excused_rel"] - 1.5e-8) < 1e-12 and abs(s["max_rel_err_not_excused"] - 5e-9) < 1e-12
    assert s["per_band"].shape == (3,) and not pa.passed(s)
    y = ref.copy(); y[0, 2] = 1.0          # NaN position lost
    assert pa.compare(y, ref)["nonfinite_mismatch"] == 1
    z = ref.copy(); z[1, 2] = -np.inf      # wrong kind of non-finite
    assert pa.compare(z, ref)["nonfinite_mismatch"] == 1
    m = pa.merge(pa.compare(x, ref, sens, band_axis=-1), pa.compare(ref, ref, band_axis=-1))
    assert m["n"] == 12 and m["n_unexplained"] == 1 and m["per_band"][0] > 4e-9


SMALL = {"c2_lines": 72, "c3_sets": 3, "c3_total": 3, "c4_members": 4, "c4_total": 4, "c5_sets": 4, "c5_grid": (2, 1, 1, 1, 2, 1)}


def test_audit_passes_on_identical_results(oracle):
    rep = pa.audit(OracleAsGort(oracle), sizes=SMALL, workers=2)
    assert rep["pass"], pa.failures(rep)
    assert set(rep["configs"]) == {"c1", "c2", "c3", "c4", "c5"}
    c2 = rep["configs"]["c2"]["outputs"]
    assert c2["rsurf"]["n"] == 72 * 2101 and len(c2["rsurf"]["max_rel_err_per_100nm_from"]) == 21
    assert {"kprop.Kc", "kprop.Kt", "scomp.C", "scomp.Z", "lut.epgap"} <= set(c2)
    assert len(rep["configs"]["c4"]["outputs"]["rsurf"]["max_rel_err_per_band"]) == 7
    assert len(rep["configs"]["c3"]["outputs"]["favegt"]["max_rel_err_per_band"]) == 211
    import json
    json.dumps(rep)                         # the report must be JSON-ready
    assert pa.headline(rep)["c2"]["rsurf"][3] == 0


def test_audit_catches_a_wrong_result(oracle):
    rep = pa.audit(OracleAsGort(oracle, poke=12345), sizes=SMALL, workers=2, configs=("c2",))
    assert not rep["pass"]
    assert rep["configs"]["c2"]["outputs"]["rsurf"]["n_unexplained"] == 1
    assert any("c2/rsurf" in f for f in pa.failures(rep))
