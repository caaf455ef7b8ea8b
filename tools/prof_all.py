"""ncu target: one small invocation of every kernel of the library (LUT full + Q08, spectra, BRDF wide / band /
scomp, energy) on shapes large enough to fill the GPU for a few hundred microseconds."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gort_b200
from gort_b200 import workloads as wk
g = gort_b200.Gort(0)
rng = np.random.Generator(np.random.PCG64(3))
st = wk.random_structures(rng, 2000)
lut = g.lut(st)                                     # lut_full_kernel, ungrouped shapes
grid = np.ascontiguousarray(wk.c5_lut_grid()["structure"][:, :16384])
lg = g.lut(grid)                                    # lut_full_kernel, shape groups of 64
lq = g.lut(st, gort_b200.LUT_Q08)                   # lut_q08_kernel
w = wk.c3_albedo(n_sets=600)
l2 = g.lut(w["structure"])
rl, tl, rs = g.spectra(w["leaf"], w["soil"], w["wavelength"])       # spectra_kernel
a, v, s = g.energy(w["structure"], l2, w["angles"], rl, tl, rs)     # energy_zenith_kernel, energy_kernel
w4 = wk.c4_enkf(n_members=20000)
l4 = g.lut(w4["structure"][:, :64])
l4 = np.repeat(l4[:1], 20000, axis=0)
r4, t4, s4 = g.spectra(w4["leaf"], w4["soil"], w4["wavelength"])
b4 = g.brdf(w4["structure"], l4, w4["angles"], r4, t4, s4)          # geom_kernel, rsurf_flat_kernel
w2 = wk.c2_hemisphere()
l1 = g.lut(w2["structure"])
r2, t2, s2 = g.spectra(w2["leaf"], w2["soil"], w2["wavelength"])
b2, sc = g.brdf(w2["structure"], l1, w2["angles"][:, :2000], r2[0], t2[0], s2[0], want_scomp=True)   # rsurf_wide_kernel<scomp>
print("ok", np.isfinite(lut).mean(), np.isfinite(a).mean(), np.isfinite(b4).mean(), np.isfinite(sc).mean())
