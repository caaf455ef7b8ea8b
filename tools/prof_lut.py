"""ncu target: LUT generation for 2000 random parameter sets (all different shapes) + one C3-sized energy call."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gort_b200
from gort_b200 import workloads as wk
g = gort_b200.Gort(0)
rng = np.random.Generator(np.random.PCG64(3))
st = wk.random_structures(rng, 2000)
lut = g.lut(st)
w = wk.c3_albedo(n_sets=600)
l2 = g.lut(w["structure"])
rl, tl, rs = g.spectra(w["leaf"], w["soil"], w["wavelength"])
a, v, s = g.energy(w["structure"], l2, w["angles"], rl, tl, rs)
print("ok", np.isfinite(lut).mean(), np.isfinite(a).mean())
