"""Write the C2 sweep (BASELINE.json configs[1]) as a gortt angle file: header "N M W_1..W_M", then N lines."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gort_b200 import workloads as wk
w = wk.c2_hemisphere(wl_step=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ang, wl = w["angles"], w["wavelength"]
with open(sys.argv[1], "w") as f:
    f.write("%d %d %s\n" % (ang.shape[1], wl.size, " ".join("%g" % x for x in wl)))
    for k in range(ang.shape[1]):
        f.write("%g %g %g %g\n" % tuple(ang[:, k]))
