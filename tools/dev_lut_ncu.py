#!/usr/bin/env python3
"""Development: one LUT call over the first N members of C4b (every member its own crown shape), for ncu."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import gort_b200
from gort_b200 import workloads as wk
from gort_b200.api import LUT_STRIDE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = gort_b200.Gort(0)
st = np.ascontiguousarray(wk.c4_enkf(n_members=n)["structure"])
dev = torch.device("cuda:0")
d = torch.from_numpy(st).to(dev)
out = torch.empty((n, LUT_STRIDE), dtype=torch.float64, device=dev)
for _ in range(2):
    g.lut_dev(d, out)
g.synchronize()
print("ok", float(torch.nansum(out)))
