#!/usr/bin/env python3
"""Development check (GPU): device time of the gap-probability kernels on the LUT workloads of BASELINE.json
(C3 10^4 sets, C4b 10^5 members all different, C4a 10^5 members LAI only, C5 131 072 grid points)."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import gort_b200
    from gort_b200 import workloads as wk
    from gort_b200.api import LUT_STRIDE
    dev = torch.device("cuda:0")
    g = gort_b200.Gort(0)
    ts = torch.cuda.Stream(device=dev)
    res = {}
    for name, st in (("c3", wk.c3_albedo()["structure"]), ("c4b", wk.c4_enkf()["structure"]),
                     ("c4a", wk.c4_enkf(vary_structure=False)["structure"]), ("c5", wk.c5_lut_grid()["structure"])):
        d_st = torch.from_numpy(np.ascontiguousarray(st)).to(dev)
        out = torch.empty((st.shape[1], LUT_STRIDE), dtype=torch.float64, device=dev)
        g.lut_dev(d_st, out, stream=ts.cuda_stream); ts.synchronize()
        best = 1e9
        for _ in range(3):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(ts); g.lut_dev(d_st, out, stream=ts.cuda_stream); b.record(ts); b.synchronize()
            best = min(best, a.elapsed_time(b))
        res[name + "_ms"] = round(best, 3)
        h = out.cpu().numpy()
        res[name + "_nan_sets"] = int(np.isnan(h).any(axis=1).sum())
        res[name + "_sum"] = float(np.nansum(h))
    print(json.dumps(res))
    g.close()


if __name__ == "__main__":
    main()
