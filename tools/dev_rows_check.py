#!/usr/bin/env python3
"""Development check (GPU): the full-spectrum rows kernel against the chunked wide kernel on the C2 sweep -- same bits --
and their device times, isolated (one call, synchronised) and back to back with and without cross-call overlap.
The experimental kernel is selected with GORT_ROWS=1 in a child process (the switch is read by gort_create)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run(tag):
    import torch
    import gort_b200
    from gort_b200 import workloads as wk
    dev = torch.device("cuda:0")
    g = gort_b200.Gort(0)
    res = {"tag": tag}
    for name, sets in (("c2", 1), ("c2x3", 3)):
        w = wk.c2_hemisphere(sets=sets)
        st, ang, wl = w["structure"], w["angles"], w["wavelength"]
        if sets > 1:
            ang = np.ascontiguousarray(ang[:, :3000])
        G, W = ang.shape[1], wl.shape[0]
        lut = g.lut(st)
        rl, tl, rs = g.spectra(w["leaf"], w["soil"], wl)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d = [T(st), T(lut), T(ang), T(rl), T(tl), T(rs)]
        out = torch.full((sets, G, 2112), -5.0, dtype=torch.float64, device=dev)
        ts = torch.cuda.Stream(device=dev)
        call = g.brdf_dev_bind(*d, out, stream=ts.cuda_stream)
        call(); ts.synchronize(); g.synchronize()
        h = out.cpu().numpy()
        res[name + "_sha"] = __import__("hashlib").sha1(h[:, :, :2101].tobytes()).hexdigest()
        res[name + "_pad_ok"] = bool(np.array_equal(h[:, :, 2101:], np.repeat(h[:, :, 2100:2101], 11, axis=2)))
        res[name + "_finite"] = bool(np.isfinite(h).all())
        if sets == 1:
            def timeit(n, sync_each):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                if sync_each:
                    tot = 0.0
                    for _ in range(n):
                        a.record(ts); call(); b.record(ts); b.synchronize(); tot += a.elapsed_time(b)
                    return tot / n * 1e3
                a.record(ts)
                for _ in range(n):
                    call()
                b.record(ts); b.synchronize()
                return a.elapsed_time(b) / n * 1e3
            for _ in range(5):
                call()
            ts.synchronize()
            res["us_isolated_call"] = timeit(30, True)
            res["us_back_to_back"] = timeit(100, False)
            g.set_overlap(True)
            for _ in range(5):
                call()
            res["us_back_to_back_overlap"] = timeit(100, False)
            g.set_overlap(False)
            ts.synchronize()
            g.profile_begin(30)
            for _ in range(30):
                call()
            ts.synchronize()
            gm, rm, n = g.profile_end()
            res["us_geom_kernel"] = gm * 1e3; res["us_per_wavelength_kernel"] = rm * 1e3
            g.kernel_stamps_enable(True)
            for _ in range(10):
                call(); ts.synchronize()
            res["stamps"] = g.kernel_stamps()
            g.kernel_stamps_enable(False)
            np.save(ROOT / "gpurun_out" / ("rows_check_%s.npy" % tag), h[0, ::97, :2101])
    g.synchronize()
    g.close()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        env = dict(os.environ)
        env.pop("GORT_ROWS", None)
        subprocess.run([sys.executable, __file__, "wide"], env=env, check=True)
        env["GORT_ROWS"] = "1"
        subprocess.run([sys.executable, __file__, "rows"], env=env, check=True)
        a = np.load(ROOT / "gpurun_out" / "rows_check_rows.npy"); b = np.load(ROOT / "gpurun_out" / "rows_check_wide.npy")
        print("rows == wide bits on the sampled lines:", bool(np.array_equal(a, b)), "max abs diff", float(np.max(np.abs(a - b))))
