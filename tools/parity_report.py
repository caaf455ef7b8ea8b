#!/usr/bin/env python3
"""Writes the parity report (profiles/r2_parity.json by default): tests/parity_audit.py at full sample sizes on cuda:0.

    python tools/parity_report.py [--out profiles/r2_parity.json] [--configs c1,c2,c3,c4,c5] [--scale 5]

--scale multiplies the C3 / C4 / C5 sample sizes (default sizes: 1024 sets, 4096 members, 4096 grid points; C2 is always
the whole sweep).
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r2_parity.json"))
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    import gort_b200
    import parity_audit as pa
    g = gort_b200.Gort(0)
    sizes = dict(pa.DEFAULT_SIZES)
    if args.scale != 1.0:
        for k, cap in (("c3_sets", sizes["c3_total"]), ("c4_members", sizes["c4_total"]), ("c5_sets", 131072)):
            sizes[k] = int(min(cap, sizes[k] * args.scale))
    rep = pa.audit(g, sizes=sizes, configs=tuple(args.configs.split(",")), log=lambda *a: print(*a, file=sys.stderr, flush=True))
    g.close()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(rep, indent=1))
    print(json.dumps({"pass": rep["pass"], "seconds": rep["seconds"], "headline": pa.headline(rep)}))
    for f in pa.failures(rep):
        print("FAIL", f, file=sys.stderr)
    return 0 if rep["pass"] else 1


if __name__ == "__main__":
    sys.exit(main())
