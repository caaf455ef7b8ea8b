#!/usr/bin/env python3
"""Writes the parity report (profiles/r2_parity.json by default): tests/parity_audit.py at full sample sizes on cuda:0.

    python tools/parity_report.py [--out profiles/r2_parity.json] [--configs c1,c2,c3,c4,c5]
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
    args = ap.parse_args()
    import gort_b200
    import parity_audit as pa
    g = gort_b200.Gort(0)
    rep = pa.audit(g, configs=tuple(args.configs.split(",")), log=lambda *a: print(*a, file=sys.stderr, flush=True))
    g.close()
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(rep, indent=1))
    print(json.dumps({"pass": rep["pass"], "seconds": rep["seconds"], "headline": pa.headline(rep)}))
    for f in pa.failures(rep):
        print("FAIL", f, file=sys.stderr)
    return 0 if rep["pass"] else 1


if __name__ == "__main__":
    sys.exit(main())
