#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: SASS opcode mix and top stall lines.
usage: ncu -i rep --page source --csv --kernel-id :::1 | python tools/ncu_sass_mix.py [ntop]"""
import collections
import csv
import sys

ntop = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter(); samp = collections.Counter(); tot = stot = 0
for r in data:
    try:
        n = int(r[iE]); s = int(r[iSm])
    except ValueError:
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    ops[op] += n; tot += n; samp[op] += s; stot += s
print("kernel:", rows[0][1][:100] if rows and len(rows[0]) > 1 else "?")
print("total warp instructions", tot, " SASS lines", len(data), " samples", stot)
for op, n in ops.most_common(ntop):
    print("%-10s %12d %5.1f%%   stall samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(stot, 1)))
print("--- top sampled SASS lines (# samples, # executed, instruction)")
for r in sorted(data, key=lambda r: -int(r[iSm]) if r[iSm].isdigit() else 0)[:ntop]:
    print("%6s %10s  %s" % (r[iSm], r[iE], r[iS].strip()[:100]))
