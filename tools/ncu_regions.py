#!/usr/bin/env python3
"""Stall samples and executed instructions per block of N SASS lines of one kernel.
usage: ncu -i rep --page source --csv --kernel-name regex:K | python tools/ncu_regions.py [N]"""
import csv, sys
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rows = list(csv.reader(sys.stdin))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
seen = set(); d2 = []
for r in rows[h + 1:]:
    if len(r) != len(hdr) or r[0] in seen or not r[hdr.index('# Samples')].isdigit():
        continue
    seen.add(r[0]); d2.append(r)
iS = hdr.index('# Samples'); iE = hdr.index('Instructions Executed'); iSrc = hdr.index('Source')
stalls = [x for x in hdr if x.startswith('stall_') and 'Not Issued' not in x]
tot = sum(int(r[iS]) for r in d2)
print('SASS lines', len(d2), 'samples', tot, 'warp instructions', sum(int(r[iE]) for r in d2))
for b in range(0, len(d2), B):
    blk = d2[b:b + B]
    s = sum(int(r[iS]) for r in blk); e = sum(int(r[iE]) for r in blk)
    st = {x: sum(int(r[hdr.index(x)]) for r in blk) for x in stalls}
    top = ' '.join('%s=%d' % (k[6:], v) for k, v in sorted(st.items(), key=lambda x: -x[1])[:4] if v)
    ops = {}
    for r in blk:
        t = r[iSrc].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] = ops.get(op, 0) + 1
    topo = ' '.join('%s:%d' % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:4])
    print('%5d-%5d samples %5d (%4.1f%%) exec %9d | %s | %s' % (b, b + B, s, 100 * s / max(tot, 1), e, top, topo))
