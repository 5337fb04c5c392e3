#!/usr/bin/env python
"""Segments of a kernel by execution count (ncu --page source --csv export): python scripts/ncu_segments.py <src.csv> [min_pct]"""
import csv, re, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]; ix = {c: i for i, c in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
seg = []; cur = None
for r in data:
    n = int(r[ix['Instructions Executed']] or 0)
    src = r[ix['Source']].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src); op = m.group(2) if m else '?'
    smp = int(r[ix['# Samples']] or 0)
    if cur and abs(n - cur[0]) <= 0.03 * max(cur[0], 1) + 5:
        cur[1] += 1; cur[2] += n; cur[3][op.split('.')[0]] += 1; cur[5] += smp
    else:
        if cur: seg.append(cur)
        cur = [n, 1, n, Counter({op.split('.')[0]: 1}), r[ix['Address']], smp]
seg.append(cur)
ts = sum(x[5] for x in seg) or 1
print("total warp-instructions %d, SASS instructions %d, samples %d" % (tot, len(data), ts))
for n, k, t, ops, addr, smp in seg:
    if 100.0 * t / tot > thr or 100.0 * smp / ts > thr:
        print(addr[-5:], "count~%d x %d instr = %.1f%% instr, %.1f%% samples" % (n, k, 100.0 * t / tot, 100.0 * smp / ts), dict(ops.most_common(5)))
