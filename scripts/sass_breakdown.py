#!/usr/bin/env python
"""Condense the source page of an ncu capture (`ncu -i rep --page source --csv`) into a per-kernel table:
share of the warp-stall samples and of the executed warp-instructions per SASS opcode, and the stall reasons.

    python scripts/sass_breakdown.py gpurun_out/prof_<tag>_src.csv > profiles/<tag>_sass_breakdown.md
"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
names = [rows[i][1] for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seen = set()
print("# SASS-level breakdown of the captured kernels (`%s`)\n" % sys.argv[1].split('/')[-1])
for k, hi in enumerate(heads):
    name = names[k] if k < len(names) else "kernel %d" % k
    if name in seen:
        continue
    seen.add(name)
    end = heads[k + 1] - 2 if k + 1 < len(heads) else len(rows)
    h = rows[hi]
    data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
    ix = {c: i for i, c in enumerate(h)}
    S = [int(r[ix['# Samples']] or 0) for r in data]
    I = [int(r[ix['Instructions Executed']] or 0) for r in data]
    ts, ti = sum(S) or 1, sum(I) or 1
    ops_s, ops_i = collections.Counter(), collections.Counter()
    for r, s, i in zip(data, S, I):
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ix['Source']])
        o = m.group(2) if m else '?'
        ops_s[o] += s
        ops_i[o] += i
    print("## `%s`\n" % name[:100])
    print("%d SASS instructions, %.1f M warp-instructions executed, %d stall samples\n" % (len(data), ti / 1e6, ts))
    print("| opcode | stall samples | warp-instructions |\n|---|---:|---:|")
    for o, c in ops_s.most_common(16):
        print("| `%s` | %.1f %% | %.1f %% (%.1f M) |" % (o, 100.0 * c / ts, 100.0 * ops_i[o] / ti, ops_i[o] / 1e6))
    st = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    tot = {c: sum(int(r[ix[c]] or 0) for r in data) for c in st}
    t = sum(tot.values()) or 1
    print("\nstall reasons (all samples): " + ", ".join("%s %.0f %%" % (c[6:], 100.0 * v / t)
                                                        for c, v in sorted(tot.items(), key=lambda x: -x[1])[:8]) + "\n")
