#!/usr/bin/env python
"""Group the SASS lines of an `ncu --page source --csv` export by execution count (= code region of a persistent,
warp-specialised kernel) and print each region's instruction mix: python tools/ncu_regions.py src.csv npixels"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1]))); px = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter()
for r in body: cnt[r[ix['Instructions Executed']]] += 1
tot_s = sum(int(r[ix['# Samples']]) for r in body)
for gk in [k for k, v in sorted(cnt.items(), key=lambda kv: -int(kv[0]) * kv[1])[:10]]:
    c = collections.Counter(); smp = 0
    for r in body:
        if r[ix['Instructions Executed']] == gk:
            s = r[ix['Source']].split(); op = s[0] if not s[0].startswith('@') else s[1]
            c[op.split('.')[0] + ('.128' if '.128' in op else '')] += 1; smp += int(r[ix['# Samples']])
    print(f"exec {gk:>9s} x {cnt[gk]:4d} lines = {int(gk)*cnt[gk]/px:6.2f}/px  samples {100*smp/tot_s:5.1f}%  {dict(c.most_common(14))}")
