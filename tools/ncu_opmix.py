#!/usr/bin/env python
"""Opcode mix of an `ncu --page source --csv` export: python tools/ncu_opmix.py src.csv npixels"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
px = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); thr = collections.Counter(); wf = collections.Counter(); smp = collections.Counter()
for r in body:
    s = r[ix['Source']].split()
    op = s[0] if not s[0].startswith('@') else s[1]
    op = op.split('.')[0] if not op.startswith(('LDS', 'STS', 'LDG', 'I2F', 'F2I')) else op
    n = float(r[ix['Instructions Executed']])
    ops[op] += n; thr[op] += float(r[ix['Thread Instructions Executed']])
    smp[op] += float(r[ix['# Samples']])
    try: wf[op] += float(r[ix['L1 Wavefronts Shared']])
    except Exception: pass
tot = sum(ops.values()); ts = sum(smp.values())
print(f"total warp instr {tot/1e6:.1f}M = {tot/px:.2f}/px")
for k, v in ops.most_common(32):
    print(f"{k:22s} {v/1e6:8.1f}M {100*v/tot:5.1f}%  /px {v/px:6.2f}  thr {thr[k]/max(v,1):5.1f}  smem wf/px {wf[k]/px:6.2f}  samples {100*smp[k]/ts:5.1f}%")
