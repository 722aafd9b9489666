#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: top SASS lines by stall samples / instructions."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in body); tot_i = sum(f(r, "Instructions Executed") for r in body)
print(f"total samples {tot_s:.0f}, total warp instr {tot_i:.0f}, SASS lines {len(body)}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in body) for s in stalls}
print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
print("--- top by samples")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:topn]:
    top = sorted(stalls, key=lambda s: -f(r, s))[:2]
    print(f"{f(r,'# Samples'):7.0f} {100*f(r,'# Samples')/tot_s:5.1f}%  inst {f(r,'Instructions Executed'):10.0f}  {r[ix['Source']][:70]:70s} {top[0]}={f(r,top[0]):.0f} {top[1]}={f(r,top[1]):.0f} bank={r[ix['L1 Conflicts Shared N-Way']]}")
