#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: total stall samples by reason and the
hottest SASS instructions.  usage: stall_summary.py src.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stalls}
all_s = sum(tot.values())
print(rows[0][1][:100], "| SASS instructions:", len(body), "| samples:", all_s)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {k:28s} {v:9d} {100*v/max(all_s,1):5.1f}%")
si = hdr.index("# Samples")
ie = hdr.index("Instructions Executed")
print("  total warp instructions executed:", sum(int(r[ie] or 0) for r in body))
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:top]:
    why = max(stalls, key=lambda i: int(r[i] or 0))
    print(f"  {int(r[si]):7d} {hdr[why]:22s} {r[1].strip()[:90]}")
