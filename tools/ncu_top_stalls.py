#!/usr/bin/env python
"""Top instructions by stall samples from an `ncu --page source --csv` export, with the dominant stall reasons.
usage: ncu_top_stalls.py src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; data = rows[2:]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "not_issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]
print("total samples", tot, "instructions", len(data))
for r in top:
    reasons = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols if r[ix[c]] not in ("", "0")), reverse=True)[:3]
    print(f"{r[ix['Address']]:>8s} {100*int(r[ix['# Samples']])/tot:5.1f}% exec {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:70]:70s} {reasons}")
