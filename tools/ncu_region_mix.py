#!/usr/bin/env python
"""Instruction mix (weighted by executed count) of an address range of an `ncu --page source --csv` export.
usage: ncu_region_mix.py src.csv start_hex end_hex"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}; data = rows[2:]
base = int(data[0][ix["Address"]], 16); lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
mix = collections.Counter(); static = collections.Counter(); tot = 0
for r in data:
    a = int(r[ix["Address"]], 16) - base
    if lo <= a < hi:
        src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()); op = src.split()[0]
        key = op if op.startswith(("VIMNMX", "VIADDMNMX", "IMAD", "LDS", "SHFL", "STS", "LDG", "STG")) else op.split(".")[0]
        n = int(r[ix["Instructions Executed"]]); mix[key] += n; static[key] += 1; tot += n
print(f"range {lo:#x}..{hi:#x}: {tot} executed warp instructions")
for k, v in mix.most_common(18):
    print(f"  {k:16s} {100*v/tot:5.1f}%  (static {static[k]})")
