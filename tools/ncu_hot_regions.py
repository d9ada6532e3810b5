#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export into code regions: samples, executed warp instructions and the main
stall reasons per contiguous hot region (loops).  usage: ncu_hot_regions.py src.csv [n_regions] [gap]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
base = int(data[0][ix["Address"]], 16)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(int(r[ix["# Samples"]]) for r in data); tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"total samples {tot_s}, warp instructions {tot_i}")
# split into regions at backward-branch targets: simpler -- fixed windows merged by activity
W = 64
wins = {}
for r in data:
    a = (int(r[ix["Address"]], 16) - base) // 16
    w = wins.setdefault(a // W, [0, 0, {s: 0 for s in stalls}, a])
    w[0] += int(r[ix["# Samples"]]); w[1] += int(r[ix["Instructions Executed"]])
    for s in stalls:
        w[2][s] += int(r[ix[s]])
# merge adjacent windows with similar per-instruction execution counts (same loop)
keys = sorted(wins)
regions = []
for k in keys:
    w = wins[k]
    if regions and k == regions[-1]["end"] + 1 and w[1] > 0 and regions[-1]["instr"] > 0 and 0.5 < (w[1] / W) / (regions[-1]["instr"] / ((regions[-1]["end"] - regions[-1]["start"] + 1) * W)) < 2.0:
        g = regions[-1]; g["end"] = k; g["samples"] += w[0]; g["instr"] += w[1]
        for s in stalls: g["st"][s] += w[2][s]
    else:
        regions.append({"start": k, "end": k, "samples": w[0], "instr": w[1], "st": dict(w[2])})
regions.sort(key=lambda g: -g["samples"])
for g in regions[:topn]:
    top = sorted(g["st"].items(), key=lambda kv: -kv[1])[:5]
    print(f"addr {g['start']*W*16:#8x}..{(g['end']+1)*W*16:#8x}  samples {100*g['samples']/tot_s:5.1f}%  instr {100*g['instr']/tot_i:5.1f}%  "
          + ", ".join(f"{k[6:]}:{100*v/max(1,g['samples']):.0f}%" for k, v in top))
