#!/usr/bin/env python
"""Innermost loops of a kernel from an `ncu --page source --csv` export: share of executed warp instructions and of stall
samples, plus a few marker counts (static) to recognise the loop.  usage: ncu_loops.py src.csv [min_share_pct]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}; data = rows[2:]
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
base = int(data[0][ix["Address"]], 16)
ins = [(int(r[ix["Address"]], 16) - base, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])) for r in data]
idx = {a: i for i, (a, _, _, _) in enumerate(ins)}
tot_i = sum(x[2] for x in ins); tot_s = sum(x[3] for x in ins)
loops = []
for i, (a, s, n, sm) in enumerate(ins):
    m = re.search(r"\bBRA\S*\s+.*?(0x[0-9a-f]+)", s)
    if m:
        t = int(m.group(1), 16) - (0 if int(m.group(1), 16) < base else base)
        if t in idx and t <= a:
            loops.append((idx[t], i))
# innermost: loops that contain no other loop
inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
covered = 0
print(f"total warp instructions {tot_i}, samples {tot_s}")
for lo, hi in sorted(inner):
    body = ins[lo:hi + 1]
    ni = sum(x[2] for x in body); ns = sum(x[3] for x in body)
    covered += ni
    if 100 * ni / tot_i < minshare: continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
    dpx = ops["VIADDMNMX"] + ops["VIMNMX3"]
    iters = max(x[2] for x in body)
    print(f"loop {ins[lo][0]:#8x}..{ins[hi][0]:#8x} len {hi-lo+1:5d}  instr {100*ni/tot_i:5.1f}%  samples {100*ns/tot_s:5.1f}%  clk/instr-ish {ns/tot_s/(ni/tot_i):4.2f}  "
          f"cells~{dpx/6:.0f} SHFL {ops['SHFL']} VOTE {ops['VOTE']} LDS {ops['LDS']} STS {ops['STS']} STG {ops['STG']} ISETP {ops['ISETP']}")
print(f"innermost loops cover {100*covered/tot_i:.1f}% of executed instructions")
