#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their instruction mix.
usage: sass_loops.py sass.txt <function-substring> [min_len]"""
import re, sys, collections
txt = open(sys.argv[1]).read().split("Function : ")
fn = [t for t in txt if sys.argv[2] in t.split("\n")[0]][0]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ins = []
for ln in fn.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
for i, (a, s) in enumerate(ins):
    m = re.search(r"\bBRA\S*\s+.*?(0x[0-9a-f]+)", s)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_idx and i - addr_idx[tgt] >= minlen:
            body = ins[addr_idx[tgt]:i + 1]
            h = collections.Counter()
            for _, b in body:
                b = re.sub(r"^@!?U?P\d+\s+", "", b)
                op = b.split()[0]
                h[op.split(".")[0] if not op.startswith(("VIMNMX", "VIADDMNMX", "IMAD", "LDS", "SHFL")) else op] += 1
            print(f"loop {tgt:#x}..{a:#x} len {len(body)}: " + ", ".join(f"{k}:{v}" for k, v in h.most_common()))
