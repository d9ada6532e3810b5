#!/usr/bin/env python
"""Compile nr_api.cu with -Xptxas -v and print one line per kernel: registers, spills, smem."""
import re, subprocess, sys, os
here = os.path.dirname(os.path.abspath(__file__))
src = os.path.join(here, "..", "nanorepeat_b200", "csrc", "nr_api.cu")
out = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                      "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-shared", "-o", "/tmp/_ptxas_report.so", src]
                     + sys.argv[1:], capture_output=True, text=True).stderr
cur = None
rows = {}
for ln in out.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout
        mm = re.search(r"nr::(\w+)<(true|false)>", d)
        cur = (mm.group(1), mm.group(2) == "true", 0) if mm else (m.group(1), False, 0)
        rows[cur] = {}
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and cur:
        rows[cur]["stack"], rows[cur]["spill"] = int(m.group(1)), int(m.group(2))
    m = re.search(r"Used (\d+) registers", ln)
    if m and cur:
        rows[cur]["regs"] = int(m.group(1))
for k in sorted(rows):
    print(f"{k[0]:14s} multi={int(k[1])} R={k[2]:2d} regs={rows[k].get('regs')} stack={rows[k].get('stack')} spill={rows[k].get('spill')}")
