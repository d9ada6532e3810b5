#!/usr/bin/env python
"""Warp-level latency vs. occupancy of the exact kernel: n identical-shape tasks (one per warp), n chosen so that each
SMSP holds 1, 2, 3 or 4 warps.  Prints clocks per column step per warp and executed GCUPS (needs a GPU)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from nanorepeat_b200 import engine
rng = np.random.default_rng(0)
sc = engine.get_preset("ont")
engine.init(0)
info = engine.device_info()
sms, mhz = info["sm_count"], info["clock_khz"] / 1e3
st = torch.cuda.Stream()
for q, t in ((256, 4000), (384, 4000), (512, 4000), (650, 4000)):
    for wps in (1, 2, 3, 4):
        n = sms * 4 * wps
        qs = ["".join(rng.choice(list("ACGT"), q)) for _ in range(8)]
        tpl = "".join(rng.choice(list("ACGT"), t))
        b = engine.Batch.tasks(sc, [qs[i % 8] for i in range(n)], [tpl] * n)
        for _ in range(2): b.run(st.cuda_stream)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); b.run(st.cuda_stream); e1.record(st); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = min(ts); s = b.stats()
        stripes = 1 if q <= 512 else -(-q // 512)
        steps = stripes * (t + 31)
        print(f"q={q} t={t} warps/SMSP={wps} ms={ms:.3f} clk/step/warp={ms*1e-3*mhz*1e6/steps:.0f} "
              f"executed GCUPS={s['executed_cells']/ms/1e6:.0f}", flush=True)
