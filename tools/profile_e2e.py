#!/usr/bin/env python
"""cProfile of one operator-API step of bench.py's workload (host-side cost breakdown).  Needs a GPU."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth
regs = synth.config2(seed=2, n_reads=int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
def step():
    rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
    nrb.estimate_regions(rrs, "ont", False)
for _ in range(3):
    step()
t0 = time.perf_counter(); step(); print("step ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
