#!/usr/bin/env python
"""Host timeline of one pipelined estimate_regions() call on bench.py's workload (needs a GPU)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, estimation as est, engine
regs = synth.config2(seed=2, n_reads=5000)
def fresh(): return [nrb.RepeatRegion.from_synth(r) for r in regs]
for _ in range(3): nrb.estimate_regions(fresh(), "ont", False)
rrs = fresh()
T = []
def mark(name): T.append((name, time.perf_counter()))
mark("start")
groups = est._pipeline_groups(rrs)
r2 = []
for g in groups:
    r2.append(est._round2_launch("ont", g)); mark("r2 launch")
r3 = []
for g, ctx in zip(groups, r2):
    est._round2_finish(ctx); mark("r2 finish")
    r3.append(est._round3_reuse_launch(False, g[0]._nr_round2[0], g)); mark("r3 launch")
for ctx in r3:
    est._round3_reuse_finish(ctx); mark("r3 finish")
t0 = T[0][1]
for (n, t), (_, p) in zip(T[1:], T[:-1]):
    print(f"{n:12s} +{(t - p) * 1e3:6.2f} ms   at {(t - t0) * 1e3:6.2f}")
# finer: pieces of one r2 launch
rrs = fresh()
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); nrb.estimate_regions(rrs, "ont", False); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
