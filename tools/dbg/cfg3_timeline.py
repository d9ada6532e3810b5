"""Absolute timeline of one config-3-slice pass through the operator API (library phases + Python phases)."""
import os, sys, time, gc
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine, estimation
regs = synth.config3(seed=3, n_loci=2000)
engine.init(0)
def fresh(): return [nrb.RepeatRegion.from_synth(r) for r in regs]
for _ in range(3): nrb.estimate_regions(fresh(), "hifi", False)
rrs = fresh(); gc.collect(); gc.freeze()
os.environ["NR_TRACE"] = "1"
marks = []
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t = time.monotonic_ns(); r = f(*a, **k); marks.append((name, t, time.monotonic_ns())); return r
    setattr(mod, name, g); return f
for n in ("_gather_chunk", "_run_chunk", "_assign_chunk", "_chunks"): wrap(estimation, n)
wrap(engine, "estimate_regions")
t0 = time.monotonic_ns(); nrb.estimate_regions(rrs, "hifi", False); t1 = time.monotonic_ns()
us = lambda t: (t / 1e3) % 1e8
print(f"=== pass: {(t1 - t0) / 1e6:.2f} ms, from {us(t0):.1f} to {us(t1):.1f} us", file=sys.stderr)
for name, a, b in sorted(marks, key=lambda x: x[1]):
    print(f"    py {name:18s} {us(a):12.1f} -> {us(b):12.1f}  ({(b - a) / 1e3:7.1f} us)", file=sys.stderr)
