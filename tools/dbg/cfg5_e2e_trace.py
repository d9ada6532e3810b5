"""Host phases of estimate_regions on the config-5 sample (NR_TRACE).  usage: cfg5_e2e_trace.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine
engine.init(0)
regs = synth.config5(seed=5, n_reads=10000)
for _ in range(2): nrb.estimate_regions([nrb.RepeatRegion.from_synth(r) for r in regs], "ont", False)
ts = []
for _ in range(3):
    rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]; t0 = time.perf_counter(); nrb.estimate_regions(rrs, "ont", False); ts.append(time.perf_counter() - t0)
print("cfg5 e2e ms", [round(t * 1e3, 1) for t in ts], flush=True)
os.environ["NR_TRACE"] = "1"
nrb.estimate_regions([nrb.RepeatRegion.from_synth(r) for r in regs], "ont", False)
