#!/usr/bin/env python
"""Where the end-to-end time of estimate_regions goes (host side): cProfile of the Python layer + the library's own phase
trace (NR_TRACE=1) for config 2 and the config-3 slice.  Needs a GPU.  usage: e2e_profile.py [cfg2|cfg3]"""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
regs = synth.config2(seed=2, n_reads=5000) if which == "cfg2" else synth.config3(seed=3, n_loci=2000)
engine.init(0)
def fresh(): return [nrb.RepeatRegion.from_synth(r) for r in regs]
for _ in range(3): nrb.estimate_regions(fresh(), "ont", False)
ts = []
for _ in range(5):
    rrs = fresh(); t0 = time.perf_counter(); nrb.estimate_regions(rrs, "ont", False); ts.append(time.perf_counter() - t0)
print(which, "e2e ms per pass:", [round(t * 1e3, 2) for t in ts])
os.environ["NR_TRACE"] = "1"          # read by the library at each call
rrs = fresh(); nrb.estimate_regions(rrs, "ont", False)
os.environ.pop("NR_TRACE")
pr = cProfile.Profile(); rrs = fresh(); pr.enable(); nrb.estimate_regions(rrs, "ont", False); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:5000])
