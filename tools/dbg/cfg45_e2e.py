"""Per-pass e2e times of configs 4 and 5 through the operator API (looking for sporadic stalls)."""
import os, sys, time, gc
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine, estimation
engine.init(0)
for name, regs in (("cfg4", synth.config4(seed=4, reads_per_locus=200)), ("cfg5", synth.config5(seed=5, n_reads=10000))):
    def fresh(): return [nrb.RepeatRegion.from_synth(r) for r in regs]
    print(name, "chunks:", [len(c) for c in estimation._chunks(fresh())])
    ts = []
    for i in range(8):
        rrs = fresh(); gc.collect(); gc.freeze()
        if i == 5: os.environ["NR_TRACE"] = "1"
        t0 = time.perf_counter(); nrb.estimate_regions(rrs, "ont", False); ts.append(time.perf_counter() - t0)
        os.environ.pop("NR_TRACE", None)
        gc.unfreeze()
    print(name, "e2e ms per pass:", [round(t * 1e3, 1) for t in ts])
