"""e2e of estimate_regions against the Python chunk size.  usage: chunk_sweep.py cfg2|cfg3"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine, estimation
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
regs = synth.config2(seed=2, n_reads=5000) if which == "cfg2" else synth.config3(seed=3, n_loci=2000)
engine.init(0)
for cm in (4096, 8192, 16384, 10**9):
    estimation.CHUNK_MIN_READS = cm
    for _ in range(2): nrb.estimate_regions([nrb.RepeatRegion.from_synth(r) for r in regs], "ont", False)
    ts = []
    for _ in range(5):
        rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]; t0 = time.perf_counter(); nrb.estimate_regions(rrs, "ont", False); ts.append(time.perf_counter() - t0)
    print(which, "chunk_min", cm, "e2e ms", round(min(ts) * 1e3, 2), round(sorted(ts)[2] * 1e3, 2), flush=True)
