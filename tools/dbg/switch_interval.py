import sys, time
sys.path.insert(0, ".")
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine, estimation
for name, regs in (("cfg2", synth.config2(seed=2, n_reads=5000)), ("cfg3", synth.config3(seed=3, n_loci=2000))):
    for sw in (0.005, 0.0005, 0.00005):
        sys.setswitchinterval(sw)
        for cm in (4096, 10**9):
            estimation.CHUNK_MIN_READS = cm
            for _ in range(2): nrb.estimate_regions([nrb.RepeatRegion.from_synth(r) for r in regs], "ont", False)
            ts=[]
            for _ in range(5):
                rrs=[nrb.RepeatRegion.from_synth(r) for r in regs]; t0=time.perf_counter(); nrb.estimate_regions(rrs, "ont", False); ts.append(time.perf_counter()-t0)
            print(name, "switch", sw, "chunk_min", cm, "e2e ms", round(min(ts)*1e3,2), round(sorted(ts)[2]*1e3,2))
