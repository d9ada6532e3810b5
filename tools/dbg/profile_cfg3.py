import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth
regs = synth.config3(seed=3, n_loci=2000)
def step():
    rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
    t0 = time.perf_counter(); nrb.estimate_regions(rrs, "hifi", False); return time.perf_counter() - t0
for _ in range(2): step()
print("cfg3 2000 loci e2e ms", min(step() for _ in range(3)) * 1e3)
rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
pr = cProfile.Profile(); pr.enable(); nrb.estimate_regions(rrs, "hifi", False); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
