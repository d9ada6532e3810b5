"""A small case that touches every kernel path (pairs, resumed round 3, redo list, long reads on concurrent stripes,
32-bit exact and ladder kernels) for compute-sanitizer runs."""
import sys
sys.path.insert(0, ".")
import numpy as np
import nanorepeat_b200 as nrb
from nanorepeat_b200 import engine, synth
engine.init(0)
rng = np.random.default_rng(3)
regs = synth.config1(seed=5, n_regions=2, reads_per_region=9)
L, R = synth.random_seq(rng, 300), synth.random_seq(rng, 300)
long_reg = synth.make_region(rng, "long", "GGGGCC", [90, 120], 6, "ont") if hasattr(synth, "make_region") else None
rrs = [nrb.RepeatRegion.from_synth(r) for r in regs] + ([nrb.RepeatRegion.from_synth(long_reg)] if long_reg else [])
nrb.estimate_regions(rrs, "ont", False)
n = sum(rd.round3_repeat_size is not None for rr in rrs for rd in rr.read_dict.values())
# undecidable ties -> redo list
sc = engine.get_preset("ont"); sc.min_dp_score = 1
left, right = synth.random_seq(rng, 50), synth.random_seq(rng, 60)
cores = ["CAG" * 5 + right[:20], left[-1:] + "CAG" * 4 + right[:1], left[-30:] + "CAG" * 6 + right[:40], "CAG" * 3]
s = engine.round3_region(sc, left, right, "CAG", cores, np.array([2, 1, 3, 0], np.int32), np.array([8, 7, 9, 6], np.int32))
# 32-bit kernels with coordinates, multi-stripe
q = synth.random_seq(rng, 900)
a = engine.score_tasks([q, q[:100]], [synth.random_seq(rng, 50) + q + synth.random_seq(rng, 50), q[:120]], engine.get_preset("ont"))
print("ok", n, s[0].tolist(), a.tolist())
