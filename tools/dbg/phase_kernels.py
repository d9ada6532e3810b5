"""One nr_phase_1d call over 500 loci x 30 sizes (for an ncu capture of the mixture-fit kernel)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from nanorepeat_b200 import engine
engine.init(0)
rng = np.random.default_rng(11)
loci = []
for g in range(500):
    a = int(rng.integers(5, 120)); b = a + int(rng.integers(6, 60))
    ks = np.where(rng.random(30) < 0.5, a, b)
    loci.append(list(np.round(ks + rng.normal(0, 0.01 * (10 + ks)), 2)))
p = engine.GmmParams(error_rate=0.07, max_mutual_overlap=0.15, max_components=22, seed=5)
fits = engine.phase_1d(p, loci)
print("components:", np.bincount([f["n"] for f in fits]))
