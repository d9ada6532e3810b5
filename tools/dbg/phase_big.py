"""Phasing of regions with thousands of reads (amplicon data: config 2 has 5 000 reads per region) and odd inputs."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from nanorepeat_b200 import engine
engine.init(0)
rng = np.random.default_rng(3)
p = engine.GmmParams(error_rate=0.07, max_mutual_overlap=0.15, max_components=22, seed=5)
for n_reads in (30, 500, 5000, 50000):
    ks = np.where(rng.random(n_reads) < 0.5, 17, 55)
    sizes = list(np.round(ks + rng.normal(0, 0.03 * (10 + ks)), 2))
    engine.phase_1d(p, [sizes[:10]])
    t0 = time.perf_counter(); fit = engine.phase_1d(p, [sizes, sizes[::-1]]); dt = time.perf_counter() - t0
    print(n_reads, "reads x 2 regions:", round(dt * 1e3, 1), "ms; n =", fit[0]["n"], "means", np.round(fit[0]["means"], 2), "labels", np.bincount(fit[0]["label"][fit[0]["label"] >= 0]))
for name, sizes in (("all equal", [20.0] * 40), ("two reads", [10.0, 30.0]), ("one outlier", [20.0] * 40 + [500.0]), ("zeros", [0.0] * 25), ("wide", list(np.linspace(0, 300, 60)))):
    fit = engine.phase_1d(p, [sizes])[0]
    print(name, "-> n", fit["n"], "means", np.round(fit["means"], 2), "labels", np.bincount(fit["label"][fit["label"] >= 0]) if (fit["label"] >= 0).any() else [])
