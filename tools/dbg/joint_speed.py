"""Device rate of the joint path on an HTT-like locus.  usage: joint_speed.py [n_reads]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from nanorepeat_b200 import synth, engine, joint
engine.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
loc = synth.joint_locus(seed=7, n_reads=n)
sc = engine.get_preset("ont")
ok1 = loc["range1"]; ok2 = loc["range2"]
s1 = joint.choose_best_step_size(3, ok1); s2 = joint.choose_best_step_size(3, ok2)
pr, p1, p2 = joint.round2_grid_points(ok1, ok2, min(a for a, _ in ok1), max(b for _, b in ok1), min(a for a, _ in ok2), max(b for _, b in ok2), s1, s2)
cells = sum(len(loc["reads"][int(r)]) * (2000 + 3 * int(k1) + 12 + 3 * int(k2)) * 2 for r, k1, k2 in zip(pr, p1, p2))
for _ in range(2):
    t0 = time.perf_counter(); rec, strand = engine.joint_grid(sc, loc["left"], loc["mid"], loc["right"], "CAG", "CCG", loc["reads"], pr, p1, p2); dt = time.perf_counter() - t0
print(f"{n} reads, steps {s1} {s2}, {len(pr)} grid points, {cells/1e9:.1f} Gcells, {dt*1e3:.1f} ms, {cells/dt/1e9:.0f} GCUPS (wall, incl. packing)")
t0 = time.perf_counter(); res = joint.quantify_two_repeats(loc["reads"], loc["left"], loc["mid"], loc["right"], "CAG", "CCG", ok1, ok2, 200, 50); dt = time.perf_counter() - t0
good = sum(abs(float(a) - t[0]) <= 1 and abs(float(b) - t[1]) <= 1 for a, b, t in zip(res["size1"], res["size2"], loc["truth"]) if a is not None)
print(f"quantify_two_repeats: {dt*1e3:.1f} ms, {good}/{n} reads within 1 unit of both simulated counts")
