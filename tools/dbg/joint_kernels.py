"""Device time per kernel of one nr_joint_grid call (round-2 grid of the HTT-like locus).  usage: joint_kernels.py [n_reads]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from nanorepeat_b200 import synth, engine, joint
engine.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
loc = synth.joint_locus(seed=7, n_reads=n)
sc = engine.get_preset("ont")
ok1, ok2 = loc["range1"], loc["range2"]
pr, p1, p2 = joint.round2_grid_points(ok1, ok2, min(a for a, _ in ok1), max(b for _, b in ok1), min(a for a, _ in ok2), max(b for _, b in ok2), 3, 2)
args = (sc, loc["left"], loc["mid"], loc["right"], "CAG", "CCG", loc["reads"], pr, p1, p2)
engine.joint_grid(*args)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    engine.joint_grid(*args)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "kernel" in e.key or "Memcpy" in e.key:
        print(f"{e.key[:70]:70s} x{e.count}  {e.device_time_total / 1e3:8.3f} ms")
