"""Resident config-2 pass: per-kernel event times over a few steps (quick look at round 2 / round 3 kernel times)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from nanorepeat_b200 import synth, engine
engine.init(0)
wl = bench.Workload("cfg2", synth.config2(seed=2, n_reads=5000))
stream = torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    wl.resident_pass(stream.cuda_stream)
torch.cuda.synchronize()
for rep in range(3):
    dev_ms, kern_ms = bench.time_resident(wl, torch, stream, flush, 4, 10, engine)
    print("ms per pass", round(sum(dev_ms) / 40, 4), "round 2 kernel", round(kern_ms[0]["paired_ms"] / 40, 4), "round 3 kernel", round(kern_ms[1]["paired_ms"] / 40, 4))
