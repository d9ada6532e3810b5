"""One resident pass over the config-5 sample (for ncu captures of the long-read kernels).  usage: cfg5_pass.py [n_reads]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from nanorepeat_b200 import synth, engine
engine.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
wl = bench.Workload("cfg5", synth.config5(seed=5, n_reads=n))
s = torch.cuda.Stream()
for _ in range(2):
    wl.resident_pass(s.cuda_stream)
torch.cuda.synchronize()
engine.set_timing(True)
wl.resident_pass(s.cuda_stream)
li = [b.launch_info() for b in (wl.b2, wl.b3)]
print("cfg5 round 2 kernel", round(li[0]["paired_ms"], 3), "ms; round 3 kernel", round(li[1]["paired_ms"], 3), "ms")
if os.environ.get("FULL"):
    print(li)
