"""Per-pass e2e times inside bench.py's own flow (Workload with resident batches alive, gc frozen) for config 3 / 5 / 2."""
import gc, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
from nanorepeat_b200 import synth, engine
engine.init(0)
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
regs = {"cfg3": lambda: synth.config3(seed=3, n_loci=2000), "cfg5": lambda: synth.config5(seed=5, n_reads=10000),
        "cfg2": lambda: synth.config2(seed=2, n_reads=5000)}[which]()
wl = bench.Workload(which, regs)
stream = torch.cuda.Stream()
for _ in range(3):
    wl.resident_pass(stream.cuda_stream)
torch.cuda.synchronize()
ts = []
for i in range(12):
    rrs = wl.fresh(); gc.collect(); gc.freeze()
    if i == 0 and os.environ.get("TRACE_FIRST"): os.environ["NR_TRACE"] = "1"
    t0 = time.perf_counter(); wl.e2e_pass(rrs); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    os.environ.pop("NR_TRACE", None)
    gc.unfreeze()
print(which, "bench-flow e2e ms per pass:", [round(t * 1e3, 1) for t in ts])
ts = []
for i in range(6):
    rrs = wl.fresh()
    t0 = time.perf_counter(); wl.e2e_pass(rrs); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(which, "no gc calls:", [round(t * 1e3, 1) for t in ts])
