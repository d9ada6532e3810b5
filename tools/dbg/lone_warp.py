import os, sys, random
sys.path.insert(0, ".")
import torch
from nanorepeat_b200 import engine
engine.init(0)
sc = engine.get_preset("ont")
rng = random.Random(1)
def rs(n): return "".join(rng.choice("ACGT") for _ in range(n))
stream = torch.cuda.Stream()
T = 4000
for q in (128, 384, 512, 600, 1024, 2048, 4096):
    qs, ts = [rs(q)], [rs(T)]
    b = engine.Batch.tasks(sc, qs, ts)
    for _ in range(3): b.run(stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); b.run(stream.cuda_stream); e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1)
    li = b.launch_info()
    print("cap", os.environ.get("NR_COOP_ROWS"), "q", q, "t", T, "entries", li["n_rest"], "ms", round(ms, 3), "clk per column", round(ms * 1e-3 * 1.965e9 / T))
    b.close()
