#!/usr/bin/env python
"""Device time of the resident round-2 and round-3 batches of bench.py's workload, separately (needs a GPU)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import nanorepeat_b200 as nrb
from nanorepeat_b200 import engine, synth
from nanorepeat_b200.estimation import ladder_bounds_array
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
regs = synth.config2(seed=2, n_reads=n)
sc = engine.get_preset("ont")
rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
nrb.estimate_regions(rrs, "ont", False)
b2 = engine.Batch.begin(sc, "round2"); b3 = engine.Batch.begin(sc, "round3")
for reg, rr in zip(regs, rrs):
    m = len(reg.repeat_unit_seq)
    r1 = max(d / m for d in reg.dist_between_anchors); T = int(r1 * 1.5) + 1
    if T < r1 + 10: T = int(r1 + 10)
    b2.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, T, reg.core_seqs)
    ok = [i for i, nme in enumerate(reg.read_names) if rr.read_dict[nme].round2_repeat_size is not None]
    lo, hi = ladder_bounds_array([rr.read_dict[reg.read_names[i]].round2_repeat_size for i in ok], False)
    b3.add_round3(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq, [reg.core_seqs[i] for i in ok], lo, hi)
b2.commit(); b3.commit()
st = torch.cuda.Stream()
for name, b in (("round2", b2), ("round3", b3)):
    for _ in range(3): b.run(st.cuda_stream)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); b.run(st.cuda_stream); e1.record(st); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    s = b.stats()
    print(name, "ms", [round(t, 3) for t in ts], "executed GCUPS", round(s["executed_cells"] / min(ts) / 1e6, 1),
          "effective GCUPS", round(s["algorithmic_cells"] / min(ts) / 1e6, 1), "launches", s["kernel_launches"])
