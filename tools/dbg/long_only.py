import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine
from nanorepeat_b200.estimation import ladder_bounds_array
engine.init(0)
sc = engine.get_preset("ont")
regs = synth.config2(seed=2, n_reads=5000)
rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
nrb.estimate_regions(rrs, "ont", False)
engine.set_timing(True)
stream = torch.cuda.Stream()
mode = int(os.environ.get("LADDER_MODE", "3")); engine.set_ladder_mode(mode)
nmax = int(os.environ.get("NMAX", "1000"))
b3 = engine.Batch.begin(sc, "round3"); n = 0
for reg, rr in zip(regs, rrs):
    idx = [i for i, nme in enumerate(reg.read_names) if len(reg.core_seqs[i]) > 384 and rr.read_dict[nme].round2_repeat_size is not None][:nmax]
    cores = [reg.core_seqs[i] for i in idx]
    lo, hi = ladder_bounds_array([rr.read_dict[reg.read_names[i]].round2_repeat_size for i in idx], False)
    b3.add_round3(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq, cores, lo, hi); n += len(idx)
b3.commit()
for _ in range(3): b3.run(stream.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); b3.run(stream.cuda_stream); e1.record(stream); e1.synchronize()
print(os.environ.get("NR_COOP_ROWS"), "n", n, "ms", round(e0.elapsed_time(e1), 3), b3.launch_info(), "qmax", max(len(c) for r in regs for c in r.core_seqs))
