#!/usr/bin/env python
"""Where a config-2 step's time goes: rounds 2 and 3 over (all reads | reads that pair up | long reads only), each batch
run alone with CUDA events around every kernel (engine.set_timing).  Needs a GPU.  usage: split_timing.py [n_reads]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine
from nanorepeat_b200.estimation import ladder_bounds_array

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
engine.init(0)
sc = engine.get_preset("ont")
regs = synth.config2(seed=2, n_reads=n_reads)
rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
nrb.estimate_regions(rrs, "ont", False)
engine.set_timing(True)
stream = torch.cuda.Stream()
for label, keep in (("all", lambda q: True), ("short (q <= 384)", lambda q: q <= 384), ("long (q > 384)", lambda q: q > 384)):
    b2 = engine.Batch.begin(sc, "round2_flags"); b3 = engine.Batch.begin(sc, "round3")
    n = 0
    for reg, rr in zip(regs, rrs):
        m = len(reg.repeat_unit_seq)
        r1max = max(float(d) / m for d in reg.dist_between_anchors)
        T = int(r1max * 1.5) + 1
        if T < r1max + 10: T = int(r1max + 10)
        idx = [i for i, nme in enumerate(reg.read_names) if keep(len(reg.core_seqs[i])) and rr.read_dict[nme].round2_repeat_size is not None]
        cores = [reg.core_seqs[i] for i in idx]
        lo, hi = ladder_bounds_array([rr.read_dict[reg.read_names[i]].round2_repeat_size for i in idx], False) if idx else (np.zeros(0, np.int32),) * 2
        b2.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, T, cores)
        b3.add_round3(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq, cores, lo, hi)
        n += len(idx)
    for name, b in (("round2", b2.commit()), ("round3", b3.commit())):
        for _ in range(3):
            b.run(stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); b.run(stream.cuda_stream); e1.record(stream); e1.synchronize()
        li = b.launch_info()
        print(json.dumps({"reads": label, "n": n, "round": name, "ms": round(e0.elapsed_time(e1), 3), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in li.items()}}), flush=True)
    b2.close(); b3.close()
