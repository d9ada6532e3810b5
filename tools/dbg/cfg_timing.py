"""Resident kernel time and executed GCUPS of rounds 2 and 3 for a scaled config (4 or 5): where the long-read path stands."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine
from nanorepeat_b200.estimation import ladder_bounds_array
which = sys.argv[1] if len(sys.argv) > 1 else "5"
regs = (synth.config5(seed=5, n_reads=10000) if which == "5" else synth.config3(seed=3, n_loci=2000) if which == "3"
        else synth.config1(seed=1) if which == "1" else synth.config4(seed=4, reads_per_locus=40))
engine.init(0)
sc = engine.get_preset("ont")
rrs = [nrb.RepeatRegion.from_synth(r) for r in regs]
nrb.estimate_regions(rrs, "ont", False)
engine.set_timing(True)
stream = torch.cuda.Stream()
b2 = engine.Batch.begin(sc, "round2_flags")
for reg in regs:
    m = len(reg.repeat_unit_seq)
    r1max = max(float(d) / m for d in reg.dist_between_anchors)
    T = int(r1max * 1.5) + 1
    if T < r1max + 10: T = int(r1max + 10)
    b2.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, T, reg.core_seqs)
b2.commit()
b3 = engine.Batch.begin_round3_from(b2)
for i, (reg, rr) in enumerate(zip(regs, rrs)):
    r2 = [rr.read_dict[n].round2_repeat_size for n in reg.read_names]
    ok = np.array([v is not None for v in r2])
    lo = np.zeros(len(r2), np.int32); hi = np.full(len(r2), -1, np.int32)
    if ok.any():
        a, b = ladder_bounds_array([v for v in r2 if v is not None], False); lo[ok], hi[ok] = a, b
    b3.add_round3_reuse(i, reg.right_anchor_seq, lo, hi)
b3.commit()
q = np.array([len(c) for r in regs for c in r.core_seqs])
print("reads", len(q), "q<=384:", int((q <= 384).sum()), "q>384:", int((q > 384).sum()), "q max", int(q.max()))
for name, b in (("round2", b2), ("round3", b3)):
    for _ in range(2): b.run(stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); b.run(stream.cuda_stream); e1.record(stream); e1.synchronize()
    li, st = b.launch_info(), b.stats()
    ms = e0.elapsed_time(e1)
    print(name, "ms", round(ms, 3), "pairs", li["n_pairs"], "rest entries", li["n_rest"], "paired cells %.3g" % li["paired_cells"], "rest cells %.3g" % li["rest_cells"],
          "executed GCUPS", round((li["paired_cells"] + li["rest_cells"]) / ms / 1e6), "algorithmic GCUPS", round(st["algorithmic_cells"] / ms / 1e6))
