#!/usr/bin/env python
"""Rounds 1-3 through the operator API on scaled-down versions of BASELINE.json's other configs (1, 3, 4, 5): wall time
of estimate_regions, reads/s, algorithmic and executed GCUPS.  Informational (bench.py's metric is config 2).
Needs a GPU.  usage: config_sweep.py [out.jsonl]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine
from nanorepeat_b200.estimation import ladder_bounds

engine.init(0)
cases = [
    ("config 1: 15 STR regions x 30 ont_q20 reads", lambda: synth.config1(seed=1)),
    ("config 3 (scaled): 2000 loci x 30 HiFi reads, 2-6 bp motifs", lambda: synth.config3(seed=3, n_loci=2000)),
    ("config 4 (scaled): C9orf72 ~1000 x GGGGCC and FMR1 ~500 x CGG, 40 R9 reads per locus", lambda: synth.config4(seed=4, reads_per_locus=40)),
    ("config 5 (scaled): 200 regions x 50 reads, k log-uniform 1..2000, ont/clr", lambda: synth.config5(seed=5, n_reads=10000)),
]
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
for name, make in cases:
    t0 = time.perf_counter(); regs = make(); gen_s = time.perf_counter() - t0
    def fresh(): return [nrb.RepeatRegion.from_synth(r) for r in regs]
    nrb.estimate_regions(fresh(), "ont", False)          # warm-up (allocator classes, kernel attributes)
    best = None
    for _ in range(3):
        rrs = fresh()
        t0 = time.perf_counter(); nrb.estimate_regions(rrs, "ont", False); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    n_reads = sum(len(r.core_seqs) for r in regs)
    cells = 0; n_r3 = 0; q = []
    for reg, rr in zip(regs, rrs):
        nl, nr_, m = len(reg.left_anchor_seq), len(reg.right_anchor_seq), len(reg.repeat_unit_seq)
        r1max = max(float(d) / m for d in reg.dist_between_anchors)
        T = int(r1max * 1.5) + 1
        if T < r1max + 10: T = int(r1max + 10)
        for nme, core in zip(reg.read_names, reg.core_seqs):
            q.append(len(core))
            r2 = rr.read_dict[nme].round2_repeat_size
            cells += len(core) * (nl + m * T)
            if r2 is not None:
                lo, hi = ladder_bounds(r2, False)
                cells += synth.algorithmic_cells(nl, nr_, m, len(core), T, lo, hi)[1]
            n_r3 += rr.read_dict[nme].round3_repeat_size is not None
    line = {"config": name, "regions": len(regs), "reads": n_reads, "reads_with_round3": n_r3,
            "core_len_median": float(np.median(q)), "core_len_max": int(max(q)),
            "e2e_ms": best * 1e3, "reads_per_s": n_reads / best, "algorithmic_gcups_e2e": cells / best / 1e9}
    print(json.dumps(line), flush=True)
    if out: out.write(json.dumps(line) + "\n")
