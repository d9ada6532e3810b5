import sys, time
sys.path.insert(0, ".")
import numpy as np
import nanorepeat_b200 as nrb
from nanorepeat_b200 import synth, engine, estimation as est
regs = synth.config2(seed=2, n_reads=5000)
sc = engine.get_preset("ont")
def fresh(): return [nrb.RepeatRegion.from_synth(r) for r in regs]
for _ in range(3): nrb.estimate_regions(fresh(), "ont", False)
def t(label, fn):
    t0 = time.perf_counter(); r = fn(); print(f"{label:28s} {(time.perf_counter()-t0)*1e3:7.3f} ms"); return r
rrs = fresh()
# python-side pieces of _round2_launch for one region
rr = rrs[0]
reads = rr.read_dict
read_list = t("list(reads.values())", lambda: list(reads.values()))
r1 = t("r1 array", lambda: (np.array([rd.dist_between_anchors for rd in read_list], dtype=np.float64) / 3.0).tolist())
def setr1():
    for rd, v in zip(read_list, r1): rd.round1_repeat_size = v
t("set r1 attrs", setr1)
qn = t("qnames", lambda: list(reads))
cores = t("cores_of", lambda: [rr.read_core_seq_dict[n] for n in qn])
buf = t("concat", lambda: engine._concat(cores))
b = engine.Batch.begin(sc, "round2_flags")
t("add_round2 (pack)", lambda: b.add_round2(rr.left_anchor_seq, rr.repeat_unit_seq, 84, cores))
t("commit (plan+upload)", lambda: b.commit())
t("run", lambda: b.run())
res = t("fetch_round2", lambda: b.fetch_round2())
score, tend, inside = res
def sel():
    ok = (score >= 80) & inside & (tend >= 1000)
    r2 = ((tend - 1000).astype(np.float64) / 3.0).tolist()
    for name, good, v in zip(qn, ok.tolist(), r2):
        if good: reads[name].round2_repeat_size = v
t("round2 selection + attrs", sel)
b3 = engine.Batch.begin_round3_from(b)
rl = [reads[n] for n in qn]
def bounds():
    r2 = [rd.round2_repeat_size for rd in rl]
    valid = np.array([v is not None for v in r2], dtype=bool)
    lo, hi = est.ladder_bounds_array([v for v in r2 if v is not None], False)
    kmin = np.zeros(len(rl), np.int32); kmax = np.full(len(rl), -1, np.int32); kmin[valid], kmax[valid] = lo, hi
    return kmin, kmax
kmin, kmax = t("ladder bounds (python)", bounds)
t("add_round3_reuse", lambda: b3.add_round3_reuse(0, rr.right_anchor_seq, kmin, kmax))
t("commit r3", lambda: b3.commit())
t("run r3", lambda: b3.run())
s = t("fetch_round3", lambda: b3.fetch_round3())
t("assign r3", lambda: est._assign_round3(rl, *s))
