import sys, json, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from nanorepeat_b200 import engine
from oracle import nr_oracle as oracle
doc = load_golden("cfg2_small")
sc = engine.get_preset("ont")
for reg in doc["regions"]:
    left, motif = reg["left"], reg["motif"]
    cores = reg["cores"]
    r1 = [d / len(motif) for d in reg["dists"]]
    mx = max(r1); T = int(mx * 1.5) + 1
    if T < mx + 10: T = int(mx + 10)
    tpl = left + motif * T
    ref = oracle.align_batch(cores, [tpl] * len(cores), n_threads=4)
    with engine.Batch.begin(sc, "round2_flags") as b:
        b.add_round2(left, motif, T, cores)
        score, tend, inside = b.commit().run().fetch_round2()
    bad = 0
    for i in range(len(cores)):
        exp = (int(ref["score"][i]), int(ref["tend"][i]), bool(ref["tstart"][i] <= len(left)))
        got = (int(score[i]), int(tend[i]), bool(inside[i]))
        if exp != got:
            bad += 1
            if bad < 6: print(reg["name"], i, "q", len(cores[i]), "t", len(tpl), "nl", len(left), "got", got, "exp", exp, "tstart", int(ref["tstart"][i]))
    print(reg["name"], "reads", len(cores), "bad", bad, "T", T)
    print("qlens", [len(c) for c in cores])
    for i in range(len(cores)):
        with engine.Batch.begin(sc, "round2_flags") as b:
            b.add_round2(left, motif, T, [cores[i]])
            s1, t1, i1 = b.commit().run().fetch_round2()
        with engine.Batch.begin(sc, "round2_flags") as b:
            b.add_round2(left, motif, T, [cores[i], cores[i]])
            s2, t2, i2 = b.commit().run().fetch_round2()
        print(i, "alone", int(s1[0]), int(t1[0]), "self-pair", int(s2[0]), int(t2[0]), int(s2[1]), int(t2[1]), "exp", int(ref["score"][i]), int(ref["tend"][i]))
