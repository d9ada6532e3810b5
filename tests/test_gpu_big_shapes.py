"""GPU parity on the shapes BASELINE.json's configs 4 and 5 actually contain (SURVEY.md section 8 shape table): the
cooperative long-read path (reads cut into stripes that run on many warps at once) against the CPU oracle, through the
C ABI and through the operator API.

    c9     C9orf72-like: GGGGCC x ~1000, core ~6.2 kb, ladder [~950, ~1050] (101 rungs), 1000-bp anchors, R9 errors
    top5   config 5's largest: 6-bp motif x 2000, core ~12.5 kb, 201 rungs, templates up to 14.6 kb, clr errors
    cap    >= 3000 units: the ladder half-width hits the 150 cap (nanoRepeat_bam.py:464-465) -> 301 rungs

The ladder bounds come from the reference's own rule applied to the oracle's round 2 (oracle/selection.py, pinned
against the reference's functions by tests/golden), so the same records serve the per-rung checks (modes 2 and 1), the
selection checks (mode 3) and the operator-API check.  One oracle pass (~5e10 cells, rung-parallel) per session."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _region_exact(rng, name, motif, ks, profile):
    """One read per entry of ks (make_region draws alleles at random; here every listed size appears once)."""
    from nanorepeat_b200 import synth
    reg = synth.SynthRegion(name, synth.random_seq(rng, 1000), synth.random_seq(rng, 1000), motif, "ont")
    for i, k in enumerate(ks):
        core, dist, _ = synth.simulate_core(rng, reg.left_anchor_seq, reg.right_anchor_seq, motif, k, profile)
        reg.read_names.append(f"{name}_read{i}")
        reg.core_seqs.append(core)
        reg.dist_between_anchors.append(dist)
        reg.true_sizes.append(k)
    return reg


@pytest.fixture(scope="module")
def big(oracle):
    """The three regions, the oracle's rounds 1-3 on them (with every rung's record), and a region of 200 short reads."""
    from nanorepeat_b200 import synth
    from oracle import selection
    rng = np.random.default_rng(20261018)
    regs = [
        _region_exact(rng, "c9", "GGGGCC", [1000, 8], "ont_r9"),          # the expansion and the normal allele beside it
        _region_exact(rng, "top5", "CTGGAA", [2000], "clr"),
        _region_exact(rng, "cap", "AC", [3050, 3000], "ont"),
    ]
    short = synth.config2(seed=5, n_reads=100)                            # 2 regions x 100 reads: pairs
    exp = []
    for reg in regs + short:
        exp.append(selection.estimate_region(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                                             reg.core_seqs, reg.dist_between_anchors, n_threads=oracle.max_threads()))
    # the shapes are the ones the test is named after (buffer = int(0.05 * r2): 49 / 99 rungs either side; 150 at the cap)
    assert len(regs[0].core_seqs[0]) > 5500 and exp[0]["kmax"][0] - exp[0]["kmin"][0] + 1 >= 95
    assert len(regs[1].core_seqs[0]) > 11500 and exp[1]["kmax"][0] - exp[1]["kmin"][0] + 1 >= 195
    assert exp[2]["kmax"][0] - exp[2]["kmin"][0] + 1 == 301 and exp[2]["kmax"][1] - exp[2]["kmin"][1] + 1 == 301
    return regs, short, exp


def _specs(regs, exp):
    """Round-3 batch inputs over the reads round 2 gave a size (region order kept)."""
    specs = []
    for reg, e in zip(regs, exp):
        idx = e["round3_idx"]
        specs.append((reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq, [reg.core_seqs[i] for i in idx],
                      np.array([e["kmin"][i] for i in idx], np.int32), np.array([e["kmax"][i] for i in idx], np.int32)))
    return specs


def _expected_selection(reg, e, min_score):
    """(top, n, sum) per read with a ladder, from the oracle's rung records (nanoRepeat_bam.py:423-431)."""
    out, off = e["round3_rungs"]
    nl, nr_, m = len(reg.left_anchor_seq), len(reg.right_anchor_seq), len(reg.repeat_unit_seq)
    rows = []
    for j, i in enumerate(e["round3_idx"]):
        a = out[int(off[j]):int(off[j + 1])]
        ks = np.arange(len(a)) + e["kmin"][i]
        tlen = nl + m * ks + nr_
        spans = (a["score"] > 0) & (tlen - a["tend"] < nr_) & (a["tstart"] < nl)
        ok = a["score"] >= max(1, min_score)
        t = int(a["score"][ok].max()) if ok.any() else 0
        sel = ks[(a["score"] == t) & spans] if t > 0 else ks[:0]
        rows.append((t, len(sel), int(sel.sum())))
    return rows


@pytest.mark.parametrize("mode", [3, 2, 1])
def test_big_ladders_equal_oracle(engine, big, mode):
    """All three long regions in ONE batch (their stripes share the launch), every ladder mode."""
    regs, _short, exp = big
    sc = engine.get_preset("ont")
    specs = _specs(regs, exp[:len(regs)])
    engine.set_ladder_mode(mode)
    try:
        with engine.Batch.begin(sc, "round3") as b:
            for spec in specs:
                b.add_round3(*spec)
            b.commit().run()
            if mode == 1:
                got = b.fetch_alns()
                ref = np.concatenate([e["round3_rungs"][0] for e in exp[:len(regs)]])
                assert len(got) == len(ref)
                bad = np.flatnonzero((got["score"] != ref["score"]) | (got["tstart"] != ref["tstart"]) | (got["tend"] != ref["tend"]))
                assert len(bad) == 0, (len(bad), bad[:5], got[bad[:5]], ref[bad[:5]])
                return
            if mode == 2:
                sum_k, n_k, top, rungs, off = b.fetch_round3(want_rungs=True)
                ref = np.concatenate([e["round3_rungs"][0] for e in exp[:len(regs)]])
                assert np.array_equal(rungs["score"], ref["score"])
            else:
                sum_k, n_k, top = b.fetch_round3()
    finally:
        engine.set_ladder_mode(3)
    want = [row for reg, e in zip(regs, exp) for row in _expected_selection(reg, e, sc.min_dp_score)]
    got = [(int(t), int(n), int(s)) for t, n, s in zip(top, n_k, sum_k)]
    assert got == want
    assert all(n >= 1 for _t, n, _s in want[:1])      # the expansion's ladder is decided by a spanning rung


def test_big_reads_mixed_with_short_pairs(engine, big):
    """The long reads' stripes and 200 short pairs in one launch (the fused kernel deals both)."""
    regs, short, exp = big
    sc = engine.get_preset("ont")
    order = [short[0], regs[1], regs[0], short[1], regs[2]]
    exps = [exp[3], exp[1], exp[0], exp[4], exp[2]]
    specs = _specs(order, exps)
    with engine.Batch.begin(sc, "round3") as b:
        for spec in specs:
            b.add_round3(*spec)
        b.commit()
        li = b.launch_info()
        assert li["n_pairs"] >= 90 and li["n_rest"] > 100          # pairs and long-read stripes side by side
        sum_k, n_k, top = b.run().fetch_round3()
    want = [row for reg, e in zip(order, exps) for row in _expected_selection(reg, e, sc.min_dp_score)]
    assert [(int(t), int(n), int(s)) for t, n, s in zip(top, n_k, sum_k)] == want


def test_big_round2_records_equal_oracle(engine, big):
    """Round 2 of the long reads (templates up to 19 kb): exact-record kind and flags kind against the oracle."""
    regs, _short, exp = big
    sc = engine.get_preset("ont")
    with engine.Batch.begin(sc, "round2") as b, engine.Batch.begin(sc, "round2_flags") as f:
        for reg, e in zip(regs, exp):
            b.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, e["T"], reg.core_seqs)
            f.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, e["T"], reg.core_seqs)
        got = b.commit().run().fetch_alns()
        score, tend, inside = f.commit().run().fetch_round2()
    ref = np.concatenate([e["round2_aln"] for e in exp[:len(regs)]])
    assert np.array_equal(got, ref), (got, ref)
    assert np.array_equal(score, ref["score"]) and np.array_equal(tend, ref["tend"])
    n_left = np.concatenate([[len(r.left_anchor_seq)] * len(r.core_seqs) for r in regs])
    assert np.array_equal(inside, ref["tstart"] <= n_left)


def test_big_reads_through_operator_api(engine, big):
    """estimate_regions on RepeatRegion / Read objects == the oracle's rounds 1-3 (r1, r2, r3 per read)."""
    import nanorepeat_b200 as nrb
    regs, short, exp = big
    rrs = [nrb.RepeatRegion.from_synth(r) for r in regs + short]
    nrb.estimate_regions(rrs, "ont", False)
    for reg, rr, e in zip(regs + short, rrs, exp):
        for i, name in enumerate(reg.read_names):
            rd = rr.read_dict[name]
            assert rd.round1_repeat_size == e["r1"][i], (reg.name, i)
            assert rd.round2_repeat_size == e["r2"][i], (reg.name, i)
            g3, e3 = rd.round3_repeat_size, e["r3"][i]
            assert (None if g3 is None else float(g3)) == (None if e3 is None else float(e3)), (reg.name, i, g3, e3)
    # the long reads were decided by their ladders, not by the fall-back to round 2
    assert abs(float(rrs[0].read_dict["c9_read0"].round3_repeat_size) - 1000) < 60
    assert abs(float(rrs[1].read_dict["top5_read0"].round3_repeat_size) - 2000) < 150
