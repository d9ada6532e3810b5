"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle and the golden fixtures.
Integer work: everything is compared bit-exact."""
import random

import numpy as np
import pytest

from conftest import GOLDEN_SETS, load_golden

pytestmark = pytest.mark.gpu


def _rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def _mutate(rng, s, rate):
    out = []
    for ch in s:
        u = rng.random()
        if u < rate / 3:
            continue
        if u < 2 * rate / 3:
            out.append(rng.choice("ACGT"))
            continue
        if u < rate:
            out.append(rng.choice("ACGT"))
        out.append(ch)
    return "".join(out)


def _assert_same(got, ref, ctx=""):
    if not np.array_equal(got, ref):
        bad = [i for i in range(len(ref)) if tuple(got[i]) != tuple(ref[i])]
        raise AssertionError(f"{ctx}: {len(bad)} of {len(ref)} records differ; first: task {bad[0]} got "
                             f"{tuple(got[bad[0]])} expected {tuple(ref[bad[0]])}")


def test_micro_known_answers(engine, oracle):
    from test_oracle import micro_cases
    cases = micro_cases()
    sc = engine.get_preset("ont")
    got = engine.score_tasks([c[1] for c in cases], [c[2] for c in cases], sc)
    for c, g in zip(cases, got):
        assert tuple(int(v) for v in g) == c[3], c[0]


@pytest.mark.parametrize("seed,qmax,tmax,n", [(1, 40, 90, 400), (2, 300, 700, 300), (3, 520, 2400, 120),
                                              (4, 1500, 3000, 40), (5, 130, 130, 300)])
def test_random_tasks_match_oracle(engine, oracle, seed, qmax, tmax, n):
    """Every stripe height (R = 4..16), single- and multi-stripe, related and unrelated pairs."""
    rng = random.Random(seed)
    qs, ts = [], []
    for it in range(n):
        ql = rng.randint(1, qmax)
        q = _rand_seq(rng, ql)
        kind = it % 4
        if kind == 0:
            t = _rand_seq(rng, rng.randint(1, tmax))
        elif kind == 1:      # query embedded with errors
            pad = rng.randint(0, max(0, tmax - ql) // 2)
            t = _rand_seq(rng, pad) + _mutate(rng, q, 0.1) + _rand_seq(rng, rng.randint(0, pad))
        elif kind == 2:      # low-complexity repeats: many ties
            unit = _rand_seq(rng, rng.randint(1, 6))
            q = (unit * (ql // len(unit) + 1))[:ql]
            t = _rand_seq(rng, rng.randint(0, 30)) + unit * rng.randint(1, max(1, tmax // (2 * len(unit)))) + _rand_seq(rng, rng.randint(0, 30))
        else:                # long gap in the middle (second affine piece)
            gap = rng.randint(1, 60)
            t = q[:ql // 2] + _rand_seq(rng, gap) + q[ql // 2:]
        qs.append(q)
        ts.append(t[:tmax] if len(t) > tmax else t)
    sc = engine.get_preset("ont")
    got = engine.score_tasks(qs, ts, sc)
    ref = oracle.align_batch(qs, ts, n_threads=oracle.max_threads())
    _assert_same(got, ref, f"seed {seed}")


def test_stripe_boundaries_exact_sizes(engine, oracle):
    """Query lengths around every stripe-shape boundary (32*R, 512, 1024 ...)."""
    rng = random.Random(9)
    qs, ts = [], []
    for ql in [1, 2, 31, 32, 33, 127, 128, 129, 160, 161, 287, 288, 289, 511, 512, 513, 544, 1023, 1024, 1025, 1537]:
        q = _rand_seq(rng, ql)
        t = _rand_seq(rng, 40) + _mutate(rng, q, 0.08) + _rand_seq(rng, 70)
        qs.append(q)
        ts.append(t)
    for tl in [1, 15, 16, 17, 31, 32, 33, 63, 64, 65]:
        q = _rand_seq(rng, 50)
        qs.append(q)
        ts.append((q + _rand_seq(rng, 100))[:tl])
    sc = engine.get_preset("hifi")
    _assert_same(engine.score_tasks(qs, ts, sc), oracle.align_batch(qs, ts), "boundaries")


def test_other_scoring_values(engine, oracle):
    rng = random.Random(21)
    qs = [_rand_seq(rng, rng.randint(20, 200)) for _ in range(60)]
    ts = [_rand_seq(rng, 20) + _mutate(rng, q, 0.15) + _rand_seq(rng, 20) for q in qs]
    for kw in (dict(match=1, mismatch=3, gap_open1=5, gap_ext1=2, gap_open2=30, gap_ext2=1),
               dict(match=3, mismatch=2, gap_open1=1, gap_ext1=3, gap_open2=10, gap_ext2=2)):
        osc = oracle.scoring(**kw)
        sc = engine.Scoring(**{n: getattr(osc, n) for n, _ in engine.Scoring._fields_})
        _assert_same(engine.score_tasks(qs, ts, sc), oracle.align_batch(qs, ts, osc), str(kw))


def test_empty_and_degenerate_inputs(engine):
    sc = engine.get_preset("ont")
    assert len(engine.score_tasks([], [], sc)) == 0
    got = engine.score_tasks(["", "ACGT", ""], ["ACGT", "", ""], sc)
    assert [tuple(int(v) for v in g) for g in got] == [(0, 0, 0)] * 3
    assert len(engine.round2_region(sc, "ACGT" * 10, "CAG", 5, [])) == 0
    s, n, t = engine.round3_region(sc, "ACGT" * 10, "TTGA" * 10, "CAG", [], np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert len(s) == 0 and len(n) == 0 and len(t) == 0
    # empty ladder (kmax < kmin) for one read
    s, n, t = engine.round3_region(sc, "ACGT" * 10, "TTGA" * 10, "CAG", ["ACGTACGTCAGCAGTTGATTGA"], [3], [2])
    assert (int(s[0]), int(n[0]), int(t[0])) == (0, 0, 0)


def test_errors_are_reported_not_fatal(engine):
    sc = engine.get_preset("ont")
    with pytest.raises(ValueError):
        engine.get_preset("pacbio")
    with pytest.raises(engine.NanoRepeatB200Error) as ei:
        engine.Batch.begin(sc, "round3").add_round3("ACGT", "ACGT", "CAG", ["ACGT"], [-1], [3])
    assert ei.value.code == -2
    # the library is still usable afterwards
    assert tuple(int(v) for v in engine.score_tasks(["ACGT"], ["ACGT"], sc)[0]) == (8, 0, 4)


def test_unscorable_tasks_are_isolated_not_fatal(engine, oracle):
    """One read beyond the packed range, one template with an N: their records are zero ("the aligner printed nothing"),
    nr_stats_t.n_skipped counts them, and every other task of the same call is scored as usual."""
    rng = random.Random(5)
    sc = engine.get_preset("ont")
    good_q = [_rand_seq(rng, rng.randint(30, 700)) for _ in range(6)]
    good_t = [_rand_seq(rng, 20) + _mutate(rng, q, 0.05) + _rand_seq(rng, 20) for q in good_q]
    qs = good_q[:3] + ["A" * 20000, "ACGTACGTAC"] + good_q[3:]
    ts = good_t[:3] + ["A" * 20000, "ACGTNCGTAC"] + good_t[3:]
    b = engine.Batch.tasks(sc, qs, ts)
    got = b.run().fetch_alns()
    assert b.stats()["n_skipped"] == 2
    b.close()
    ref = oracle.align_batch(good_q, good_t)
    _assert_same(np.concatenate([got[:3], got[5:]]), ref, "tasks beside the skipped ones")
    assert [tuple(int(v) for v in g) for g in got[3:5]] == [(0, 0, 0)] * 2
    # operator layer: a region whose left anchor holds an N sits between two normal regions
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth
    regs = synth.config1(seed=41, n_regions=3, reads_per_region=8)
    plain = [nrb.RepeatRegion.from_synth(r) for r in regs]
    nrb.estimate_regions(plain, "ont", False)
    mixed = [nrb.RepeatRegion.from_synth(r) for r in regs]
    mixed[1].left_anchor_seq = mixed[1].left_anchor_seq[:500] + "N" + mixed[1].left_anchor_seq[501:]
    nrb.estimate_regions(mixed, "ont", False)
    for i in (0, 2):
        for name, rd in plain[i].read_dict.items():
            o = mixed[i].read_dict[name]
            assert (rd.round1_repeat_size, rd.round2_repeat_size, rd.round3_repeat_size) == \
                   (o.round1_repeat_size, o.round2_repeat_size, o.round3_repeat_size)
    for rd in mixed[1].read_dict.values():
        assert rd.round1_repeat_size is not None and rd.round2_repeat_size is None and rd.round3_repeat_size is None


def _with_ambiguous(rng, s, rate):
    """Replace a fraction of the bases by N / other IUPAC letters / lower-case n."""
    out = list(s)
    for i in range(len(out)):
        if rng.random() < rate:
            out[i] = rng.choice("NNNnRYK")
    return "".join(out)


def test_ambiguous_read_bases_equal_oracle(engine, oracle):
    """Reads with bases other than ACGT (the reference accepts them; minimap2 scores them -1 against anything):
    exact records, round-2 flags kind, and every ladder mode, on paired, single and multi-stripe reads, next to clean
    reads of the same region."""
    rng = random.Random(606)
    sc = engine.get_preset("ont")
    qs, ts = [], []
    for i in range(90):
        q = _rand_seq(rng, rng.choice([5, 40, 130, 400, 700, 1500]))
        t = _rand_seq(rng, rng.randint(0, 50)) + _mutate(rng, q, 0.06) + _rand_seq(rng, rng.randint(0, 50))
        qs.append(_with_ambiguous(rng, q, rng.choice([0.0, 0.01, 0.2])) if i % 7 else "N" * len(q))
        ts.append(t)
    _assert_same(engine.score_tasks(qs, ts, sc), oracle.align_batch(qs, ts, n_threads=oracle.max_threads()), "ambiguous queries")
    # round 2 + round 3 of one region: clean and ambiguous reads pair up with each other
    left, right, motif = _rand_seq(rng, 200), _rand_seq(rng, 180), "CAG"
    cores, kmin, kmax = [], [], []
    for i in range(41):
        k = rng.choice([4, 17, 55, 150, 300])          # up to ~1.1 kb: stripes
        core = _mutate(rng, left[-70:] + motif * k + right[:80], 0.04)
        if i % 3:
            core = _with_ambiguous(rng, core, rng.choice([0.005, 0.05]))
        cores.append(core + ("\n" if i % 5 == 0 else ""))      # white space around a read is not part of it
        kmin.append(max(0, k - rng.randint(2, 9))); kmax.append(k + rng.randint(2, 9))
    kmin, kmax = np.array(kmin, np.int32), np.array(kmax, np.int32)
    stripped = [c.strip() for c in cores]
    T = 320
    ref2 = oracle.align_batch(stripped, [left + motif * T] * len(cores), n_threads=oracle.max_threads())
    with engine.Batch.begin(sc, "round2_flags") as b:
        b.add_round2(left, motif, T, cores, lines=False)      # (cores with a newline cannot travel as lines)
        score, tend, inside = b.commit().run().fetch_round2()
        assert b.stats()["n_skipped"] == 0
    assert np.array_equal(score, ref2["score"])
    spans = ref2["tend"] >= len(left)
    assert np.array_equal(tend[spans], ref2["tend"][spans])
    live = (ref2["score"] > 0) & spans
    assert np.array_equal(inside[live], (ref2["tstart"] <= len(left))[live])
    ref, roff = oracle.align_ladders(stripped, left, right, motif, kmin, kmax, n_threads=oracle.max_threads())
    try:
        for mode in (3, 2, 1, 0):
            engine.set_ladder_mode(mode)
            with engine.Batch.round3(sc, left, right, motif, cores, kmin, kmax) as b:
                b.run()
                if mode >= 2:
                    _assert_flag_ladder(b, ref, roff, kmin, len(left), len(right), 3, sc.min_dp_score,
                                        f"ambiguous reads, mode {mode}", rungs_too=mode == 2)
                else:
                    _assert_same(b.fetch_alns(), ref, f"ambiguous reads, mode {mode}")
    finally:
        engine.set_ladder_mode(3)


def test_scratch_reuse_across_launches_and_eras(engine, oracle):
    """The long reads' boundary rows live in pooled scratch that still holds the tagged entries of earlier launches;
    tags carry a 10-bit launch epoch, and a buffer is cleared when the epochs wrap (every 1023 launches).  Same batch
    launched across an era boundary, and a different batch over the same pooled buffer: always the oracle's records."""
    rng = random.Random(99)
    sc = engine.get_preset("ont")

    def make(n):
        qs = [_rand_seq(rng, rng.randint(600, 1400)) for _ in range(n)]
        ts = [_rand_seq(rng, 40) + _mutate(rng, q, 0.08) + _rand_seq(rng, 40) for q in qs]
        return qs, ts

    qa, ta = make(6)
    ref_a = oracle.align_batch(qa, ta, n_threads=oracle.max_threads())
    b = engine.Batch.tasks(sc, qa, ta)
    for it in range(1100):                     # > 1023 launches: crosses an era boundary at least once
        b.run()
        if it % 97 == 0 or it > 1090:
            _assert_same(b.fetch_alns(), ref_a, f"launch {it}")
    b.close()                                  # its scratch returns to the pool, full of valid-looking entries
    for _ in range(3):
        qb, tb = make(6)
        _assert_same(engine.score_tasks(qb, tb, sc), oracle.align_batch(qb, tb, n_threads=oracle.max_threads()), "reused scratch")


def test_round3_rungs_match_oracle(engine, oracle):
    from nanorepeat_b200 import synth
    reg = synth.config1(seed=3, n_regions=1, reads_per_region=10)[0]
    sc = engine.get_preset("ont_q20")
    n = len(reg.core_seqs)
    kmin = np.array([max(0, k - 15) for k in reg.true_sizes], dtype=np.int32)
    kmax = np.array([k + 15 for k in reg.true_sizes], dtype=np.int32)
    engine.set_ladder_mode(2)            # the paired ladder (mode 3) keeps no rung records
    try:
        sum_k, n_k, top, rungs, off = engine.round3_region(sc, reg.left_anchor_seq, reg.right_anchor_seq,
                                                           reg.repeat_unit_seq, reg.core_seqs, kmin, kmax, want_rungs=True)
    finally:
        engine.set_ladder_mode(3)
    ref, roff = oracle.align_ladders(reg.core_seqs, reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                                     kmin, kmax, n_threads=oracle.max_threads())
    assert np.array_equal(off, roff)
    nl, nr_, m = len(reg.left_anchor_seq), len(reg.right_anchor_seq), len(reg.repeat_unit_seq)
    for r in range(n):
        for i in range(int(off[r]), int(off[r + 1])):
            k = int(kmin[r]) + i - int(off[r])
            a = ref[i]
            assert int(rungs[i]["score"]) == int(a["score"])
            in_right = bool(a["score"] > 0 and (nl + m * k + nr_) - a["tend"] < nr_)
            assert bool(rungs[i]["ends_in_right"]) == in_right
            assert bool(rungs[i]["starts_in_left"]) == bool(in_right and a["tstart"] < nl)     # conjunction (:427)


def _assert_flag_ladder(b, ref, roff, kmin, n_left, n_right, m, min_score, ctx, rungs_too=True):
    """Flag ladder (modes 2 and 3) == oracle on what the reference reads per rung (score, span predicates, :423-427;
    mode 2 only: mode 3 keeps no rung records) and on the selection (top score, tied rungs that span both flanks)."""
    if rungs_too:
        sum_k, n_k, top, rungs, off = b.fetch_round3(want_rungs=True)
        assert np.array_equal(off, roff), ctx
        assert np.array_equal(rungs["score"], ref["score"]), ctx
    else:
        sum_k, n_k, top = b.fetch_round3()
        off = roff
    for r in range(len(kmin)):
        lo, hi = int(off[r]), int(off[r + 1])
        ks = np.arange(hi - lo) + int(kmin[r])
        a = ref[lo:hi]
        tlen = n_left + m * ks + n_right
        in_right = (a["score"] > 0) & (tlen - a["tend"] < n_right)
        in_left = in_right & (a["tstart"] < n_left)
        if rungs_too:
            assert np.array_equal(rungs["ends_in_right"][lo:hi].astype(bool), in_right), (ctx, r)
            assert np.array_equal(rungs["starts_in_left"][lo:hi].astype(bool), in_left), (ctx, r)
        ok = a["score"] >= max(1, min_score)
        t = int(a["score"][ok].max()) if ok.any() else 0
        sel = ks[(a["score"] == t) & in_left] if t > 0 else ks[:0]
        assert (int(top[r]), int(n_k[r]), int(sum_k[r])) == (t, len(sel), int(sel.sum())), (ctx, r)


def _ladder_case(rng, n_left, n_right, m, n_reads, kspan, qmode):
    left, right = _rand_seq(rng, n_left), _rand_seq(rng, n_right)
    motif = _rand_seq(rng, m)
    cores, kmin, kmax = [], [], []
    for _ in range(n_reads):
        k_true = rng.randint(0, kspan)
        lf = left[-rng.randint(0, min(n_left, 60)):] if n_left and rng.random() < 0.9 else ""
        rf = right[:rng.randint(0, min(n_right, 60))] if n_right and rng.random() < 0.9 else ""
        if qmode == "long":
            k_true += 200
        core = _mutate(rng, lf + motif * k_true + rf, rng.choice([0.0, 0.03, 0.12]))
        if qmode == "junk" or not core:
            core = _rand_seq(rng, rng.randint(1, 80))
        lo = max(0, k_true - rng.randint(0, 12))
        hi = k_true + rng.randint(0, 12)
        cores.append(core); kmin.append(lo); kmax.append(hi)
    return left, right, motif, cores, np.array(kmin, np.int32), np.array(kmax, np.int32)


@pytest.mark.parametrize("seed,n_left,n_right,m,n_reads,kspan,qmode", [
    (1, 40, 50, 3, 40, 20, "std"), (2, 0, 30, 2, 30, 10, "std"), (3, 25, 0, 4, 30, 10, "std"),
    (4, 0, 0, 5, 20, 8, "std"), (5, 300, 250, 1, 30, 30, "std"), (6, 120, 130, 6, 30, 60, "std"),
    (7, 60, 60, 3, 25, 15, "junk"), (8, 80, 90, 3, 12, 30, "long"), (9, 1000, 1000, 5, 16, 50, "std"),
    (10, 7, 9, 2, 40, 5, "std")])
def test_ladder_shared_sweeps_equal_independent_rectangles(engine, oracle, seed, n_left, n_right, m, n_reads, kspan, qmode):
    """Round 3 through the shared-sweep ladder kernel == every rung as its own rectangle == the oracle, on the full
    (score, tstart, tend) record of every rung (north_star: prefix sharing only if scores stay identical)."""
    rng = random.Random(1000 + seed)
    left, right, motif, cores, kmin, kmax = _ladder_case(rng, n_left, n_right, m, n_reads, kspan, qmode)
    sc = engine.get_preset("ont")
    ref, roff = oracle.align_ladders(cores, left, right, motif, kmin, kmax, n_threads=oracle.max_threads())
    got = {}
    try:
        for mode in (3, 2, 1, 0):
            engine.set_ladder_mode(mode)
            b = engine.Batch.round3(sc, left, right, motif, cores, kmin, kmax)
            b.run()
            if mode >= 2:
                _assert_flag_ladder(b, ref, roff, kmin, n_left, n_right, m, sc.min_dp_score,
                                    f"flag ladder mode {mode}, seed {seed}", rungs_too=mode == 2)
                with pytest.raises(engine.NanoRepeatB200Error):
                    b.fetch_alns()                      # no coordinates on flag words
                if mode == 3:
                    with pytest.raises(engine.NanoRepeatB200Error):
                        b.fetch_round3(want_rungs=True)  # no rung records from the paired kernel
            else:
                got[mode] = b.fetch_alns()
            b.close()
    finally:
        engine.set_ladder_mode(3)
    _assert_same(got[0], ref, f"independent rectangles, seed {seed}")
    _assert_same(got[1], ref, f"shared sweeps, seed {seed}")


def test_multi_region_batches_equal_per_region_calls(engine):
    """estimate_regions (one launch per round over all regions) == the two operators called region by region."""
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth
    regs = synth.config1(seed=11, n_regions=6, reads_per_region=12) + synth.config3(seed=12, n_loci=5, reads_per_locus=9)
    one = [nrb.RepeatRegion.from_synth(r) for r in regs]
    for reg, rr in zip(regs, one):
        nrb.round1_and_round2_estimation(reg.data_type, rr, 1)
        nrb.round3_estimation(reg.data_type, False, rr, 1)
    many = [nrb.RepeatRegion.from_synth(r) for r in regs]
    for reg, rr in zip(regs, many):
        rr.data_type = reg.data_type
    nrb.estimate_regions(many)
    n_r3 = 0
    for a, b in zip(one, many):
        for name in a.read_dict:
            ra, rb = a.read_dict[name], b.read_dict[name]
            assert (ra.round1_repeat_size, ra.round2_repeat_size) == (rb.round1_repeat_size, rb.round2_repeat_size)
            assert ra.round3_repeat_size == rb.round3_repeat_size and type(ra.round3_repeat_size) is type(rb.round3_repeat_size)
            n_r3 += ra.round3_repeat_size is not None
    assert n_r3 > 100


def test_sharded_pieces_equal_unsharded(engine):
    """Region splitting + dealing to 4 ranks (run one after the other here) gives every read the same numbers as the
    unsharded call: T is pinned region-wide, results do not depend on batch composition."""
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth, sharding
    regs = synth.config1(seed=21, n_regions=4, reads_per_region=11) + synth.config2(seed=22, n_reads=23)
    whole = [nrb.RepeatRegion.from_synth(r) for r in regs]
    nrb.estimate_regions(whole, "ont", False)
    parts = [nrb.RepeatRegion.from_synth(r) for r in regs]
    seen = []
    for rank in range(4):
        seen += sharding.estimate_regions_sharded(parts, "ont", False, max_reads_per_piece=5, rank=rank, world_size=4,
                                                  gather=False)
    assert sorted(seen) == list(range(len(seen))) and len(seen) > len(regs)
    for a, b in zip(whole, parts):
        for name in a.read_dict:
            ra, rb = a.read_dict[name], b.read_dict[name]
            assert (ra.round1_repeat_size, ra.round2_repeat_size, ra.round3_repeat_size) == \
                   (rb.round1_repeat_size, rb.round2_repeat_size, rb.round3_repeat_size), name


def test_ladder_long_expanded_allele(engine, oracle):
    """cfg4-like FMR1 shape: multi-stripe read, 1000-bp anchors, a +/-25 ladder around 500 units."""
    from nanorepeat_b200 import synth
    rng = np.random.default_rng(44)
    L, R = synth.random_seq(rng, 1000), synth.random_seq(rng, 1000)
    cores = [synth.simulate_core(rng, L, R, "CGG", k, "ont_r9")[0] for k in (480, 500, 523)]
    kmin = np.array([470, 490, 515], np.int32)
    kmax = np.array([486, 506, 530], np.int32)
    sc = engine.get_preset("ont")
    ref, roff = oracle.align_ladders(cores, L, R, "CGG", kmin, kmax, n_threads=oracle.max_threads())
    try:
        for mode in (3, 2, 1):
            engine.set_ladder_mode(mode)
            b = engine.Batch.round3(sc, L, R, "CGG", cores, kmin, kmax)
            b.run()
            if mode >= 2:
                _assert_flag_ladder(b, ref, roff, kmin, 1000, 1000, 3, sc.min_dp_score, "long flag ladder", rungs_too=mode == 2)
            else:
                _assert_same(b.fetch_alns(), ref, "long ladder")
            st = b.stats()
            assert st["executed_cells"] * 5 < st["algorithmic_cells"]
            b.close()
    finally:
        engine.set_ladder_mode(3)


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_golden_fixtures_through_operator_api(engine, name):
    """Reads like the reference's own use: fill a RepeatRegion, call the two operators, compare the attributes
    with what the reference's unmodified functions produced (tests/golden/make_golden.py)."""
    import nanorepeat_b200 as nrb
    doc = load_golden(name)
    for reg in doc["regions"]:
        rr = nrb.RepeatRegion()
        rr.left_anchor_seq, rr.right_anchor_seq = reg["left"], reg["right"]
        rr.left_anchor_len, rr.right_anchor_len = len(reg["left"]), len(reg["right"])
        rr.repeat_unit_seq = reg["motif"]
        for nme, core, dist in zip(reg["read_names"], reg["cores"], reg["dists"]):
            rr.read_dict[nme] = nrb.Read(nme, dist)
            rr.read_core_seq_dict[nme] = core
        nrb.round1_and_round2_estimation(reg["data_type"], rr, 1)
        nrb.round3_estimation(reg["data_type"], doc["fast_mode"], rr, 1)
        for nme, exp in zip(reg["read_names"], reg["expected"]):
            rd = rr.read_dict[nme]
            assert rd.round1_repeat_size == exp["r1"], (reg["name"], nme)
            assert rd.round2_repeat_size == exp["r2"], (reg["name"], nme)
            g3 = rd.round3_repeat_size
            assert (None if g3 is None else float(g3)) == exp["r3"], (reg["name"], nme, g3, exp["r3"])


def test_results_independent_of_batch_composition(engine):
    """Determinism clause of the boundary: a task's record does not depend on its neighbours."""
    rng = random.Random(31)
    qs = [_rand_seq(rng, rng.randint(10, 900)) for _ in range(50)]
    ts = [_rand_seq(rng, 30) + _mutate(rng, q, 0.1) + _rand_seq(rng, 30) for q in qs]
    sc = engine.get_preset("ont")
    full = engine.score_tasks(qs, ts, sc)
    perm = list(range(50))
    rng.shuffle(perm)
    part = engine.score_tasks([qs[i] for i in perm[:17]], [ts[i] for i in perm[:17]], sc)
    for j, i in enumerate(perm[:17]):
        assert tuple(part[j]) == tuple(full[i])


def test_long_expanded_allele_shape(engine, oracle):
    """cfg4-like: core of a few kb (many stripes) against a template of ~3.5 kb, with the round-trip property
    that a perfect read scores 2*|core| and spans the template exactly."""
    from nanorepeat_b200 import synth
    rng = np.random.default_rng(4)
    L, R = synth.random_seq(rng, 1000), synth.random_seq(rng, 1000)
    k = 500
    perfect = L[-100:] + "CGG" * k + R[:100]
    noisy, _, _ = synth.simulate_core(rng, L, R, "CGG", k, "ont_r9")
    tpl = L + "CGG" * k + R
    sc = engine.get_preset("ont")
    got = engine.score_tasks([perfect, noisy], [tpl, tpl], sc)
    assert tuple(int(v) for v in got[0]) == (2 * len(perfect), 900, 1000 + 3 * k + 100)
    ref = oracle.align_batch([perfect, noisy], [tpl, tpl], n_threads=2)
    _assert_same(got, ref, "long allele")


@pytest.mark.gpu
def test_round3_over_reused_round2_reads_equals_fresh_batches(engine):
    """nr_batch_begin_round3_from / add_round3_reuse (reads stay packed on the device) == a fresh round-3 batch."""
    from nanorepeat_b200 import synth
    regs = synth.config1(seed=11, n_regions=3, reads_per_region=12)
    sc = engine.get_preset("ont")
    b2 = engine.Batch.begin(sc, "round2")
    for reg in regs:
        b2.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, 60, reg.core_seqs)
    b2.commit().run().fetch_alns()
    rng = np.random.default_rng(5)
    b3 = engine.Batch.begin_round3_from(b2)
    fresh = engine.Batch.begin(sc, "round3")
    skip_all = []
    for i in (2, 0, 1):                                   # any order
        reg = regs[i]
        n = len(reg.core_seqs)
        kmin = np.array([max(0, k - 15) for k in reg.true_sizes], np.int32)
        kmax = np.array([k + 15 for k in reg.true_sizes], np.int32)
        skip = rng.random(n) < 0.25
        kmin[skip], kmax[skip] = 0, -1
        skip_all.append(skip)
        b3.add_round3_reuse(i, reg.right_anchor_seq, kmin, kmax)
        keep = np.flatnonzero(~skip)
        fresh.add_round3(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                         [reg.core_seqs[j] for j in keep], kmin[keep], kmax[keep])
    b2.close()                                            # the device pool must outlive its owner
    got = b3.commit().run().fetch_round3()
    exp = fresh.commit().run().fetch_round3()
    keep_all = ~np.concatenate(skip_all)
    for g, e in zip(got, exp):
        assert np.array_equal(g[keep_all], e)
        assert not g[~keep_all].any()
    b3.close(); fresh.close()


@pytest.mark.gpu
def test_operator_layer_drops_whitespace_like_the_fastq_round_trip(engine):
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth
    reg = synth.config1(seed=12, n_regions=1, reads_per_region=6)[0]
    clean = nrb.RepeatRegion.from_synth(reg)
    dirty = nrb.RepeatRegion.from_synth(reg)
    for n in list(dirty.read_core_seq_dict)[::2]:
        dirty.read_core_seq_dict[n] = dirty.read_core_seq_dict[n] + "\n"
    for rr in (clean, dirty):
        nrb.round1_and_round2_estimation("ont", rr, 1)
        nrb.round3_estimation("ont", False, rr, 1)
    for n in clean.read_dict:
        a, b = clean.read_dict[n], dirty.read_dict[n]
        assert (a.round2_repeat_size, a.round3_repeat_size) == (b.round2_repeat_size, b.round3_repeat_size)


@pytest.mark.parametrize("seed,n_left,m,T,n_reads,qmax", [(1, 40, 3, 30, 41, 200), (2, 0, 2, 12, 30, 90), (3, 300, 5, 40, 64, 520),
                                                          (4, 25, 4, 0, 17, 60), (5, 1000, 3, 80, 33, 700), (6, 5, 1, 7, 50, 40)])
def test_round2_flags_kind_equals_oracle(engine, oracle, seed, n_left, m, T, n_reads, qmax):
    """NR_KIND_ROUND2_FLAGS (paired u16x2 kernel for reads up to 512 bases, 32-bit kernel beside it for the rest):
    AS, tend and the span predicate tstart <= |left| (nanoRepeat_bam.py:373) against the oracle's records, and against
    the exact-record kind."""
    rng = random.Random(2000 + seed)
    left, motif = _rand_seq(rng, n_left), _rand_seq(rng, m)
    tpl = left + motif * T
    cores = []
    for i in range(n_reads):
        k = rng.randint(0, T + 3)
        kind = i % 5
        if kind == 0:
            core = _rand_seq(rng, rng.randint(1, min(qmax, 80)))                  # unrelated
        elif kind == 1:
            core = motif * max(1, k)                                              # no flank at all: starts past |left|
        elif kind == 2:
            core = _mutate(rng, left[-rng.randint(0, min(n_left, 100)):] if n_left else "", 0.05) + motif * k
        elif kind == 3:
            core = left[-1:] + motif * k if n_left else motif * k               # one flank base: ties on tstart
        else:
            core = _mutate(rng, (left[-60:] if n_left else "") + motif * k + _rand_seq(rng, 30), rng.choice([0.0, 0.1]))
        cores.append((core or "A")[:qmax])
    sc = engine.get_preset("ont")
    ref = oracle.align_batch(cores, [tpl] * n_reads, n_threads=oracle.max_threads())
    with engine.Batch.begin(sc, "round2_flags") as b:
        b.add_round2(left, motif, T, cores)
        score, tend, inside = b.commit().run().fetch_round2()
        with pytest.raises(engine.NanoRepeatB200Error):
            b.fetch_alns()
    with engine.Batch.begin(sc, "round2") as b:
        b.add_round2(left, motif, T, cores)
        exact = b.commit().run().fetch_alns()
    _assert_same(exact, ref, f"round-2 exact kind, seed {seed}")
    assert np.array_equal(score, ref["score"]), seed
    # tend and the predicate are exact wherever the span test (:373) can pass, i.e. tend >= |left|; an alignment that
    # ends before the repeat is reported as such (some tend < |left|), which is all the selection needs to drop it
    spans = ref["tend"] >= n_left
    assert np.array_equal(tend[spans], ref["tend"][spans]), seed
    assert (tend[~spans] < max(n_left, 1)).all(), seed
    live = (ref["score"] > 0) & spans
    assert np.array_equal(inside[live], (ref["tstart"] <= n_left)[live]), seed


def test_paired_ladder_undecidable_ties_are_rescored(engine, oracle):
    """Reads built so that a marked and an unmarked alignment tie for the top rung (no left-flank bases, one base, a
    mismatching flank): the paired kernel must hand them to the 32-bit ladder and the results must equal mode 2."""
    rng = random.Random(77)
    left, right, motif = _rand_seq(rng, 50), _rand_seq(rng, 60), "CAG"
    cores, kmin, kmax = [], [], []
    for i in range(48):
        k = rng.randint(1, 12)
        lf = ["", left[-1:], left[-2:], _rand_seq(rng, 3), left[-30:]][i % 5]
        rf = [right[:20], right[:1], "", right[:40]][i % 4]
        cores.append(lf + motif * k + rf)
        kmin.append(max(0, k - rng.randint(0, 4))); kmax.append(k + rng.randint(0, 4))
    kmin, kmax = np.array(kmin, np.int32), np.array(kmax, np.int32)
    sc = engine.get_preset("ont")
    sc.min_dp_score = 1
    ref, roff = oracle.align_ladders(cores, left, right, motif, kmin, kmax, n_threads=oracle.max_threads())
    try:
        for mode in (3, 2):
            engine.set_ladder_mode(mode)
            with engine.Batch.round3(sc, left, right, motif, cores, kmin, kmax) as b:
                b.run()
                _assert_flag_ladder(b, ref, roff, kmin, 50, 60, 3, 1, f"tie cases, mode {mode}", rungs_too=mode == 2)
    finally:
        engine.set_ladder_mode(3)


def test_paired_and_long_reads_in_one_batch(engine, oracle):
    """cfg2-like mix: most reads pair up (q <= 384), a few expanded alleles take the 32-bit kernels on the side stream;
    several regions in one batch; odd read counts leave a single read without a partner."""
    from nanorepeat_b200 import synth
    rng = np.random.default_rng(8)
    sc = engine.get_preset("ont")
    specs, refs = [], []
    for motif, ks in (("CAG", [17, 17, 55, 120, 150, 18, 54]), ("CCG", [7, 10, 7, 10, 9]), ("GGGGCC", [3, 80, 4])):
        L, R = synth.random_seq(rng, 1000), synth.random_seq(rng, 1000)
        cores = [synth.simulate_core(rng, L, R, motif, k, "ont")[0] for k in ks]
        kmin = np.array([max(0, k - 15) for k in ks], np.int32)
        kmax = np.array([k + 15 for k in ks], np.int32)
        specs.append((L, R, motif, cores, kmin, kmax))
        refs.append(oracle.align_ladders(cores, L, R, motif, kmin, kmax, n_threads=oracle.max_threads()))
    got = engine.round3_regions(sc, specs)
    pos = 0
    for (L, R, motif, cores, kmin, kmax), (ref, roff) in zip(specs, refs):
        for r in range(len(cores)):
            lo, hi = int(roff[r]), int(roff[r + 1])
            ks = np.arange(hi - lo) + int(kmin[r])
            a = ref[lo:hi]
            tlen = len(L) + len(motif) * ks + len(R)
            spans = (a["score"] > 0) & (tlen - a["tend"] < len(R)) & (a["tstart"] < len(L))
            ok = a["score"] >= sc.min_dp_score
            t = int(a["score"][ok].max()) if ok.any() else 0
            sel = ks[(a["score"] == t) & spans] if t > 0 else ks[:0]
            assert (int(got[2][pos]), int(got[1][pos]), int(got[0][pos])) == (t, len(sel), int(sel.sum())), (motif, r)
            pos += 1


def test_round3_resumed_from_round2_state_equals_fresh(engine, oracle):
    """The production flow (round 3 over a round-2 batch's reads, forward sweeps resumed from the DP state round 2 kept
    at the end of the left anchor) == fresh round-3 batches == the oracle's selection; reads round 2 gave no size keep
    their half of the pair empty."""
    from nanorepeat_b200 import synth
    rng = np.random.default_rng(17)
    sc = engine.get_preset("ont")
    regs = synth.config1(seed=31, n_regions=4, reads_per_region=15) + synth.config2(seed=32, n_reads=41)
    b2 = engine.Batch.begin(sc, "round2_flags")
    Ts = []
    for reg in regs:
        m = len(reg.repeat_unit_seq)
        T = int(max(d / m for d in reg.dist_between_anchors) * 1.5) + 1
        Ts.append(T)
        b2.add_round2(reg.left_anchor_seq, reg.repeat_unit_seq, T, reg.core_seqs)
    b2.commit().run().fetch_round2()
    b3 = engine.Batch.begin_round3_from(b2)
    fresh = engine.Batch.begin(sc, "round3")
    keep_all, refs = [], []
    for i, reg in enumerate(regs):
        n = len(reg.core_seqs)
        kmin = np.array([max(0, k - int(rng.integers(3, 16))) for k in reg.true_sizes], np.int32)
        kmax = np.array([k + int(rng.integers(3, 16)) for k in reg.true_sizes], np.int32)
        skip = rng.random(n) < 0.2
        kmin[skip], kmax[skip] = 0, -1
        b3.add_round3_reuse(i, reg.right_anchor_seq, kmin, kmax)
        keep = np.flatnonzero(~skip)
        keep_all.append(~skip)
        cores = [reg.core_seqs[j] for j in keep]
        fresh.add_round3(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq, cores, kmin[keep], kmax[keep])
        refs.append((oracle.align_ladders(cores, reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                                          kmin[keep], kmax[keep], n_threads=oracle.max_threads()), kmin[keep], reg))
    li = b3.commit().launch_info()
    assert li["n_pairs"] > 0
    got = b3.run().fetch_round3()
    exp = fresh.commit().run().fetch_round3()
    keep_all = np.concatenate(keep_all)
    for g, e in zip(got, exp):
        assert np.array_equal(g[keep_all], e)
        assert not g[~keep_all].any()
    pos = 0
    for (ref, roff), kmin, reg in refs:
        nl, nr_, m = len(reg.left_anchor_seq), len(reg.right_anchor_seq), len(reg.repeat_unit_seq)
        for r in range(len(kmin)):
            lo, hi = int(roff[r]), int(roff[r + 1])
            ks = np.arange(hi - lo) + int(kmin[r])
            a = ref[lo:hi]
            spans = (a["score"] > 0) & ((nl + m * ks + nr_) - a["tend"] < nr_) & (a["tstart"] < nl)
            ok = a["score"] >= sc.min_dp_score
            t = int(a["score"][ok].max()) if ok.any() else 0
            sel = ks[(a["score"] == t) & spans] if t > 0 else ks[:0]
            assert (int(exp[2][pos]), int(exp[1][pos]), int(exp[0][pos])) == (t, len(sel), int(sel.sum())), (reg.name, r)
            pos += 1
    b2.close(); b3.close(); fresh.close()


@pytest.mark.parametrize("rows", ["4", "6", "12", ""])
def test_many_long_reads_on_concurrent_stripes(engine, oracle, rows, monkeypatch):
    """Stress for the cooperative long-read path: dozens of multi-stripe tasks in flight at once, at every stripe
    height the host may pick (NR_COOP_ROWS pins it; "" = the host's own choice), exact records and ladders against the
    oracle.  A lost or stale boundary entry / token would show up as a wrong score here."""
    if rows:
        monkeypatch.setenv("NR_COOP_ROWS", rows)
    else:
        monkeypatch.delenv("NR_COOP_ROWS", raising=False)
    rng = random.Random(4242)
    qs, ts = [], []
    for i in range(48):
        ql = rng.randint(520, 2600)
        q = _rand_seq(rng, ql)
        t = _rand_seq(rng, rng.randint(0, 300)) + _mutate(rng, q, rng.choice([0.02, 0.1])) + _rand_seq(rng, rng.randint(0, 300))
        if i % 6 == 0:
            t = _rand_seq(rng, rng.randint(600, 3000))       # unrelated
        qs.append(q); ts.append(t)
    sc = engine.get_preset("ont")
    got = engine.score_tasks(qs, ts, sc)
    ref = oracle.align_batch(qs, ts, n_threads=oracle.max_threads())
    _assert_same(got, ref, f"long tasks, rows {rows!r}")
    # ladders: long reads of one region, modes 3 (long reads take the 32-bit flag ladder inside the fused launch) and 1
    left, right, motif = _rand_seq(rng, 300), _rand_seq(rng, 280), "GGGGCC"
    cores, kmin, kmax = [], [], []
    for i in range(20):
        k = rng.randint(70, 330)
        cores.append(_mutate(rng, left[-80:] + motif * k + right[:90], 0.06))
        kmin.append(max(0, k - rng.randint(3, 20))); kmax.append(k + rng.randint(3, 20))
    cores += [_mutate(rng, left[-50:] + motif * 9 + right[:60], 0.03) for _ in range(5)]      # short ones in the same batch
    kmin += [2] * 5; kmax += [20] * 5
    kmin, kmax = np.array(kmin, np.int32), np.array(kmax, np.int32)
    ref, roff = oracle.align_ladders(cores, left, right, motif, kmin, kmax, n_threads=oracle.max_threads())
    try:
        for mode in (3, 1):
            engine.set_ladder_mode(mode)
            with engine.Batch.round3(sc, left, right, motif, cores, kmin, kmax) as b:
                b.run()
                if mode == 3:
                    _assert_flag_ladder(b, ref, roff, kmin, 300, 280, 6, sc.min_dp_score, f"long ladders, rows {rows!r}", rungs_too=False)
                else:
                    _assert_same(b.fetch_alns(), ref, f"long ladders with coordinates, rows {rows!r}")
    finally:
        engine.set_ladder_mode(3)


def test_config2_full_size_properties(engine, oracle):
    """BASELINE.json's config 2 at full size (5 000 reads x 2 regions), through the operator API, by properties that do
    not need the oracle on every read: (1) planted error-free reads recover their repeat count exactly and a core that
    is its own template scores 2 * |core|; (2) the per-read results do not depend on the order of the reads (other
    pairs, other warps, other pipeline groups); (3) a random sample agrees with the CPU oracle's full selection."""
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth
    from oracle import selection
    regs = synth.config2(seed=2, n_reads=5000)
    rng = np.random.default_rng(99)
    cag = regs[0]
    planted = {}
    for j, k in enumerate([5, 17, 18, 40, 55, 56, 90, 121, 150, 33] * 3):
        name = f"perfect{j}"
        core = cag.left_anchor_seq[-100:] + "CAG" * k + cag.right_anchor_seq[:100]
        cag.read_names.append(name); cag.core_seqs.append(core); cag.dist_between_anchors.append(3 * k); cag.true_sizes.append(k)
        planted[name] = (k, core)

    def run(order_seed):
        rrs = []
        for reg in regs:
            rr = nrb.RepeatRegion.from_synth(reg)
            if order_seed is not None:
                names = list(rr.read_dict)
                np.random.default_rng(order_seed).shuffle(names)
                rr.read_dict = {n: rr.read_dict[n] for n in names}
            rrs.append(rr)
        nrb.estimate_regions(rrs, "ont", False)
        return rrs

    a, b = run(None), run(7)
    n_r3 = 0
    for ra, rb in zip(a, b):
        for name, rd in ra.read_dict.items():
            other = rb.read_dict[name]
            assert (rd.round1_repeat_size, rd.round2_repeat_size, rd.round3_repeat_size) == \
                   (other.round1_repeat_size, other.round2_repeat_size, other.round3_repeat_size), name
            n_r3 += rd.round3_repeat_size is not None
    assert n_r3 > 9900
    for name, (k, core) in planted.items():
        rd = a[0].read_dict[name]
        # round 2 has no right anchor to stop at (the first bases behind the repeat may extend it); round 3 is exact
        assert k <= rd.round2_repeat_size <= k + 4 and float(rd.round3_repeat_size) == float(k), (name, k, rd.round2_repeat_size, rd.round3_repeat_size)
    sc = engine.get_preset("ont")
    cores = [c for _k, c in planted.values()]
    tpls = [cag.left_anchor_seq + "CAG" * k + cag.right_anchor_seq for k, _c in planted.values()]
    got = engine.score_tasks(cores, tpls, sc)
    assert all(int(g["score"]) == 2 * len(c) and int(g["tstart"]) == 900 for g, c in zip(got, cores))
    # a sample of 24 reads per region against the oracle's rounds 1-3
    for reg, rr in zip(regs, a):
        idx = rng.choice(len(reg.read_names), 24, replace=False)
        # T is region-wide (nanoRepeat_bam.py:344): the oracle is told the region's longest distance between anchors
        exp = selection.estimate_region(reg.left_anchor_seq, reg.right_anchor_seq, reg.repeat_unit_seq,
                                        [reg.core_seqs[i] for i in idx], [reg.dist_between_anchors[i] for i in idx],
                                        n_threads=oracle.max_threads(), max_dist=max(reg.dist_between_anchors))
        for j, i in enumerate(idx):
            rd = rr.read_dict[reg.read_names[i]]
            g3 = rd.round3_repeat_size
            assert rd.round2_repeat_size == exp["r2"][j], (reg.name, i)
            assert (None if g3 is None else float(g3)) == (None if exp["r3"][j] is None else float(exp["r3"][j])), (reg.name, i)


@pytest.mark.parametrize("rows", ["4", "8", ""])
def test_round2_pairs_of_long_reads_equal_oracle(engine, oracle, rows, monkeypatch):
    """Round 2, flags kind, reads longer than one paired stripe (513 ... 3000 bases): two reads of a region share the
    u16x2 words stripe by stripe (cooperative stripes on different warps).  Several regions, odd counts, partners of
    very different lengths (those stay on the 32-bit kernels), short reads beside them; every stripe height."""
    if rows:
        monkeypatch.setenv("NR_COOP_ROWS", rows)
    else:
        monkeypatch.delenv("NR_COOP_ROWS", raising=False)
    rng = random.Random(515)
    sc = engine.get_preset("ont")
    specs, refs = [], []
    for g in range(4):
        n_left, m = rng.choice([60, 400, 1000]), rng.randint(2, 6)
        left, motif = _rand_seq(rng, n_left), _rand_seq(rng, m)
        T = rng.randint(300, 700)
        tpl = left + motif * T
        cores = []
        for i in range(rng.choice([7, 12, 15])):
            k = rng.choice([rng.randint(90, T), rng.randint(90, T), rng.randint(5, 60)])
            kind = i % 4
            core = left[-rng.randint(20, min(n_left, 120)):] + motif * k + _rand_seq(rng, rng.randint(0, 80))
            if kind == 1:
                core = _mutate(rng, core, 0.08)
            elif kind == 2:
                core = motif * k                                # no flank: starts past |left|
            elif kind == 3:
                core = _rand_seq(rng, rng.randint(520, 1500))  # unrelated long read
            cores.append(core[:3000])
        specs.append((left, motif, T, cores))
        refs.append(oracle.align_batch(cores, [tpl] * len(cores), n_threads=oracle.max_threads()))
    with engine.Batch.begin(sc, "round2_flags") as b:
        for left, motif, T, cores in specs:
            b.add_round2(left, motif, T, cores)
        b.commit()
        score, tend, inside = b.run().fetch_round2()
        score2, tend2, inside2 = b.run().fetch_round2()         # a second launch over the same scratch
    assert np.array_equal(score, score2) and np.array_equal(tend, tend2) and np.array_equal(inside, inside2)
    pos = 0
    for (left, motif, T, cores), ref in zip(specs, refs):
        n, n_left = len(cores), len(left)
        sl = slice(pos, pos + n)
        pos += n
        assert np.array_equal(score[sl], ref["score"]), (rows, [len(c) for c in cores])
        spans = ref["tend"] >= n_left
        assert np.array_equal(tend[sl][spans], ref["tend"][spans])
        assert (tend[sl][~spans] < max(n_left, 1)).all()
        live = (ref["score"] > 0) & spans
        assert np.array_equal(inside[sl][live], (ref["tstart"] <= n_left)[live])


def test_long_pairs_undecidable_ties_are_rescored(engine, oracle):
    """The same crafted ties as test_paired_ladder_undecidable_ties_are_rescored, on reads of 600 ... 1 000 bases: pairs
    of long reads run stripe by stripe on u16x2 words; a read whose selection hinges on a tie between a marked and an
    unmarked candidate comes back on the host's redo list and is rescored on 32-bit words.  Results == mode 2 == oracle."""
    rng = random.Random(78)
    left, right, motif = _rand_seq(rng, 50), _rand_seq(rng, 60), "CAG"
    cores, kmin, kmax = [], [], []
    for i in range(26):
        k = rng.randint(190, 310)
        lf = ["", left[-1:], left[-2:], _rand_seq(rng, 3), left[-30:]][i % 5]
        rf = [right[:20], right[:1], "", right[:40]][i % 4]
        cores.append(lf + motif * k + rf)
        kmin.append(max(0, k - rng.randint(0, 4))); kmax.append(k + rng.randint(0, 4))
    kmin, kmax = np.array(kmin, np.int32), np.array(kmax, np.int32)
    sc = engine.get_preset("ont")
    sc.min_dp_score = 1
    ref, roff = oracle.align_ladders(cores, left, right, motif, kmin, kmax, n_threads=oracle.max_threads())
    try:
        for mode in (3, 2):
            engine.set_ladder_mode(mode)
            with engine.Batch.round3(sc, left, right, motif, cores, kmin, kmax) as b:
                b.run()
                _assert_flag_ladder(b, ref, roff, kmin, 50, 60, 3, 1, f"long tie cases, mode {mode}", rungs_too=mode == 2)
                if mode == 3:
                    assert b.launch_info()["n_pairs"] == 0 and b.stats()["executed_cells"] > 0      # all reads in long pairs
                    b.run()                                                                        # and again over the same scratch
                    _assert_flag_ladder(b, ref, roff, kmin, 50, 60, 3, 1, "long tie cases, second launch", rungs_too=False)
    finally:
        engine.set_ladder_mode(3)
