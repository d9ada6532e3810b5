"""GPU parity of the joint path: nr_window_tasks / nr_joint_grid against the CPU oracle and the
golden vectors whose window scores the reference's own CIGAR re-scoring confirmed (tests/golden/make_golden_window.py)."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def _rs(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def _mut(rng, s, rate):
    out = []
    for ch in s:
        u = rng.random()
        if u < rate / 3:
            continue
        if u < 2 * rate / 3:
            out.append(rng.choice("ACGT")); continue
        if u < rate:
            out.append(rng.choice("ACGT"))
        out.append(ch)
    return "".join(out)


def _revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def test_window_tasks_equal_golden_vectors(engine):
    with open(os.path.join(GOLDEN, "joint_dp_cases.json")) as f:
        cases = json.load(f)["cases"]
    sc = engine.get_preset("ont")
    got = engine.window_tasks([c["query"] for c in cases], [c["target"] for c in cases], [c["win_a"] for c in cases],
                              [c["win_b"] for c in cases], sc, reverse=[c["reverse"] for c in cases])
    bad = [i for i, c in enumerate(cases) if (int(got["score"][i]), int(got["window_score"][i])) != (c["score"], c["window_score"])]
    assert not bad, (len(bad), bad[:5], [tuple(got[i]) for i in bad[:5]], [(cases[i]["score"], cases[i]["window_score"]) for i in bad[:5]])


def test_window_tasks_equal_oracle_long_reads_both_strands(engine, oracle):
    """Reads of many stripes (the joint CLI aligns whole amplicon reads, not cores), both strands, windows anywhere."""
    rng = random.Random(31)
    sc = engine.get_preset("ont")
    qs, ts, aa, bb, rv = [], [], [], [], []
    for it in range(60):
        L, R = _rs(rng, rng.choice([200, 1000])), _rs(rng, rng.choice([200, 1000]))
        m1, m2, mid = _rs(rng, 3), _rs(rng, 3), _rs(rng, 12)
        k1, k2 = rng.randint(0, 60), rng.randint(0, 20)
        tpl = L + m1 * k1 + mid + m2 * k2 + R
        read = _mut(rng, _rs(rng, rng.randint(0, 1500)) + L + m1 * (k1 + rng.randint(-3, 3) if k1 > 3 else k1) + mid + m2 * k2 + R +
                    _rs(rng, rng.randint(0, 1500)), rng.choice([0.02, 0.1]))
        reverse = it % 2
        qs.append(_revcomp(read) if reverse else read); ts.append(tpl); rv.append(reverse)
        a = max(len(L) - 10, 0); b = min(len(L) + 3 * k1 + 12 + 3 * k2 + 10, len(tpl))
        if it % 5 == 0:
            a, b = rng.randint(0, 100), rng.randint(len(tpl) - 100, len(tpl))
        aa.append(a); bb.append(b)
    got = engine.window_tasks(qs, ts, aa, bb, sc, reverse=rv)
    for i in range(len(qs)):
        exp = oracle.align_window(qs[i], ts[i], aa[i], bb[i], reverse=rv[i])
        assert (int(got["score"][i]), int(got["window_score"][i])) == exp, (i, tuple(got[i]), exp)
    # degenerate inputs
    z = engine.window_tasks(["", "ACGT"], ["ACGT", ""], [0, 0], [4, 0], sc)
    assert [tuple(int(v) for v in r) for r in z] == [(0, 0), (0, 0)]


def test_quantify_two_repeats_equals_oracle_driven_run(engine, oracle):
    """nanoRepeat-joint's rounds 2 and 3 for an HTT-like locus (CAG / CCG, README.md:167) through nr_joint_grid ==
    the same host logic fed by the CPU oracle, read by read; and the estimates land on the simulated alleles."""
    from nanorepeat_b200 import joint
    rng = random.Random(5)
    left, right, mid = _rs(rng, 1000), _rs(rng, 1000), "CAACAGCCGCCA"
    reads, truth, range1, range2 = [], [], [], []
    for i in range(24):
        k1, k2 = rng.choice([17, 55]), rng.choice([7, 10])
        amp = left[-rng.randint(150, 400):] + "CAG" * k1 + mid + "CCG" * k2 + right[:rng.randint(150, 400)]
        read = _mut(rng, amp, 0.05)
        reads.append(_revcomp(read) if i % 3 == 0 else read)
        truth.append((k1, k2))
        range1.append((max(0, k1 - rng.randint(8, 20)), k1 + rng.randint(8, 20)))        # what the initial estimate would give
        range2.append((max(0, k2 - rng.randint(4, 7)), k2 + rng.randint(4, 9)))
    range1[5] = None                                                                       # a read the initial estimate dropped

    def oracle_grid(sc, left, mid, right, motif1, motif2, reads_, pr, p1, p2):
        rec = np.zeros(len(pr), dtype=engine.WINDOW_DTYPE)
        strand = np.zeros(len(pr), dtype=np.uint8)
        for i, (r, k1, k2) in enumerate(zip(pr, p1, p2)):
            tpl = left + motif1 * k1 + mid + motif2 * k2 + right
            a, b = max(len(left) - 10, 0), min(len(left) + len(motif1) * k1 + len(mid) + len(motif2) * k2 + 10, len(tpl))
            f = oracle.align_window(reads_[r], tpl, a, b, reverse=False)
            v = oracle.align_window(reads_[r], tpl, a, b, reverse=True)
            best, strand[i] = (v, 1) if v > f else (f, 0)
            rec[i] = best
        return rec, strand

    got = joint.quantify_two_repeats(reads, left, mid, right, "CAG", "CCG", range1, range2, 200, 50)
    exp = joint.quantify_two_repeats(reads, left, mid, right, "CAG", "CCG", range1, range2, 200, 50, align=oracle_grid)
    assert got["step1"] == exp["step1"] and got["step2"] == exp["step2"]
    for i in range(len(reads)):
        assert (got["size1"][i] is None) == (exp["size1"][i] is None)
        if got["size1"][i] is not None:
            assert float(got["size1"][i]) == float(exp["size1"][i]) and float(got["size2"][i]) == float(exp["size2"][i]), i
    assert got["size1"][5] is None
    close = sum(abs(float(got["size1"][i]) - truth[i][0]) <= 2 and abs(float(got["size2"][i]) - truth[i][1]) <= 2
                for i in range(len(reads)) if got["size1"][i] is not None)
    assert close >= 18


def _grid_points(rng, n_reads, k1_lists, k2_lists):
    pr, p1, p2 = [], [], []
    for r in range(n_reads):
        for k1 in k1_lists[r]:
            for k2 in k2_lists[r]:
                pr.append(r); p1.append(k1); p2.append(k2)
    return pr, p1, p2


@pytest.mark.parametrize("n_left,n_right,mid", [(1000, 1000, "CAACAGCCGCCA"), (300, 10, "CAACAGCCGCCA"), (7, 40, "CAACAGCCGCCA"),
                                                (200, 200, ""), (150, 120, "T")])
def test_joint_grid_shared_sweeps_equal_rectangles(engine, oracle, n_left, n_right, mid):
    """nr_joint_grid scores full K1 x K2 grids with shared sweeps (nr_window_ladder.cuh: arithmetic K1 -> one prefix sweep and
    a continuation per k1, otherwise one forward sweep per k1; an empty mid with k2 from 0 takes the latter); every grid
    point must equal the rectangle of its own template (nr_window_tasks, both strands, better strand), and the CPU oracle on a sample."""
    rng = random.Random(1000 * n_left + n_right)
    sc = engine.get_preset("ont")
    left, right = _rs(rng, n_left), _rs(rng, n_right)
    m1, m2 = "CAG", "CCG"
    reads, K1s, K2s = [], [], []
    for i in range(36):
        k1, k2 = rng.choice([17, 55, 3]), rng.choice([7, 10, 0])
        amp = left[-rng.randint(1, 400):] + m1 * k1 + mid + m2 * k2 + right[:rng.randint(1, 400)]
        read = _mut(rng, amp, rng.choice([0.02, 0.08, 0.15]))
        if i % 7 == 3:
            read = read[:len(read) // 2] + "N" + read[len(read) // 2 + 1:]
        if i % 11 == 5:
            read = _rs(rng, rng.randint(1, 700))                                    # unrelated read
        reads.append(_revcomp(read) if i % 2 else read)
        lo1, step1 = max(0, k1 - rng.randint(0, 12)), rng.choice([1, 1, 2, 5])
        lo2, step2 = max(0, k2 - rng.randint(0, 6)), rng.choice([1, 1, 2, 3])
        K1s.append([lo1 + step1 * j for j in range(rng.randint(1, 6))] if i % 5 else [lo1, lo1 + 1, lo1 + 7])
        K2s.append([lo2 + step2 * j for j in range(rng.randint(1, 9))])
    K2s[4] = [2, 3, 7]                                                               # not arithmetic: template by template
    pr, p1, p2 = _grid_points(rng, len(reads), K1s, K2s)
    pr += [0, 0]; p1 += [91, 92]; p2 += [1, 5]                                       # read 0 is no longer a full grid
    got, strand = engine.joint_grid(sc, left, mid, right, m1, m2, reads, pr, p1, p2)
    qs, ts, aa, bb, rv = [], [], [], [], []
    for r, k1, k2 in zip(pr, p1, p2):
        tpl = left + m1 * k1 + mid + m2 * k2 + right
        a, b = max(n_left - 10, 0), min(n_left + 3 * k1 + len(mid) + 3 * k2 + 10, len(tpl))
        for rev in (0, 1):
            qs.append(reads[r]); ts.append(tpl); aa.append(a); bb.append(b); rv.append(rev)
    rect = engine.window_tasks(qs, ts, aa, bb, sc, reverse=rv)
    bad = []
    for i in range(len(pr)):
        f = (int(rect["score"][2 * i]), int(rect["window_score"][2 * i]))
        v = (int(rect["score"][2 * i + 1]), int(rect["window_score"][2 * i + 1]))
        exp, exp_strand = (v, 1) if v > f else (f, 0)
        if (int(got["score"][i]), int(got["window_score"][i])) != exp or int(strand[i]) != exp_strand:
            bad.append((i, pr[i], p1[i], p2[i], tuple(int(x) for x in got[i]), int(strand[i]), exp, exp_strand))
    assert not bad, (len(bad), len(pr), bad[:8])
    for i in rng.sample(range(len(pr)), 25):
        tpl = left + m1 * p1[i] + mid + m2 * p2[i] + right
        a, b = max(n_left - 10, 0), min(n_left + 3 * p1[i] + len(mid) + 3 * p2[i] + 10, len(tpl))
        f = oracle.align_window(reads[pr[i]], tpl, a, b, reverse=False)
        v = oracle.align_window(reads[pr[i]], tpl, a, b, reverse=True)
        assert (int(got["score"][i]), int(got["window_score"][i])) == max(f, v), (i, tuple(got[i]), f, v)
