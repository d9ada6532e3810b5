"""CPU suite: the oracle against hand-derived vectors, against its own two restatements, and the selection
restatement against fixtures produced by the reference's unmodified functions (tests/golden/make_golden.py)."""
import ctypes
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN_DIR, GOLDEN_SETS, load_golden


def expand(expr, doc):
    return "".join(doc[p] if p in ("A", "B") else p for p in expr.split("+"))


def micro_cases():
    with open(os.path.join(GOLDEN_DIR, "micro_cases.json")) as f:
        doc = json.load(f)
    return [(c["name"], expand(c["query"], doc), expand(c["target"], doc), tuple(c["expected"])) for c in doc["cases"]]


@pytest.mark.parametrize("name,query,target,expected", micro_cases(), ids=[c[0] for c in micro_cases()])
def test_micro_known_answers(oracle, name, query, target, expected):
    assert oracle.align(query, target) == expected
    assert oracle.align_py(query, target) == expected


def test_three_restatements_agree(oracle):
    L = oracle.lib()
    L.nro_align_tuple.argtypes = L.nro_align.argtypes
    sc = oracle.scoring()
    rng = random.Random(11)
    for it in range(600):
        n, m = rng.randint(1, 50), rng.randint(1, 80)
        alpha = "ACGT" if it % 5 else "ACGTN"
        q = "".join(rng.choice(alpha) for _ in range(n))
        t = "".join(rng.choice(alpha) for _ in range(m))
        if it % 2:
            t = t[:m // 2] + q[:n // 2] + "A" * rng.randint(0, 30) + q[n // 2:] + t[m // 2:]
        out = np.zeros(1, dtype=oracle.ALN_DTYPE)
        L.nro_align_tuple(ctypes.byref(sc), q.encode(), len(q), t.encode(), len(t), out.ctypes.data)
        tup = tuple(int(x) for x in out[0])
        fast = oracle.align(q, t)
        assert fast == tup, (q, t)
        if it < 150:
            assert fast == oracle.align_py(q, t), (q, t)


def test_empty_and_batch(oracle):
    assert oracle.align("", "ACGT") == (0, 0, 0)
    assert oracle.align("ACGT", "") == (0, 0, 0)
    out = oracle.align_batch([], [])
    assert len(out) == 0
    out = oracle.align_batch(["ACGT", "AC"], ["ACGT", "GGACGG"], n_threads=2)
    assert [tuple(int(v) for v in r) for r in out] == [(8, 0, 4), (4, 2, 4)]


def test_ladder_matches_independent_tasks(oracle):
    rng = random.Random(5)
    left = "".join(rng.choice("ACGT") for _ in range(120))
    right = "".join(rng.choice("ACGT") for _ in range(90))
    motif = "CAG"
    cores = [left[-40:] + motif * k + right[:35] for k in (3, 7)]
    out, off = oracle.align_ladders(cores, left, right, motif, [0, 2], [9, 12], n_threads=2)
    for r, (lo, hi) in enumerate([(0, 9), (2, 12)]):
        for k in range(lo, hi + 1):
            exp = oracle.align(cores[r], left + motif * k + right)
            got = tuple(int(v) for v in out[off[r] + k - lo])
            assert got == exp


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_selection_restatement_matches_reference_fixtures(oracle, name):
    """oracle.selection (our restatement of nanoRepeat_bam.py:334-347,:373-384,:463-472,:423-433) must give the
    numbers the reference's own functions produced for the same inputs."""
    from oracle import selection
    doc = load_golden(name)
    sc = oracle.scoring(**doc["scoring"])
    for reg in doc["regions"]:
        res = selection.estimate_region(reg["left"], reg["right"], reg["motif"], reg["cores"], reg["dists"],
                                        fast_mode=doc["fast_mode"], sc=sc, n_threads=oracle.max_threads())
        for i, exp in enumerate(reg["expected"]):
            assert res["r1"][i] == exp["r1"]
            assert res["r2"][i] == exp["r2"]
            r3 = res["r3"][i]
            assert (None if r3 is None else float(r3)) == exp["r3"], (reg["name"], i)


def test_ladder_bounds_truncation():
    from oracle import selection
    assert selection.ladder_bounds(0.0) == (0, 15)
    assert selection.ladder_bounds(7.666666666666667) == (0, 22)
    assert selection.ladder_bounds(20.9) == (5, 35)
    assert selection.ladder_bounds(400.5) == (380, 420)          # buffer = int(20.025) = 20
    assert selection.ladder_bounds(5000.0) == (4850, 5150)       # buffer capped at 150
    assert selection.ladder_bounds(5000.0, fast_mode=True) == (4985, 5015)
    assert selection.ladder_bounds(-0.6) == (0, 14)              # int() truncates toward zero; kmin clamps


def test_round1_template_size():
    from oracle import selection
    r1, T = selection.round1([60, 30], 3)
    assert r1 == [20.0, 10.0] and T == 31
    r1, T = selection.round1([9], 3)        # 3.0*1.5+1 = 5 < 13 -> int(13.0)
    assert T == 13
    r1, T = selection.round1([-3], 5)       # negative dist allowed (> -10): max r1 = -0.6 -> T = int(9.4) = 9
    assert T == 9
