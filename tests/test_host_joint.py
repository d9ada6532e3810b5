"""Host logic of the joint path (nanorepeat_b200/joint.py, vectorised) against plain-loop restatements of the reference's
loops (nanoRepeat_joint.py:296-333, :397-410, :427-478) on seeded cases.  No GPU, no library call."""
import random

import numpy as np

from nanorepeat_b200 import engine, joint


def _loop_round2(range1, range2, min1, max1, min2, max2, step1, step2):
    pr, p1, p2 = [], [], []
    for k1 in range(min1, max1 + 1, step1):
        for k2 in range(min2, max2 + 1, step2):
            for r, (a, b) in enumerate(zip(range1, range2)):
                if a is None or b is None:
                    continue
                if a[0] <= k1 < a[1] and b[0] <= k2 < b[1]:
                    pr.append(r); p1.append(k1); p2.append(k2)
    return pr, p1, p2


def _loop_round3(range1, range2, size1, size2, buffer1, buffer2):
    s1 = [v for v, w in zip(size1, size2) if v is not None and w is not None]
    s2 = [w for v, w in zip(size1, size2) if v is not None and w is not None]
    if not s1:
        return [], [], []
    min_size1, max_size1 = max(int(min(s1) - buffer1), 0), int(max(s1) + buffer1 + 2)
    min_size2, max_size2 = max(int(min(s2) - buffer2), 0), int(max(s2) + buffer2 + 2)
    pr, p1, p2 = [], [], []
    for k1 in range(min_size1, max_size1):
        for k2 in range(min_size2, max_size2):
            for r, (v, w) in enumerate(zip(size1, size2)):
                if v is None or w is None:
                    continue
                if k1 < v - buffer1 or k1 >= v + buffer1 or k2 < w - buffer2 or k2 >= w + buffer2:
                    continue
                a, b = range1[r], range2[r]
                if k1 < a[0] or k1 >= a[1] or k2 < b[0] or k2 >= b[1]:
                    continue
                pr.append(r); p1.append(k1); p2.append(k2)
    return pr, p1, p2


def _loop_estimate(n_reads, point_read, point_k1, point_k2, records, min_dp_score=80):
    size1, size2 = [None] * n_reads, [None] * n_reads
    per_read = {}
    for r, k1, k2, rec in zip(point_read, point_k1, point_k2, records):
        if rec["score"] <= 0 or rec["score"] < min_dp_score:
            continue
        per_read.setdefault(int(r), []).append((int(rec["window_score"]), int(k1), int(k2)))
    for r, rows in per_read.items():
        top = max(s for s, _a, _b in rows)
        size1[r] = np.mean([k1 for s, k1, _k2 in rows if s == top])
        size2[r] = np.mean([k2 for s, _k1, k2 in rows if s == top])
    return size1, size2


def _same(a, b):
    return [list(map(int, x)) for x in a] == [list(map(int, x)) for x in b]


def test_grid_points_and_selection_equal_the_loops():
    rng = random.Random(3)
    for case in range(40):
        n = rng.randint(1, 60)
        range1 = [None if rng.random() < 0.1 else (lo, lo + rng.randint(0, 30)) for lo in (rng.randint(0, 60) for _ in range(n))]
        range2 = [None if rng.random() < 0.1 else (lo, lo + rng.randint(0, 12)) for lo in (rng.randint(0, 20) for _ in range(n))]
        ok1, ok2 = [a for a in range1 if a], [b for b in range2 if b]
        if not ok1 or not ok2:
            continue
        min1, max1 = min(a for a, _ in ok1), max(b for _, b in ok1)
        min2, max2 = min(a for a, _ in ok2), max(b for _, b in ok2)
        step1, step2 = rng.randint(1, 6), rng.randint(1, 4)
        got = joint.round2_grid_points(range1, range2, min1, max1, min2, max2, step1, step2)
        exp = _loop_round2(range1, range2, min1, max1, min2, max2, step1, step2)
        assert _same(got, exp), case
        pr, p1, p2 = exp
        rec = np.zeros(len(pr), dtype=engine.WINDOW_DTYPE)
        rec["score"] = [rng.choice([0, 50, 79, 80, 300, 900]) for _ in pr]
        rec["window_score"] = [rng.randint(-40, 40) // 4 * 4 for _ in pr]            # many ties
        g1, g2 = joint.estimate_two_repeats(n, got[0], got[1], got[2], rec, 80)
        e1, e2 = _loop_estimate(n, pr, p1, p2, rec, 80)
        assert [None if v is None else float(v) for v in g1] == [None if v is None else float(v) for v in e1], case
        assert [None if v is None else float(v) for v in g2] == [None if v is None else float(v) for v in e2], case
        got3 = joint.round3_grid_points(range1, range2, g1, g2, step1, step2)
        exp3 = _loop_round3(range1, range2, e1, e2, step1, step2)
        assert _same(got3, exp3), case
    assert _same(joint.round3_grid_points([(0, 5)], [(0, 5)], [None], [None], 2, 2), ([], [], []))
    assert joint.estimate_two_repeats(2, [], [], [], np.zeros(0, dtype=engine.WINDOW_DTYPE)) == ([None, None], [None, None])
