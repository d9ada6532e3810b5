"""The joint path's rescoring / selection restatement (oracle/joint.py) against golden vectors
produced by the reference's own functions (tests/golden/make_golden_joint.py)."""
import json
import os

import pytest

from oracle import joint

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_window_stats_equal_reference():
    doc = _load("joint_window_cases.json")
    assert len(doc["cases"]) > 400
    for c in doc["cases"]:
        got = joint.window_stats(c["cigar"], c["tstart"], c["tend"], c["a"], c["b"])
        assert got == c["expected"], (c["cigar"], c["tstart"], c["a"], c["b"], c["note"], got, c["expected"])


def test_known_window_answers():
    # SURVEY.md section 4's probe and two hand-derived ones: +2 / -4 / gap -4 - 2 (l - 1), window edges as tk.py has them
    assert joint.window_stats("100=", 0, 100, 10, 90)["score"] == 160
    assert joint.window_stats("50=20D50=", 0, 120, 60, 100) == dict(num_match=30, num_mismatch=0, num_ins=0, num_del=10, score=60 - 4 - 2 * 9)
    assert joint.window_stats("50=7I50=", 0, 100, 50, 100)["num_ins"] == 0          # insertion AT the window start: not counted
    assert joint.window_stats("50=7I50=", 0, 100, 49, 100)["score"] == 2 * 51 - 4 - 2 * 6
    with pytest.raises(ValueError):
        joint.window_stats("", 0, 0, 0, 1)
    with pytest.raises(ValueError):
        joint.window_stats("10=3Q", 0, 10, 0, 5)


def test_two_repeat_selection_equals_reference():
    doc = _load("joint_selection_cases.json")
    for c in doc["cases"]:
        recs = []
        for line in c["paf_lines"]:
            col = line.split("\t")
            k1, k2 = (int(x) for x in col[5].split("-"))
            cigar = [x[5:] for x in col[12:] if x.startswith("cg:Z:")][0]
            recs.append((col[0], k1, k2, int(col[6]), int(col[7]), int(col[8]), cigar))
        got = joint.two_repeat_sizes(recs, c["left_len"], c["mid_len"], c["m1"], c["m2"])
        assert {q: list(v) for q, v in got.items()} == c["expected"]


def test_window_dp_equals_golden_and_the_reference_rescoring_of_its_own_cigar():
    """oracle/nr_oracle.c nro_align_window (score, window score carried through the DP) on the committed cases: equal to
    the recorded values, equal with and without the traceback, and the recorded CIGAR re-scored by the restatement of
    tk.target_region_alignment_stats_from_cigar (itself pinned above against the reference) gives the window score.
    (At generation time the reference's own function was run on the same CIGARs: tests/golden/make_golden_window.py.)"""
    from oracle import nr_oracle
    nr_oracle.build()
    doc = _load("joint_dp_cases.json")
    assert len(doc["cases"]) >= 150
    for c in doc["cases"]:
        got = nr_oracle.align_window(c["query"], c["target"], c["win_a"], c["win_b"], reverse=c["reverse"], want_cigar=True)
        assert got == (c["score"], c["window_score"], c["tstart"], c["tend"], c["cigar"])
        assert nr_oracle.align_window(c["query"], c["target"], c["win_a"], c["win_b"], reverse=c["reverse"]) == (c["score"], c["window_score"])
        if c["score"] > 0:
            assert joint.window_stats(c["cigar"], c["tstart"], c["tend"], c["win_a"], c["win_b"])["score"] == c["window_score"]


def test_grid_rounds_equal_the_reference_pipeline_end_to_end():
    """nanorepeat_b200.joint.quantify_two_repeats (grid enumeration, step sizes, windows, tie selection), fed by the CPU
    oracle's DP, returns per read exactly what the reference's own fine_tune_read_count returned when its pyminimap2
    was replaced by a PAF printer over the same DP (tests/golden/make_golden_joint_pipeline.py): temp files, PAF text and
    CIGAR re-scoring on one side, binary records with the window score carried through the DP on the other."""
    import numpy as np
    from nanorepeat_b200 import engine, joint as njoint
    from oracle import nr_oracle
    nr_oracle.build()
    doc = _load("joint_pipeline_cases.json")
    assert len(doc["cases"]) >= 6

    def oracle_grid(sc, left, mid, right, motif1, motif2, reads_, pr, p1, p2):
        rec = np.zeros(len(pr), dtype=engine.WINDOW_DTYPE)
        strand = np.zeros(len(pr), dtype=np.uint8)
        for i, (r, k1, k2) in enumerate(zip(pr, p1, p2)):
            tpl = left + motif1 * int(k1) + mid + motif2 * int(k2) + right
            a, b = max(len(left) - 10, 0), min(len(left) + len(motif1) * int(k1) + len(mid) + len(motif2) * int(k2) + 10, len(tpl))
            f = nr_oracle.align_window(reads_[int(r)], tpl, a, b, reverse=False)
            v = nr_oracle.align_window(reads_[int(r)], tpl, a, b, reverse=True)
            best, strand[i] = (v, 1) if v > f else (f, 0)
            rec[i] = best
        return rec, strand

    rounds3 = 0
    for c in doc["cases"]:
        r1 = [None if x is None else tuple(x) for x in c["range1"]]
        r2 = [None if x is None else tuple(x) for x in c["range2"]]
        got = njoint.quantify_two_repeats(c["reads"], c["left"], c["mid"], c["right"], c["motif1"], c["motif2"], r1, r2,
                                          c["max_size1"], c["max_size2"], align=oracle_grid)
        assert (got["step1"], got["step2"]) == (c["step1"], c["step2"])
        assert [None if v is None else float(v) for v in got["size1"]] == c["size1"]
        assert [None if v is None else float(v) for v in got["size2"]] == c["size2"]
        rounds3 += c["alignment_calls"] > 100
        close = sum(v is not None and abs(v - t[0]) <= 2 for v, t in zip(c["size1"], c["truth"]))
        assert close >= 7, (close, c["size1"], c["truth"])
    assert rounds3 >= 4
