"""CPU restatement of the joint path's CIGAR-window rescoring and two-repeat selection (TEST INFRASTRUCTURE ONLY).

SURVEY.md section 8a, row a7 -- the part of nanoRepeat-joint that sits behind the alignment engine:

    window_stats         <- reference src/NanoRepeat/tk.py:435-500  target_region_alignment_stats_from_cigar
    two_repeat_window    <- reference src/NanoRepeat/nanoRepeat_joint.py:448-452 (the window of a grid point)
    two_repeat_sizes     <- reference src/NanoRepeat/nanoRepeat_joint.py:427-478 estimate_two_repeats_from_paf

Pinned: tests/golden/joint_window_cases.json and joint_selection_cases.json hold what the reference's own, unmodified
functions return on 411 + 12 seeded cases (tests/golden/make_golden_joint.py); tests/test_oracle_joint.py compares.
The device side of a7 (a DP that carries the window score of the optimal path) is not built yet;
this module is what it will be held against.
"""
import re

import numpy as np

_RUN = re.compile(r"(\d+)([=XIDNSHPM])")

MATCH, MISMATCH, GAP_OPEN, GAP_EXT = 2, -4, -4, -2          # tk.py:444-447


def parse_cigar(cigar):
    """-> list of (length, op).  tk.py:378-398 (unknown characters are an error there; here too)."""
    runs = [(int(n), op) for n, op in _RUN.findall(cigar)]
    if sum(len(str(n)) + 1 for n, _ in runs) != len(cigar):
        raise ValueError(f"unknown CIGAR operation in {cigar!r}")
    return runs


def _overlap(lo, hi, a, b):
    return max(0, min(hi, b) - max(lo, a))                   # tk.py:366-371


def window_stats(cigar, tstart, tend, a, b):
    """Matches, mismatches, inserted and deleted bases and the score of the alignment inside target window [a, b).

    tk.py:435-500.  Scored per run: '=' +2 and 'X' -4 per base inside the window; a deletion by the part of it inside
    the window, -4 - 2 (part - 1); an insertion in full, -4 - 2 (len - 1), when its target position p satisfies
    a < p < b - 1 (both strict, :477); walking stops once the position has passed b (:490); window bases the
    alignment does not reach on either side count as mismatches but do not change the score (:492-496)."""
    if not cigar:
        raise ValueError("cigar string is empty")
    n_match = n_mis = n_ins = n_del = score = 0
    pos = tstart
    for length, op in parse_cigar(cigar):
        if op == "=":
            inside = _overlap(pos, pos + length, a, b)
            n_match += inside
            score += MATCH * inside
            pos += length
        elif op == "X":
            inside = _overlap(pos, pos + length, a, b)
            n_mis += inside
            score += MISMATCH * inside
            pos += length
        elif op == "I":
            if a < pos < b - 1:
                n_ins += length
                score += GAP_OPEN + GAP_EXT * (length - 1)
        elif op == "D":
            inside = _overlap(pos, pos + length, a, b)
            if inside:
                n_del += inside
                score += GAP_OPEN + GAP_EXT * (inside - 1)
            pos += length
        elif op == "S":
            continue
        else:
            raise ValueError(f"unsupported cigar operation: {op}")
        if pos > b:
            break
    if tend < b:
        n_mis += b - tend
    if tstart > a:
        n_mis += tstart - a
    return dict(num_match=n_match, num_mismatch=n_mis, num_ins=n_ins, num_del=n_del, score=score)


def two_repeat_window(left_len, mid_len, m1, m2, k1, k2, tlen):
    """The window of grid point (k1, k2): both repeats, the piece between them and 10 bases on either side, clipped to
    the template (nanoRepeat_joint.py:448-451)."""
    return max(left_len - 10, 0), min(left_len + m1 * k1 + mid_len + m2 * k2 + 10, tlen)


def two_repeat_sizes(records, left_len, mid_len, m1, m2):
    """records: iterable of (qname, k1, k2, tlen, tstart, tend, cigar) -- one per PAF line, in file order.
    -> {qname: (size1, size2)}: per read the grid points whose window score is the highest; the two sizes are the
    means of their k1 and of their k2, separately (nanoRepeat_joint.py:457-476; the sort is stable and only the top
    group is read, so file order does not matter)."""
    per_read = {}
    for qname, k1, k2, tlen, tstart, tend, cigar in records:
        a, b = two_repeat_window(left_len, mid_len, m1, m2, k1, k2, tlen)
        per_read.setdefault(qname, []).append((window_stats(cigar, tstart, tend, a, b)["score"], k1, k2))
    out = {}
    for qname, rows in per_read.items():
        top = max(s for s, _, _ in rows)
        out[qname] = (float(np.mean([k1 for s, k1, _ in rows if s == top])),
                      float(np.mean([k2 for s, _, k2 in rows if s == top])))
    return out
