#!/usr/bin/env python
"""Golden vectors for the joint path's alignment-with-window-score DP.

Runs only in the build container (needs /root/reference).  For seeded (read, template, window) cases of the joint
CLI's shape -- left + motif1*k1 + mid + motif2*k2 + right, window = both repeats +- 10 (nanoRepeat_joint.py:448-451) --
the oracle (oracle/nr_oracle.c, nro_align_window_cigar) computes the canonical optimal alignment, its score, its window
score carried through the DP as a payload, and the alignment's CIGAR from a traceback.  The REFERENCE's own, unmodified
tk.target_region_alignment_stats_from_cigar (src/NanoRepeat/tk.py:435-500) is then run on that CIGAR: its .score must
equal the payload, which pins the payload arithmetic (what the CUDA kernel computes without any CIGAR) to the reference's
re-scoring rules.  The alignment engine itself stays PARITY UNPINNED (pyminimap2 is absent).

Usage: python tests/golden/make_golden_window.py   -> tests/golden/joint_dp_cases.json
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import nr_oracle, joint          # noqa: E402
import make_golden_joint                     # noqa: E402

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def main():
    tk, _ = make_golden_joint.import_reference()
    rng = random.Random(20261018)

    def rs(n):
        return "".join(rng.choice("ACGT") for _ in range(n))

    def mut(s, rate):
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

    cases = []
    for it in range(160):
        big = it % 8 == 0                                  # a few with 1000-bp anchors and reads of several stripes
        nl, nr_ = (1000, 1000) if big else (rng.randint(30, 200), rng.randint(30, 200))
        L, R, mid = rs(nl), rs(nr_), rs(rng.choice([0, 5, 12]))
        m1, m2 = rs(rng.randint(2, 6)), rs(rng.randint(2, 6))
        k1, k2 = rng.randint(0, 40 if big else 15), rng.randint(0, 15)
        tpl = L + m1 * k1 + mid + m2 * k2 + R
        t1, t2 = max(0, k1 + rng.randint(-4, 4)), max(0, k2 + rng.randint(-3, 3))
        read = mut(L[-rng.randint(10, min(nl, 400)):] + m1 * t1 + mid + m2 * t2 + R[:rng.randint(10, min(nr_, 400))],
                   rng.choice([0.0, 0.05, 0.15]))
        if it % 11 == 0:
            read = read[:len(read) // 2] + "N" + read[len(read) // 2:]
        if it % 13 == 0:
            read = rs(rng.randint(20, 150))                # unrelated read
        reverse = it % 3 == 0
        query = "".join(COMP[c] for c in reversed(read)) if reverse else read
        a, b = joint.two_repeat_window(nl, len(mid), len(m1), len(m2), k1, k2, len(tpl))
        if it % 7 == 0:
            a, b = rng.randint(0, len(tpl) // 2), rng.randint(len(tpl) // 2, len(tpl))
        score, wscore, ts, te, cigar = nr_oracle.align_window(query, tpl, a, b, reverse=reverse, want_cigar=True)
        assert score == nr_oracle.align(read, tpl)[0]
        if score > 0:
            ref = tk.target_region_alignment_stats_from_cigar(cigar, ts, te, a, b).score
            assert ref == wscore, (it, ref, wscore, cigar)
        cases.append(dict(query=query, target=tpl, win_a=a, win_b=b, reverse=int(reverse), score=score, window_score=wscore,
                          tstart=ts, tend=te, cigar=cigar))
    with open(os.path.join(HERE, "joint_dp_cases.json"), "w") as f:
        json.dump(dict(source="oracle nro_align_window_cigar; window_score == reference tk.target_region_alignment_stats_from_cigar(cigar).score "
                              "asserted at generation time", cases=cases), f, indent=0, separators=(",", ":"))
    print(len(cases), "cases,", sum(c["score"] > 0 for c in cases), "with an alignment")


if __name__ == "__main__":
    main()
