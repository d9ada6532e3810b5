#!/usr/bin/env python
"""Golden vectors for the joint path's CIGAR-window rescoring and two-repeat selection.

Runs only in the build container (needs /root/reference).  Calls the REFERENCE's own, unmodified
tk.target_region_alignment_stats_from_cigar (src/NanoRepeat/tk.py:435-500) and
nanoRepeat_joint.estimate_two_repeats_from_paf (src/NanoRepeat/nanoRepeat_joint.py:427-478) on seeded random and
hand-made inputs and records their outputs in tests/golden/joint_window_cases.json / joint_selection_cases.json.
The third-party imports the reference makes at module level (pyminimap2, pysam, Levenshtein, matplotlib, distutils)
are stubbed: neither function touches them.

Usage: python tests/golden/make_golden_joint.py
"""
import json
import os
import random
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")


def import_reference():
    for name in ("pysam", "Levenshtein", "pyminimap2"):
        sys.modules.setdefault(name, types.ModuleType(name))
    du = types.ModuleType("distutils"); du.spawn = types.ModuleType("distutils.spawn")
    sys.modules.setdefault("distutils", du); sys.modules.setdefault("distutils.spawn", du.spawn)
    mpl = types.ModuleType("matplotlib"); mpl.__path__ = []; mpl.use = lambda *a, **k: None; mpl.rcParams = {}
    for sub, attrs in (("pyplot", ()), ("colors", ("Normalize",)), ("cm", ()), ("patches", ())):
        mod = types.ModuleType("matplotlib." + sub)
        for a in attrs:
            setattr(mod, a, object)
        setattr(mpl, sub, mod)
        sys.modules["matplotlib." + sub] = mod
    sys.modules["matplotlib"] = mpl
    from NanoRepeat import tk, nanoRepeat_joint
    return tk, nanoRepeat_joint


def random_cigar(rng, n_ops, max_len):
    ops, prev = [], "-"
    for _ in range(n_ops):
        op = rng.choice([o for o in "=XID" if o != prev and not (prev in "ID" and o in "ID")] or ["="])
        ln = rng.randint(1, max_len if op == "=" else max(1, max_len // 6))
        ops.append((op, ln)); prev = op
    if ops[0][0] in "ID":
        ops[0] = ("=", ops[0][1])
    if ops[-1][0] in "ID":
        ops[-1] = ("=", ops[-1][1])
    return ops


def main():
    tk, joint = import_reference()
    rng = random.Random(20260101)
    cases = []

    def add(ops, tstart, a, b, note=""):
        cigar = "".join(f"{n}{o}" for o, n in ops)
        tend = tstart + sum(n for o, n in ops if o in "=XD")
        r = tk.target_region_alignment_stats_from_cigar(cigar, tstart, tend, a, b)
        cases.append(dict(cigar=cigar, tstart=tstart, tend=tend, a=a, b=b, note=note,
                          expected=dict(num_match=r.num_match, num_mismatch=r.num_mismatch, num_ins=r.num_ins,
                                        num_del=r.num_del, score=r.score)))

    # hand-made: the probe of SURVEY.md section 4, runs straddling the window edges, insertions on the edges
    add([("=", 100)], 0, 10, 90, "plain matches")
    add([("=", 50), ("D", 20), ("=", 50)], 0, 60, 100, "deletion straddles the window start")
    add([("=", 50), ("D", 20), ("=", 50)], 0, 10, 60, "deletion straddles the window end")
    add([("=", 50), ("I", 7), ("=", 50)], 0, 50, 100, "insertion exactly at the window start (not counted: strict >)")
    add([("=", 50), ("I", 7), ("=", 50)], 0, 49, 100, "insertion one base inside")
    add([("=", 50), ("I", 7), ("=", 50)], 0, 0, 51, "insertion at end - 1 (not counted: strict < end - 1)")
    add([("=", 50), ("I", 7), ("=", 50)], 0, 0, 52, "insertion at end - 2")
    add([("=", 30), ("X", 3), ("=", 30), ("I", 2), ("=", 10), ("D", 1), ("=", 30)], 900, 990, 1060, "mixed, window inside")
    add([("=", 40)], 100, 50, 200, "alignment inside the window: both uncovered ends count as mismatches")
    add([("=", 40)], 100, 120, 130, "window inside one run")
    add([("X", 5), ("=", 5)], 10, 0, 12, "early break when the position passes the window end")
    for _ in range(400):
        ops = random_cigar(rng, rng.randint(1, 25), rng.choice([5, 30, 200]))
        tstart = rng.randint(0, 1200)
        span = sum(n for o, n in ops if o in "=XD")
        kind = rng.random()
        if kind < 0.6:
            a = tstart + rng.randint(-20, max(1, span // 2))
            b = a + rng.randint(1, max(2, span))
        elif kind < 0.8:
            a, b = tstart - rng.randint(0, 50), tstart + span + rng.randint(0, 50)
        else:
            a = rng.randint(0, 1500); b = a + rng.randint(1, 300)
        add(ops, tstart, max(a, 0), max(b, 1))
    with open(os.path.join(HERE, "joint_window_cases.json"), "w") as f:
        json.dump(dict(source="tk.target_region_alignment_stats_from_cigar (reference, unmodified)", cases=cases), f, indent=0)

    # selection: PAF files with tname "k1-k2", per read the grid point(s) with the best window score
    sel_cases = []
    R1, R2 = types.SimpleNamespace(repeat_unit_size=3), types.SimpleNamespace(repeat_unit_size=3)
    for c in range(12):
        left_len, mid_len = rng.choice([50, 200, 1000]), rng.choice([0, 12, 40])
        R1.repeat_unit_size, R2.repeat_unit_size = rng.randint(2, 6), rng.randint(2, 6)
        lines = []
        for r in range(rng.randint(1, 6)):
            qlen = rng.randint(200, 600)
            for _ in range(rng.randint(1, 12)):
                k1, k2 = rng.randint(0, 30), rng.randint(0, 15)
                tlen = left_len + R1.repeat_unit_size * k1 + mid_len + R2.repeat_unit_size * k2 + rng.choice([0, 5, 300])
                ops = random_cigar(rng, rng.randint(1, 15), rng.choice([10, 60]))
                span = sum(n for o, n in ops if o in "=XD")
                tstart = rng.randint(0, max(0, tlen - span)) if span <= tlen else 0
                tend = min(tstart + span, tlen) if span <= tlen else span
                cigar = "".join(f"{n}{o}" for o, n in ops)
                lines.append("\t".join(str(x) for x in (f"read{r}", qlen, 0, qlen, "+", f"{k1}-{k2}", max(tlen, tend), tstart,
                                                        tend, span, span, 60, f"AS:i:{rng.randint(50, 900)}", f"cg:Z:{cigar}", "tp:A:P")))
        rng.shuffle(lines)
        with tempfile.NamedTemporaryFile("w", suffix=".paf", delete=False) as f:
            f.write("\n".join(lines) + "\n")
            path = f.name
        est = joint.estimate_two_repeats_from_paf(path, left_len, mid_len, R1, R2)
        os.unlink(path)
        sel_cases.append(dict(left_len=left_len, mid_len=mid_len, m1=R1.repeat_unit_size, m2=R2.repeat_unit_size, paf_lines=lines,
                              expected={q: [float(est.repeat1_count_dict[q]), float(est.repeat2_count_dict[q])]
                                        for q in est.repeat1_count_dict}))
    with open(os.path.join(HERE, "joint_selection_cases.json"), "w") as f:
        json.dump(dict(source="nanoRepeat_joint.estimate_two_repeats_from_paf (reference, unmodified)", cases=sel_cases), f, indent=0)
    print(len(cases), "window cases,", len(sel_cases), "selection cases")


if __name__ == "__main__":
    main()
