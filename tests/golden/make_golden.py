#!/usr/bin/env python
"""Generate tests/golden/*.json by running the REFERENCE's own, unmodified hot-path functions.

Runs only in the build container (needs /root/reference).  The reference's decision code
(round1_and_round2_estimation, round3_estimation: src/NanoRepeat/nanoRepeat_bam.py:334-500, and the PAF
parser src/NanoRepeat/paf.py) is imported as-is; its absent third-party imports are stubbed, and
`pyminimap2.main` -- the alignment engine the reference shells out to, which is not in the reference tree --
is replaced by a function that parses the same command line, reads the same FASTA/FASTQ temp files and
prints PAF lines whose AS/tstart/tend come from the oracle DP (oracle/nr_oracle.c).  So the fixtures pin
everything the reference itself decides (task generation, int() truncations, span predicates, tie
averaging, fall-backs) and record the oracle DP's numbers for the engine part (PARITY UNPINNED there).

Usage: python tests/golden/make_golden.py [name ...]   (rewrites the named fixtures, default all, in place)
"""
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import nr_oracle  # noqa: E402
from nanorepeat_b200 import synth  # noqa: E402

SC = nr_oracle.scoring()


def _read_fasta(path):
    names, seqs = [], []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                names.append(line[1:].split()[0]); seqs.append("")
            else:
                seqs[-1] += line
    return names, seqs


def _read_fastq(path):
    names, seqs = [], []
    with open(path) as f:
        while True:
            l1 = f.readline(); l2 = f.readline(); l3 = f.readline(); l4 = f.readline()
            if not l4:
                break
            names.append(l1.strip()[1:].split()[0]); seqs.append(l2.strip())
    return names, seqs


def fake_pymm2_main(cmd):
    """Stand-in for pyminimap2.main(cmd) -> (stdout, stderr): last two tokens are <target.fa> <query.fx>."""
    toks = cmd.split()
    tfile, qfile = toks[-2], toks[-1]
    tnames, tseqs = _read_fasta(tfile)
    if qfile.endswith(".fastq") or qfile.endswith(".fq"):
        qnames, qseqs = _read_fastq(qfile)
    else:
        qnames, qseqs = _read_fasta(qfile)
    lines = []
    for qn, qs in zip(qnames, qseqs):
        res = nr_oracle.align_batch([qs] * len(tseqs), tseqs, SC, n_threads=nr_oracle.max_threads())
        order = sorted(range(len(tseqs)), key=lambda i: -int(res["score"][i]))   # minimap2 prints best first
        for i in order:
            s, ts, te = int(res["score"][i]), int(res["tstart"][i]), int(res["tend"][i])
            if s <= 0 or s < SC.min_dp_score:
                continue
            # qstart/qend/n_match/align_len/mapq are never read by rounds 2-3; fill with placeholders
            lines.append("\t".join(str(x) for x in (qn, len(qs), 0, len(qs), "+", tnames[i], len(tseqs[i]), ts, te,
                                                    te - ts, te - ts, 60, f"AS:i:{s}", "tp:A:P")))
    return "\n".join(lines) + ("\n" if lines else ""), ""


def import_reference():
    for name in ("pysam", "Levenshtein"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    # matplotlib is only used by the plotting code (out of scope); give it the attributes touched at import
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    mpl.use = lambda *a, **k: None
    mpl.rcParams = {}
    for sub, attrs in (("pyplot", ()), ("colors", ("Normalize",)), ("cm", ()), ("patches", ())):
        mod = types.ModuleType("matplotlib." + sub)
        for a in attrs:
            setattr(mod, a, object)
        setattr(mpl, sub, mod)
        sys.modules["matplotlib." + sub] = mod
    sys.modules["matplotlib"] = mpl
    mm = types.ModuleType("pyminimap2")
    mm.main = fake_pymm2_main
    sys.modules["pyminimap2"] = mm
    from NanoRepeat import nanoRepeat_bam, repeat_region
    nanoRepeat_bam.pymm2 = mm
    return nanoRepeat_bam, repeat_region


def run_reference(nrb, rr, reg, fast_mode, tmp):
    """Drive the reference's round 1/2/3 functions on one synthetic region."""
    R = rr.RepeatRegion()
    R.left_anchor_seq = reg.left_anchor_seq
    R.right_anchor_seq = reg.right_anchor_seq
    R.left_anchor_len = len(reg.left_anchor_seq)
    R.right_anchor_len = len(reg.right_anchor_seq)
    R.repeat_unit_seq = reg.repeat_unit_seq
    R.temp_out_dir = tmp
    R.core_seq_fq_file = os.path.join(tmp, "core_sequences.fastq")
    with open(R.core_seq_fq_file, "w") as f:
        for name, core, dist in zip(reg.read_names, reg.core_seqs, reg.dist_between_anchors):
            rd = rr.Read()
            rd.read_name = name
            rd.dist_between_anchors = dist
            R.read_dict[name] = rd
            R.read_core_seq_dict[name] = core
            f.write(f"@{name}\n{core}\n+\n{'0' * len(core)}\n")
    nrb.round1_and_round2_estimation(reg.data_type, R, 1)
    nrb.round3_estimation(reg.data_type, fast_mode, R, 1)
    out = []
    for name in reg.read_names:
        rd = R.read_dict[name]
        out.append(dict(r1=rd.round1_repeat_size, r2=rd.round2_repeat_size,
                        r3=None if rd.round3_repeat_size is None else float(rd.round3_repeat_size)))
    return out


def region_to_json(reg):
    return dict(name=reg.name, data_type=reg.data_type, left=reg.left_anchor_seq, right=reg.right_anchor_seq,
                motif=reg.repeat_unit_seq, read_names=reg.read_names, cores=reg.core_seqs,
                dists=reg.dist_between_anchors, true_sizes=reg.true_sizes)


def crafted_regions():
    """Hand-made edge cases (SURVEY.md section 8c 'golden vectors to author')."""
    import numpy as np
    rng = np.random.default_rng(77)
    regs = []
    L, R = synth.random_seq(rng, 300), synth.random_seq(rng, 300)
    # (i) perfect reads; half-unit read -> tie between k and k+1; short flanks (chromosome end)
    reg = synth.SynthRegion("crafted_perfect", L, R, "CAG", "hifi")
    for i, (k, extra) in enumerate([(10, ""), (20, ""), (7, "CA"), (0, ""), (1, ""), (33, "C")]):
        core = L[-100:] + "CAG" * k + extra + R[:100]
        reg.read_names.append(f"p{i}"); reg.core_seqs.append(core)
        reg.dist_between_anchors.append(3 * k + len(extra)); reg.true_sizes.append(k)
    regs.append(reg)
    # (ii) reads missing a flank / garbage reads -> r2 None or span predicates fail; slightly negative dist
    reg = synth.SynthRegion("crafted_edge", L, R, "TATTG", "ont")
    cases = [
        (L[-100:] + "TATTG" * 12, 60),                        # no right flank: round 3 cannot span -> r3 = r2
        ("TATTG" * 12 + R[:100], 60),                         # no left flank: round 2 filter fails -> None
        (synth.random_seq(rng, 250), 50),                     # garbage: below min_dp_score everywhere
        (L[-100:] + R[:100], -3),                             # zero units, negative dist (> -10 allowed, :213)
        (L[-100:] + "TATTG" * 40 + R[:100], 200),
        (L[-30:] + "TATTG" * 9 + R[:25], 45),                 # short buffers
    ]
    for i, (core, dist) in enumerate(cases):
        reg.read_names.append(f"e{i}"); reg.core_seqs.append(core)
        reg.dist_between_anchors.append(dist); reg.true_sizes.append(-1)
    regs.append(reg)
    # (iii) long-gap piece: reads with 30-40 bp deletions inside the repeat relative to their neighbours
    reg = synth.SynthRegion("crafted_longgap", L, R, "GGGGCC", "ont")
    for i, k in enumerate([30, 45, 60]):
        core = L[-100:] + "GGGGCC" * k + R[:100]
        reg.read_names.append(f"g{i}"); reg.core_seqs.append(core)
        reg.dist_between_anchors.append(6 * k + (40 if i == 1 else 0)); reg.true_sizes.append(k)
    regs.append(reg)
    return regs


def main():
    nrb, rr = import_reference()
    fixtures = {}
    # small slices of the BASELINE configs + crafted cases
    fixtures["cfg1_small"] = (synth.config1(seed=1, n_regions=3, reads_per_region=6), False)
    fixtures["cfg2_small"] = (synth.config2(seed=2, n_reads=8), False)
    fixtures["cfg3_small"] = (synth.config3(seed=3, n_loci=4, reads_per_locus=4), False)
    fixtures["cfg5_small_fast"] = (synth.config5(seed=5, n_reads=12, reads_per_region=3, k_max=120), True)
    fixtures["crafted"] = (crafted_regions(), False)
    # the long shapes: config 4's expansions (C9orf72 ~6.4 kb core x 101 rungs, FMR1 ~1.7 kb x 51) and config 5 with the
    # full ladder rule (cores up to 10.6 kb, up to 167 rungs) -- a minute of oracle time each
    fixtures["cfg4_small"] = (synth.config4(seed=4, reads_per_locus=5), False)
    fixtures["cfg5_small"] = (synth.config5(seed=8, n_reads=8, reads_per_region=2, k_max=2000), False)
    only = set(sys.argv[1:])
    for name, (regs, fast_mode) in fixtures.items():
        if only and name not in only:
            continue
        doc = dict(fast_mode=fast_mode, scoring={n: getattr(SC, n) for n, _ in nr_oracle.Scoring._fields_},
                   generated_by="tests/golden/make_golden.py (reference functions from /root/reference, DP from oracle)",
                   regions=[])
        for reg in regs:
            with tempfile.TemporaryDirectory() as tmp:
                res = run_reference(nrb, rr, reg, fast_mode, tmp)
            d = region_to_json(reg)
            d["expected"] = res
            doc["regions"].append(d)
        path = os.path.join(HERE, name + ".json")
        with open(path, "w") as f:
            json.dump(doc, f, indent=0, separators=(",", ":"))
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
