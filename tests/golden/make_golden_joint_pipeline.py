#!/usr/bin/env python
"""Golden vectors for the joint path's grid rounds END TO END.

Runs only in the build container (needs /root/reference).  The REFERENCE's own, unmodified
nanoRepeat_joint.fine_tune_read_count (src/NanoRepeat/nanoRepeat_joint.py:234-273: round2_estimation_of_repeat_size,
round3_estimation_of_repeat_size, choose_best_step_size, estimate_two_repeats_from_paf, the temp FASTQ / FASTA / PAF
files and all) is run on seeded HTT-like loci with only `pyminimap2.main` replaced: a PAF printer over the oracle's DP
(oracle/nr_oracle.c, nro_align_window_cigar: score, coordinates and the CIGAR of the canonical optimal alignment; both
strands, the better one printed; nothing below -s 80; output APPENDED to the -o file as the joint CLI relies on).  The
recorded per-read sizes and step sizes are what nanorepeat_b200.joint.quantify_two_repeats must return when its
alignments come from the same DP (tests/test_oracle_joint.py on CPU; on the GPU the CUDA path replaces the oracle and
tests/test_gpu_joint.py holds the two to each other).

Usage: python tests/golden/make_golden_joint_pipeline.py   -> tests/golden/joint_pipeline_cases.json
"""
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import nr_oracle                 # noqa: E402
import make_golden_joint                     # noqa: E402

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def main():
    tk, ref = make_golden_joint.import_reference()
    nr_oracle.build()
    rng = random.Random(20261019)

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
    for case in range(6):
        m1, m2 = ("CAG", "CCG") if case % 2 == 0 else (rs(rng.randint(2, 5)), rs(rng.randint(2, 4)))
        mid = "CAACAGCCGCCA" if case % 3 else rs(rng.choice([0, 6]))
        flank_l, flank_r = rng.choice([150, 260]), rng.choice([150, 240])
        kref1, kref2 = 19, 7
        chrom = rs(flank_l) + m1 * kref1 + mid + m2 * kref2 + rs(flank_r)
        r1 = ref.Repeat(); r2 = ref.Repeat()
        r1.chrom = r2.chrom = "chrT"
        r1.start, r1.end = flank_l, flank_l + len(m1) * kref1
        r2.start, r2.end = r1.end + len(mid), r1.end + len(mid) + len(m2) * kref2
        r1.repeat_unit, r2.repeat_unit = m1, m2
        r1.repeat_unit_size, r2.repeat_unit_size = len(m1), len(m2)
        r1.min_size = r2.min_size = 0
        r1.max_size, r2.max_size = 200, 50
        left, midseq, right = ref.extract_anchor_seq_for_two_repeats(chrom, r1, r2, 1000)
        alleles = [(rng.randint(8, 30), rng.randint(4, 12)), (rng.randint(35, 60), rng.randint(4, 12))]
        reads, init = {}, ref.Round1Estimation()
        for i in range(10):
            k1, k2 = alleles[i % 2]
            amp = left[-rng.randint(40, len(left)):] + m1 * k1 + midseq + m2 * k2 + right[:rng.randint(40, len(right))]
            seq = mut(amp, rng.choice([0.03, 0.08]))
            if i % 3 == 1:
                seq = "".join(COMP[c] for c in reversed(seq))
            name = f"read{i}"
            reads[name] = seq
            if i != 7:                                      # one read the initial estimate dropped
                init.repeat1_count_range_dict[name] = (max(0, k1 - rng.randint(6, 14)), k1 + rng.randint(6, 14))
                init.repeat2_count_range_dict[name] = (max(0, k2 - rng.randint(3, 6)), k2 + rng.randint(3, 7))
        if case == 4:                                       # narrow ranges: the coarse step is 1, no third round
            for name in init.repeat1_count_range_dict:
                a, b = init.repeat1_count_range_dict[name]; c = (a + b) // 2
                init.repeat1_count_range_dict[name] = (max(0, c - 1), c + 2)
                a, b = init.repeat2_count_range_dict[name]; c = (a + b) // 2
                init.repeat2_count_range_dict[name] = (max(0, c - 1), c + 2)

        calls = {"n": 0}

        def fake_main(cmd):
            tok = cmd.split()
            paf_path = tok[tok.index("-o") + 1]
            ref_fa, fq = tok[-2], tok[-1]
            with open(ref_fa) as f:
                tname = f.readline().strip()[1:]
                tseq = f.readline().strip()
            k1, k2 = (int(x) for x in tname.split("-"))
            a = max(len(left) - 10, 0)
            b = min(len(left) + len(m1) * k1 + len(midseq) + len(m2) * k2 + 10, len(tseq))
            lines = []
            with open(fq) as f:
                rec = f.read().strip().split("\n")
            for j in range(0, len(rec), 4):
                qn, qs = rec[j][1:].split()[0], rec[j + 1].strip()
                fwd = nr_oracle.align_window(qs, tseq, a, b, reverse=False, want_cigar=True)
                rev = nr_oracle.align_window(qs, tseq, a, b, reverse=True, want_cigar=True)
                best, strand = (rev, "-") if rev[:2] > fwd[:2] else (fwd, "+")
                s, _w, ts, te, cigar = best
                if s < 80:
                    continue
                lines.append("\t".join(str(x) for x in (qn, len(qs), 0, len(qs), strand, tname, len(tseq), ts, te, te - ts, te - ts, 60,
                                                        f"AS:i:{s}", "tp:A:P", f"cg:Z:{cigar}")))
            with open(paf_path, "a") as f:                   # appended: nanoRepeat_joint.py:393-395, :416
                f.write("".join(ln + "\n" for ln in lines))
            calls["n"] += 1
            return "", ""

        ref.pymm2.main = fake_main
        with tempfile.TemporaryDirectory() as tmp:
            fq = os.path.join(tmp, "reads.fastq")
            with open(fq, "w") as f:
                for n, s in reads.items():
                    f.write(f"@{n}\n{s}\n+\n{'0' * len(s)}\n")
            est = ref.fine_tune_read_count(init, fq, chrom, r1, r2, "ont", 1, tmp)
        names = list(reads)
        cases.append(dict(left=left, mid=midseq, right=right, motif1=m1, motif2=m2, max_size1=r1.max_size, max_size2=r2.max_size,
                          reads=[reads[n] for n in names],
                          range1=[init.repeat1_count_range_dict.get(n) for n in names],
                          range2=[init.repeat2_count_range_dict.get(n) for n in names],
                          size1=[None if n not in est.repeat1_count_dict else float(est.repeat1_count_dict[n]) for n in names],
                          size2=[None if n not in est.repeat2_count_dict else float(est.repeat2_count_dict[n]) for n in names],
                          step1=int(est.step_size1), step2=int(est.step_size2), alignment_calls=calls["n"], truth=[alleles[i % 2] for i in range(10)]))
        print("case", case, "calls", calls["n"], "steps", est.step_size1, est.step_size2,
              "sizes", [None if v is None else round(v, 1) for v in cases[-1]["size1"]][:5])
    with open(os.path.join(HERE, "joint_pipeline_cases.json"), "w") as f:
        json.dump(dict(source="nanoRepeat_joint.fine_tune_read_count (reference, unmodified; pyminimap2.main replaced by a PAF printer over "
                              "oracle/nr_oracle.c)", cases=cases), f, indent=0)


if __name__ == "__main__":
    main()
