#!/usr/bin/env python
"""Golden vectors for Step 1 (anchor finding + core extraction).

Runs only in the build container (needs /root/reference).  The REFERENCE's own, unmodified
find_anchor_locations_in_reads and make_core_seq_fastq (src/NanoRepeat/nanoRepeat_bam.py:165-331, with paf.py's strand
flip) are run on seeded raw reads with only `pyminimap2.main` replaced: a PAF printer over the oracle's DP that prints,
per read and anchor, the hits the repo's stated rule defines (nanorepeat_b200/anchoring.py: the best exact local
alignment on each strand that reaches -s 80, best first, mapq 60), in standard PAF coordinates (query = the read, on
its original strand; target = the anchor).  What the reference then decides -- which reads it keeps, strand,
dist_between_anchors, the core / middle positions with their 100-base buffers and clamps, the sliced sequences -- is
recorded; oracle/anchoring.step1 (the checker of the GPU path) must reproduce it (tests/test_oracle_anchoring.py).

Usage: python tests/golden/make_golden_anchoring.py   -> tests/golden/anchoring_cases.json
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import nr_oracle                  # noqa: E402
from nanorepeat_b200 import synth             # noqa: E402
import make_golden                            # noqa: E402

COMP = str.maketrans("ACGT", "TGCA")


def fake_main(cmd):
    toks = cmd.split()
    tnames, tseqs = make_golden._read_fasta(toks[-2])           # anchors.fasta: left_anchor, right_anchor
    qnames, qseqs = make_golden._read_fastq(toks[-1])           # the region's raw reads
    lines = []
    for qn, qs in zip(qnames, qseqs):
        rc = qs.translate(COMP)[::-1]
        for tn, tseq in zip(tnames, tseqs):
            # the anchor is the DP's query, the read (or its reverse complement) its target: (tstart, tend) are read coordinates
            res = nr_oracle.align_batch([tseq, tseq], [qs, rc], n_threads=2)
            hits = []
            for k, strand in enumerate("+-"):
                s, ts, te = int(res["score"][k]), int(res["tstart"][k]), int(res["tend"][k])
                if s <= 0 or s < 80:
                    continue
                qstart, qend = (ts, te) if strand == "+" else (len(qs) - te, len(qs) - ts)      # PAF: original strand
                hits.append((s, "\t".join(str(x) for x in (qn, len(qs), qstart, qend, strand, tn, len(tseq), 0, len(tseq), te - ts, te - ts,
                                                          60, f"AS:i:{s}", "tp:A:P"))))
            hits.sort(key=lambda h: -h[0])
            lines += [h[1] for h in hits]
    return "\n".join(lines) + ("\n" if lines else ""), ""


def main():
    nrb, rr = make_golden.import_reference()
    nrb.pymm2.main = fake_main
    nr_oracle.build()
    cases = []
    for seed, motif, alleles, flank in ((31, "CAG", (17, 55), 1000), (32, "GGGGCC", (8, 30), 400), (33, "AT", (20, 21), 120),
                                        (34, "CTG", (5, 140), 1000)):
        reg, names, seqs, truth = synth.region_reads(seed=seed, n_reads=16, motif=motif, alleles=alleles, flank=flank, outer=600)
        if seed == 33:
            seqs[2] = reg.left_anchor_seq[-60:]                                      # a read with a left hit only
            seqs[5] = seqs[5][: len(seqs[5]) // 3]                                   # truncated: anchors missing
        R = rr.RepeatRegion()
        R.left_anchor_seq, R.right_anchor_seq, R.repeat_unit_seq = reg.left_anchor_seq, reg.right_anchor_seq, motif
        with tempfile.TemporaryDirectory() as tmp:
            R.temp_out_dir = tmp
            R.region_fq_file = os.path.join(tmp, "region.fastq")
            with open(R.region_fq_file, "w") as f:
                for n, s in zip(names, seqs):
                    f.write(f"@{n}\n{s}\n+\n{'0' * len(s)}\n")
            nrb.find_anchor_locations_in_reads("ont", R, 1)
            nrb.make_core_seq_fastq(R)
            _n, mids = make_golden._read_fastq(R.mid_seq_fq_file)
            mid_of = dict(zip(_n, mids))
        kept = {}
        for n, rd in R.read_dict.items():
            kept[n] = dict(strand=rd.strand, dist=rd.dist_between_anchors, core_start=rd.core_seq_start_pos, core_end=rd.core_seq_end_pos,
                           mid_start=rd.mid_seq_start_pos, mid_end=rd.mid_seq_end_pos, left_buffer=rd.left_buffer_len,
                           right_buffer=rd.right_buffer_len, core=R.read_core_seq_dict[n], mid=mid_of[n])
        cases.append(dict(left=reg.left_anchor_seq, right=reg.right_anchor_seq, motif=motif, names=names, reads=seqs, kept=kept))
        print(motif, "reads", len(names), "kept", len(kept), "strands", "".join(v["strand"] for v in kept.values()))
    with open(os.path.join(HERE, "anchoring_cases.json"), "w") as f:
        json.dump(dict(source="nanoRepeat_bam.find_anchor_locations_in_reads + make_core_seq_fastq (reference, unmodified; pyminimap2.main "
                              "replaced by a PAF printer over oracle/nr_oracle.c with the hit rule of nanorepeat_b200/anchoring.py)",
                       cases=cases), f, indent=0)


if __name__ == "__main__":
    main()
