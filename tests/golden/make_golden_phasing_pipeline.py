#!/usr/bin/env python
"""Golden vectors for Step 4 END TO END.

Runs only in the build container (needs /root/reference).  The REFERENCE's own, unmodified
nanoRepeat_bam.split_allele_using_gmm_1d (src/NanoRepeat/nanoRepeat_bam.py:515-575: 3-sigma trim, simulate_reads,
auto_GMM_1d on scikit-learn, create_allele_list_1d, remove_noisy_reads_1d, the output functions with no_details) is run
on seeded regions whose alleles lie well apart, with Python's and numpy's global generators seeded (the reference
draws from both, unseeded).  Recorded: the number of alleles, each allele's median size and read count, every read's
allele and confidence.  The CUDA path draws different random numbers by construction; on alleles this far apart it must
find the same alleles, put every read into the same one and agree on the medians (tests/test_gpu_gmm.py).

Usage: python tests/golden/make_golden_phasing_pipeline.py   -> tests/golden/phasing_pipeline_cases.json
"""
import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden_phasing                   # noqa: E402


def main():
    sa, nb = make_golden_phasing.import_reference()
    from NanoRepeat import repeat_region as rrmod
    rng = random.Random(20261020)
    cases = []
    for case, (centers, counts, ploidy, noisy) in enumerate([((17, 48), (25, 30), 2, False), ((33,), (40,), 2, False),
                                                             ((12, 45, 110), (20, 25, 22), 3, False), ((60, 85), (30, 30), 2, False),
                                                             ((20, 70), (40, 35), 2, True), ((150, 400), (18, 22), 2, False)]):
        sizes = {}
        for c, n in zip(centers, counts):
            for _ in range(n):
                sizes[f"read{len(sizes)}"] = round(c + rng.gauss(0, 0.02 * (10 + c)), 2)
        if noisy:
            for v in (118.0, 120.5, 121.0):                                    # a small third cluster: removed as noise (:502-514)
                sizes[f"read{len(sizes)}"] = v
        names = list(sizes)
        rng.shuffle(names)
        R = rrmod.RepeatRegion(no_details=True)
        R.chrom, R.start_pos, R.end_pos, R.repeat_unit_seq = "chrT", 100, 200, "CAG"
        for n in names:
            rd = rrmod.Read()
            rd.read_name = n
            rd.round3_repeat_size = sizes[n]
            R.read_dict[n] = rd
        random.seed(1000 + case)
        np.random.seed(2000 + case)
        with tempfile.TemporaryDirectory() as tmp:
            R.region_fq_file = os.path.join(tmp, "region.fastq")
            with open(R.region_fq_file, "w") as f:
                for n in names:
                    f.write(f"@{n}\nACGT\n+\n0000\n")
            R.out_prefix = os.path.join(tmp, "out")
            nb.split_allele_using_gmm_1d(R, ploidy, 0.07, 0.15, ploidy + 20, noisy)
        reads = {n: dict(allele_id=q.allele_id, confidence=q.phasing_confidence, size=float(q.repeat_size1))
                 for n, q in R.results.quantified_read_dict.items()}
        cases.append(dict(sizes={n: sizes[n] for n in names}, ploidy=ploidy, remove_noisy_reads=noisy, error_rate=0.07, max_mutual_overlap=0.15,
                          num_alleles=R.results.num_alleles,
                          alleles=[dict(median=int(a.repeat_size1), num_reads=int(a.num_supp_reads)) for a in R.results.quantified_allele_list],
                          reads=reads))
        print("case", case, centers, "->", R.results.num_alleles, [(a.repeat_size1, a.num_supp_reads) for a in R.results.quantified_allele_list])
    with open(os.path.join(HERE, "phasing_pipeline_cases.json"), "w") as f:
        json.dump(dict(source="nanoRepeat_bam.split_allele_using_gmm_1d (reference, unmodified; random.seed / numpy.random.seed set per case)",
                       cases=cases), f, indent=0)


if __name__ == "__main__":
    main()
