#!/usr/bin/env python
"""Golden vectors for the 1-D phasing's host bookkeeping.

Runs only in the build container (needs /root/reference).  Calls the REFERENCE's own, unmodified
split_alleles.remove_outlier_reads_1d (src/NanoRepeat/split_alleles.py:141-154), split_alleles.create_allele_list_1d
(:258-293), nanoRepeat_bam.remove_noisy_reads_1d (nanoRepeat_bam.py:502-514) and split_alleles.interval_has_overlap
(:90-96) on seeded inputs; the fitted mixture handed to create_allele_list_1d is a stand-in object carrying the recorded
means / variances / weights whose predict / predict_proba are scikit-learn's own (a GaussianMixture with its parameters
set).  Third-party imports the reference makes at module level and never uses here are stubbed.

Usage: python tests/golden/make_golden_phasing.py      -> tests/golden/phasing_cases.json
"""
import json
import os
import random
import sys
import types

import numpy as np

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
    from NanoRepeat import split_alleles, nanoRepeat_bam
    return split_alleles, nanoRepeat_bam


def mixture_with(weights, means, variances):
    from sklearn.mixture import GaussianMixture
    g = GaussianMixture(n_components=len(means), covariance_type="diag")
    g.weights_ = np.array(weights, dtype=float)
    g.means_ = np.array(means, dtype=float).reshape(-1, 1)
    g.covariances_ = np.array(variances, dtype=float).reshape(-1, 1)
    g.precisions_cholesky_ = 1.0 / np.sqrt(g.covariances_)
    return g


def main():
    sa, nb = import_reference()
    rng = random.Random(20260318)
    cases = []
    for case in range(40):
        k = rng.choice([1, 2, 2, 3, 4])
        centers = sorted(rng.sample(range(6, 200, 9), k))
        counts = [rng.choice([1, 3, 8, 20, 40]) for _ in range(k)]
        sizes = {}
        for c, n in zip(centers, counts):
            for _ in range(n):
                sizes[f"r{len(sizes)}"] = round(c + rng.gauss(0, 0.03 * (10 + c)), 2)
        if case % 4 == 0:
            sizes[f"r{len(sizes)}"] = 900.0                                   # an outlier for the 3-sd trim
        names = list(sizes)
        rng.shuffle(names)
        sizes = {n: sizes[n] for n in names}
        if len(sizes) < 2:
            continue
        kept_names, kept_sizes = sa.remove_outlier_reads_1d(sizes)
        # a mixture near the truth (what a fit would return), sometimes with an empty extra component
        means = [c + rng.uniform(-0.3, 0.3) for c in centers]
        variances = [max(0.05, (0.03 * (10 + c)) ** 2 * rng.uniform(0.5, 2.0)) for c in centers]
        weights = [n / sum(counts) for n in counts]
        if case % 5 == 1:
            means.append(500.0); variances.append(1.0); weights = [w * 0.99 for w in weights] + [0.01]
        g = mixture_with(weights, means, variances)
        arr = np.array(kept_sizes).reshape(-1, 1)
        alleles = sa.create_allele_list_1d(len(means), g, kept_names, arr, sizes, 0.95)
        before = [dict(mean=float(a.gmm_mean1), sd=float(a.gmm_sd1), reads=list(a.readname_list), sizes=[float(s) for s in a.repeat1_size_list],
                       proba=[float(p) for p in a.probability_list], num_reads=a.num_reads, median=a.repeat1_median_size,
                       gmm_min=float(a.gmm_min1), gmm_max=float(a.gmm_max1), confidence=list(a.confidence_list)) for a in alleles]
        ploidy = rng.choice([1, 2, 2, 3])
        alleles2, removed = nb.remove_noisy_reads_1d(list(alleles), ploidy)
        cases.append(dict(sizes=sizes, kept=kept_names, weights=weights, means=means, variances=variances,
                          label=[int(l) for l in g.predict(arr)], proba=[float(p) for p in g.predict_proba(arr).max(axis=1)],
                          alleles=before, ploidy=ploidy, after_noise_removal=[list(a.readname_list) for a in alleles2], removed=removed))
    overlaps = []
    for _ in range(200):
        a = sorted(rng.uniform(0, 50) for _ in range(2)); b = sorted(rng.uniform(0, 50) for _ in range(2))
        if rng.random() < 0.2:
            b[0] = a[1]
        overlaps.append(dict(a=a, b=b, expected=bool(sa.interval_has_overlap(tuple(a), tuple(b)))))
    with open(os.path.join(HERE, "phasing_cases.json"), "w") as f:
        json.dump(dict(source="split_alleles.remove_outlier_reads_1d / create_allele_list_1d / interval_has_overlap, "
                              "nanoRepeat_bam.remove_noisy_reads_1d (reference, unmodified)", cases=cases, overlaps=overlaps), f, indent=0)
    print(len(cases), "cases,", len(overlaps), "interval pairs")


if __name__ == "__main__":
    main()
