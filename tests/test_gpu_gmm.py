"""GPU parity of the batched 1-D phasing (csrc/nr_gmm.cu through the C ABI) with the CPU checker
that shares its random-number generator (oracle/gmm.py, itself pinned to scikit-learn in tests/test_oracle_gmm.py), and
statistically with scikit-learn run the way the reference runs it.  Floating point: tolerances are written at each check."""
import math
import random
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixture(rng, centers, n_each, err=0.02):
    x = []
    for c, n in zip(centers, n_each):
        x += [round(c + rng.gauss(0, err * (10 + c)), 2) for _ in range(n)]
    rng.shuffle(x)
    return x


def _regions(seed, n):
    rng = random.Random(seed)
    out = []
    for g in range(n):
        k = rng.choice([1, 2, 2, 2, 3])
        centers = sorted(rng.sample(range(8, 160, 12), k))
        out.append(_mixture(rng, centers, [rng.randint(12, 40) for _ in range(k)]))
    return out


def test_bootstrap_equals_checker(engine):
    from oracle import gmm
    p = engine.GmmParams(error_rate=0.07, seed=99)
    regions = [[17.0, 18.5, 44.0], [120.25], [3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0]]
    got = engine.gmm_bootstrap(p, regions, region_id_base=5)
    for g, xs in enumerate(regions):
        exp = gmm.bootstrap(xs, 0.07, 99, 5 + g)
        assert got[g].shape == exp.shape
        assert np.allclose(got[g], exp, rtol=0, atol=1e-11), g               # log / cos differ in the last ulp at most
    assert abs(np.concatenate(got[:1]).reshape(100, 3)[:, 0].std() - 0.07 * 27) < 0.5


def test_fit_equals_checker_from_the_same_starts(engine):
    """Best of 10 starts, same starts (hashed), same EM: parameters to 1e-7 relative (summation order is the only difference)."""
    from oracle import gmm
    rng = random.Random(4)
    p = engine.GmmParams(seed=21, max_components=6)
    data, ncs = [], []
    for case in range(10):
        k = rng.choice([1, 2, 3, 4])
        xs = _mixture(rng, sorted(rng.sample(range(10, 150, 10), k)), [rng.randint(100, 600) for _ in range(k)], err=0.04)
        data.append(xs); ncs.append(rng.choice([k, k, max(1, k - 1), k + 1]))
    got = engine.gmm1d_fit(p, data, ncs, region_ids=list(range(100, 110)))
    for i, xs in enumerate(data):
        w, m, v, lower, it, conv = gmm.best_fit(np.array(xs), ncs[i], 21, 100 + i)
        assert abs(got["lower"][i] - lower) < 1e-9 and got["iters"][i] == it, (i, got["lower"][i], lower, got["iters"][i], it)
        assert np.allclose(got["weights"][i], w, rtol=1e-7, atol=1e-12), i
        assert np.allclose(got["means"][i], m, rtol=1e-7), i
        assert np.allclose(got["variances"][i], v, rtol=1e-6), i


def test_phase_equals_checker_and_is_independent_of_batching(engine):
    from oracle import gmm
    regions = _regions(8, 14) + [[30.0], [], [20.0, 20.0], [25.0] * 20 + [400.0]]
    p = engine.GmmParams(error_rate=0.07, max_mutual_overlap=0.15, max_components=5, seed=3)
    got = engine.phase_1d(p, regions)
    for g, xs in enumerate(regions):
        exp = gmm.phase_1d(xs, 0.07, 5, 0.15, seed=3, region=g)
        assert got[g]["n"] == exp["n"], (g, got[g]["n"], exp["n"])
        lab = [int(l) for l in got[g]["label"]]
        assert [i for i, l in enumerate(lab) if l >= 0] == (exp["kept"] if exp["n"] else []), g
        if exp["n"]:
            assert np.allclose(got[g]["means"], exp["means"], rtol=1e-6), g
            assert np.allclose(got[g]["variances"], exp["variances"], rtol=1e-5), g
            assert [lab[i] for i in exp["kept"]] == [int(l) for l in exp["label"]], g
            assert np.allclose([got[g]["proba"][i] for i in exp["kept"]], exp["proba"], rtol=1e-6, atol=1e-9), g
    assert got[-1]["label"][-1] == -1 and got[-1]["n"] == 1                     # the 400 is trimmed
    assert got[-3]["n"] == 0 and got[-4]["n"] == 0
    # the same regions alone, with their ids: identical numbers
    for g in (0, 5, 9):
        alone = engine.phase_1d(p, [regions[g]], region_id_base=g)[0]
        assert alone["n"] == got[g]["n"] and np.array_equal(alone["means"], got[g]["means"]) and np.array_equal(alone["label"], got[g]["label"])


def test_phase_agrees_with_the_reference_recipe_on_sklearn(engine):
    """split_alleles.auto_GMM_1d's recipe with scikit-learn itself (own bootstrap draws, own k-means starts): on alleles that
    are apart the number of alleles and every read's allele agree; means within 3 standard errors of the bootstrap."""
    from sklearn.mixture import GaussianMixture
    from scipy.stats import norm
    from nanorepeat_b200 import phasing
    rng = random.Random(12)
    regions = [_mixture(rng, c, n) for c, n in [((17, 48), (25, 30)), ((33,), (40,)), ((12, 45, 110), (20, 25, 22)), ((60, 75), (30, 30))]]
    dicts = [{f"read{i}": x for i, x in enumerate(xs)} for xs in regions]
    res = phasing.phase_regions_1d(dicts, ploidy=2, error_rate=0.07, max_mutual_overlap=0.15, max_num_components=6, seed=1)
    for g, xs in enumerate(regions):
        mean, sd = np.mean(xs), np.std(xs)
        kept = [x for x in xs if not (x < max(mean - 3 * sd, 0) or x > mean + 3 * sd)]
        sim = [x + rng.gauss(0, 0.07 * (10 + x)) for x in kept * 100]
        X = np.array(sim).reshape(-1, 1)
        n_best, gm = 6, None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for n in range(2, 7):
                gm = GaussianMixture(n_components=n, covariance_type="diag", n_init=10, random_state=n).fit(X)
                clash = False
                for i in range(n):
                    for j in range(i + 1, n):
                        si, sj = max(1.0, math.sqrt(gm.covariances_[i][0])), max(1.0, math.sqrt(gm.covariances_[j][0]))
                        a = (norm.isf(0.85, gm.means_[i][0], si), norm.isf(0.15, gm.means_[i][0], si))
                        b = (norm.isf(0.85, gm.means_[j][0], sj), norm.isf(0.15, gm.means_[j][0], sj))
                        clash = clash or max(a[0], b[0]) - min(a[1], b[1]) <= 0
                if clash:
                    n_best = n - 1
                    break
            gm = GaussianMixture(n_components=n_best, covariance_type="diag", n_init=10, random_state=7).fit(X)
        alleles, removed = res[g]
        assert len(alleles) == n_best, (g, len(alleles), n_best)
        sk_means = sorted(gm.means_[:, 0])
        for a, mref in zip(alleles, sk_means):                                   # alleles come sorted by gmm_mean1
            se = a.gmm_sd1 / math.sqrt(100 * a.num_reads)
            assert abs(a.gmm_mean1 - mref) < 6 * se + 0.02, (g, a.gmm_mean1, mref, se)
        rank = {int(c): r for r, c in enumerate(np.argsort(gm.means_[:, 0]))}
        sk_lab = {f"read{i}": rank[int(l)] for i, l in zip([i for i, x in enumerate(xs) if x in kept], gm.predict(np.array(kept).reshape(-1, 1)))}
        ours = {name: r for r, a in enumerate(alleles) for name in a.readname_list}
        differ = [k for k in ours if ours[k] != sk_lab[k]]
        assert len(differ) <= 1, (g, differ)
        assert all(c in ("HIGH", "LOW") for a in alleles for c in a.confidence_list)


def test_steps_1_to_4_in_memory_find_the_simulated_alleles(engine):
    """pipeline.quantify_regions(phase=True): raw reads of several regions -> anchors, cores, rounds 1-3, phased alleles,
    all batched; the alleles' medians land on the simulated repeat counts."""
    import nanorepeat_b200 as nrb
    from nanorepeat_b200 import synth, pipeline
    regions, reads, alleles = [], [], [(17, 55), (30, 30), (12, 40)]
    for g, al in enumerate(alleles):
        reg, names, seqs, truth = synth.region_reads(seed=20 + g, n_reads=50, motif="CAG", alleles=al)
        rr = nrb.RepeatRegion()
        rr.left_anchor_seq, rr.right_anchor_seq, rr.repeat_unit_seq = reg.left_anchor_seq, reg.right_anchor_seq, "CAG"
        regions.append(rr); reads.append(dict(zip(names, seqs)))
    pipeline.quantify_regions(regions, reads, "ont", phase=True, seed=4)
    for rr, al in zip(regions, alleles):
        assert rr.allele_list is not None
        medians = sorted(a.repeat1_median_size for a in rr.allele_list if a.num_reads >= 5)
        want = sorted(set(al))
        assert len(medians) == len(want) and all(abs(m - w) <= 2 for m, w in zip(medians, want)), (al, medians)
        assert sum(a.num_reads for a in rr.allele_list) <= len(rr.read_dict)


def test_many_samples_take_the_cluster_path_and_equal_the_checker(engine):
    """Regions of 16 384 bootstrapped samples and more are fitted by a cluster of thread blocks (partial sums through
    distributed shared memory): same starts, same EM as the checker; the summation order is all that differs."""
    from oracle import gmm
    rng = random.Random(9)
    p = engine.GmmParams(seed=77, max_components=4)
    data = [_mixture(rng, (20, 60), (9000, 11000), err=0.05), _mixture(rng, (15, 40, 90), (7000, 6000, 8000), err=0.04),
            _mixture(rng, (33,), (17000,), err=0.03), _mixture(rng, (20, 60), (300, 200), err=0.05)]
    ncs = [2, 3, 2, 2]
    got = engine.gmm1d_fit(p, data, ncs, region_ids=[5, 6, 7, 8])
    for i, xs in enumerate(data):
        w, m, v, lower, it, conv = gmm.best_fit(np.array(xs), ncs[i], 77, 5 + i)
        assert abs(got["lower"][i] - lower) < 1e-8 and got["iters"][i] == it, (i, got["lower"][i], lower, got["iters"][i], it)
        assert np.allclose(got["weights"][i], w, rtol=1e-6, atol=1e-10) and np.allclose(got["means"][i], m, rtol=1e-7), i
        assert np.allclose(got["variances"][i], v, rtol=1e-5), i
    # a region of 400 reads (40 000 samples) through the whole phasing, beside a small one
    sizes = _mixture(rng, (17, 55), (210, 190), err=0.02)
    pp = engine.GmmParams(error_rate=0.07, max_mutual_overlap=0.15, max_components=4, seed=2)
    got = engine.phase_1d(pp, [sizes, sizes[:30]])
    exp = gmm.phase_1d(sizes, 0.07, 4, 0.15, seed=2, region=0)
    assert got[0]["n"] == exp["n"] == 2 and np.allclose(got[0]["means"], exp["means"], rtol=1e-6)
    assert [int(l) for l in got[0]["label"] if l >= 0] == [int(l) for l in exp["label"]]
    assert got[1]["n"] >= 1


def test_phase_equals_the_reference_step4_on_separated_alleles(engine):
    """The reference's own, unmodified split_allele_using_gmm_1d (scikit-learn, its own random draws, seeded when the
    fixture was made: tests/golden/make_golden_phasing_pipeline.py) against phasing.phase_regions_1d on the same sizes:
    same number of alleles, same median size and read count per allele, every read in the same allele."""
    import json
    import os
    from nanorepeat_b200 import phasing
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phasing_pipeline_cases.json")) as f:
        doc = json.load(f)
    assert len(doc["cases"]) >= 6
    dicts = [c["sizes"] for c in doc["cases"]]
    for g, c in enumerate(doc["cases"]):
        alleles, removed = phasing.phase_regions_1d([dicts[g]], ploidy=c["ploidy"], error_rate=c["error_rate"],
                                                    max_mutual_overlap=c["max_mutual_overlap"], max_num_components=c["ploidy"] + 20,
                                                    remove_noisy_reads=c["remove_noisy_reads"], seed=g, region_id_base=g)[0]
        assert len(alleles) == c["num_alleles"], (g, len(alleles), c["num_alleles"])
        assert [(a.repeat1_median_size, a.num_reads) for a in alleles] == [(e["median"], e["num_reads"]) for e in c["alleles"]], g
        ours = {name: i + 1 for i, a in enumerate(alleles) for name in a.readname_list}
        assert ours == {n: r["allele_id"] for n, r in c["reads"].items()}, g
        conf = {name: cf for a in alleles for name, cf in zip(a.readname_list, a.confidence_list)}
        differ = [n for n, r in c["reads"].items() if conf[n] != r["confidence"]]
        assert len(differ) <= max(1, len(conf) // 20), (g, differ)          # HIGH / LOW sits on a fitted 2-sd bound: a boundary read may flip
