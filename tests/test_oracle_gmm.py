"""oracle/gmm.py against scikit-learn and scipy themselves -- the libraries the reference calls
(split_alleles.py:171-200): EM from identical starting parameters, the overlap rule, the whole auto-GMM on well
separated alleles."""
import math
import random
import warnings

import numpy as np
import pytest

from oracle import gmm


def _mixture(rng, centers, n_each, err=0.03):
    x = []
    for c, n in zip(centers, n_each):
        x += [c + rng.gauss(0, err * (10 + c)) for _ in range(n)]
    return np.array(x)


def test_em_equals_sklearn_from_the_same_start():
    from sklearn.mixture import GaussianMixture
    rng = random.Random(1)
    for case in range(12):
        n = rng.choice([2, 3, 4])
        x = _mixture(rng, [rng.uniform(5, 200) for _ in range(n)], [rng.randint(30, 400) for _ in range(n)])
        w0 = np.full(n, 1.0 / n)
        m0 = np.array(sorted(rng.uniform(x.min(), x.max()) for _ in range(n)))
        v0 = np.full(n, float(np.var(x)) / n)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sk = GaussianMixture(n_components=n, covariance_type="diag", n_init=1, weights_init=w0, means_init=m0.reshape(-1, 1),
                                 precisions_init=(1.0 / v0).reshape(-1, 1)).fit(x.reshape(-1, 1))
        w, m, v, lower, it, conv = gmm.em_fit(x, w0, m0, v0)
        assert it == sk.n_iter_ and conv == sk.converged_, (case, it, sk.n_iter_)
        assert np.allclose(w, sk.weights_, rtol=1e-9, atol=1e-12) and np.allclose(m, sk.means_[:, 0], rtol=1e-9)
        assert np.allclose(v, sk.covariances_[:, 0], rtol=1e-8) and abs(lower - sk.lower_bound_) < 1e-9
        lab, pr = gmm.labels(x, w, m, v)
        assert np.array_equal(lab, sk.predict(x.reshape(-1, 1)))
        assert np.allclose(pr, sk.predict_proba(x.reshape(-1, 1)).max(axis=1), rtol=1e-9)


def test_isf_and_overlap_rule_equal_scipy():
    from scipy.stats import norm
    for o in (0.1, 0.05, 0.2, 0.01, 0.4999):
        assert abs(gmm.std_isf(o) - norm.isf(o)) < 1e-12
    rng = random.Random(2)
    for _ in range(300):
        n = rng.randint(2, 4)
        means = [rng.uniform(0, 60) for _ in range(n)]
        var = [rng.uniform(0.01, 30) for _ in range(n)]
        o = rng.choice([0.1, 0.05, 0.2])
        exp = False
        for i in range(n):
            for j in range(i + 1, n):
                si, sj = max(1.0, math.sqrt(var[i])), max(1.0, math.sqrt(var[j]))
                a = (norm.isf(1 - o, means[i], si), norm.isf(o, means[i], si))
                b = (norm.isf(1 - o, means[j], sj), norm.isf(o, means[j], sj))
                exp = exp or (max(a[0], b[0]) - min(a[1], b[1]) <= 0)
        assert gmm.overlap(means, var, o) == exp


def test_trim_is_three_sigma_with_a_floor_at_zero():
    sizes = [20.0] * 30 + [21.0] * 30 + [400.0]
    kept = gmm.trim(sizes)
    assert 60 not in kept and len(kept) == 60
    lo, hi = gmm.outlier_cutoffs([1.0, 2.0, 90.0])
    assert lo == 0.0 and hi > 90


@pytest.mark.parametrize("alleles,counts", [((17, 48), (25, 25)), ((30,), (40,)), ((12, 40, 90), (20, 30, 25))])
def test_auto_gmm_finds_the_alleles_like_the_reference_pipeline(alleles, counts):
    """Same steps with sklearn in the middle (what split_alleles.auto_GMM_1d runs) on the same bootstrap: the number of
    components and the means agree (SURVEY.md section 4's probe: 2 components at 16.99 / 48.01)."""
    from sklearn.mixture import GaussianMixture
    from scipy.stats import norm
    rng = random.Random(7)
    sizes = list(_mixture(rng, alleles, counts, err=0.02))
    res = gmm.phase_1d(sizes, 0.03, 6, 0.1, seed=11)
    xs = [sizes[i] for i in res["kept"]]
    sim = gmm.bootstrap(xs, 0.03, 11, 0).reshape(-1, 1)
    best = None
    for n in range(2, 7):
        g = GaussianMixture(n_components=n, covariance_type="diag", n_init=10, random_state=0).fit(sim)
        clash = False
        for i in range(n):
            for j in range(i + 1, n):
                si, sj = max(1.0, math.sqrt(g.covariances_[i][0])), max(1.0, math.sqrt(g.covariances_[j][0]))
                a = (norm.isf(0.9, g.means_[i][0], si), norm.isf(0.1, g.means_[i][0], si))
                b = (norm.isf(0.9, g.means_[j][0], sj), norm.isf(0.1, g.means_[j][0], sj))
                clash = clash or max(a[0], b[0]) - min(a[1], b[1]) <= 0
        if clash:
            best = n - 1
            break
    assert res["n"] == best == len(alleles)
    g = GaussianMixture(n_components=best, covariance_type="diag", n_init=10, random_state=0).fit(sim)
    assert np.allclose(sorted(res["means"]), sorted(g.means_[:, 0]), atol=0.05)
    order_ours, order_sk = np.argsort(res["means"]), np.argsort(g.means_[:, 0])
    rank_ours = {int(c): r for r, c in enumerate(order_ours)}
    rank_sk = {int(c): r for r, c in enumerate(order_sk)}
    sk_lab = g.predict(np.array(xs).reshape(-1, 1))
    assert [rank_ours[int(l)] for l in res["label"]] == [rank_sk[int(l)] for l in sk_lab]
