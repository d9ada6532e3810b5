"""CPU restatement of the 1-D allele phasing (TEST INFRASTRUCTURE ONLY: tests/, smoke() and bench.py's CPU legs).

SURVEY.md section 8(f) row f3 -- what the reference does after round 3 for every region:

    outlier_cutoffs      <- reference src/NanoRepeat/split_alleles.py:98-113   mean +- 3 sd, lower bound floored at 0
    trim                 <- split_alleles.py:141-154                            keeps min <= size <= max
    em_fit               <- sklearn.mixture.GaussianMixture(covariance_type='diag', tol=1e-3, reg_covar=1e-6,
                            max_iter=100): the E / M steps and the stopping rule of BaseMixture.fit_predict
                            (scikit-learn 1.9: E step, M step, |change of mean log-likelihood| < tol), started from
                            explicit parameters
    overlap              <- split_alleles.py:90-96, :171-200                    [isf(1-o), isf(o)] intervals, sd floored at 1.0
    auto_gmm             <- split_alleles.py:171-200                            n = 2, 3, ... until two components overlap
    labels               <- split_alleles.py:258-279 (predict / predict_proba on the trimmed sizes)

Pinned: em_fit against scikit-learn itself from identical starting parameters (tests/test_oracle_gmm.py, 1e-9), the
overlap rule against scipy.stats.norm.isf.  NOT pinned, by construction: the reference draws its bootstrap noise
(split_alleles.py:82-88, random.gauss) and sklearn's k-means starts from unseeded global generators, so no two runs of
the reference agree bit for bit.  The repo replaces both draws by a counter-based generator (`mix64` below) that this
file and the CUDA kernel (nanorepeat_b200/csrc/nr_gmm.cu) share, so THEY agree to rounding, and holds the result to the
reference statistically (same number of alleles, same labels, means within the bootstrap's standard error).
"""
import math

import numpy as np

MASK = (1 << 64) - 1
LOG_2PI = math.log(2.0 * math.pi)
BOOTSTRAP = 100                    # split_alleles.py:83


def mix64(x):
    """splitmix64's output function on a Python int."""
    x = (x + 0x9E3779B97F4A7C15) & MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK
    return x ^ (x >> 31)


def key(seed, region, stream, idx):
    return mix64((mix64((mix64(seed ^ ((0xD1B54A32D192ED03 * (region + 1)) & MASK)) + stream) & MASK) + idx) & MASK)


def uniform(h):
    """(0, 1]"""
    return ((h >> 11) + 1) * (1.0 / 9007199254740992.0)


def gauss(seed, region, idx):
    u1, u2 = uniform(key(seed, region, 1, idx)), uniform(key(seed, region, 2, idx))
    return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


def outlier_cutoffs(sizes):
    mean, sd = float(np.mean(sizes)), float(np.std(sizes))
    return max(mean - 3 * sd, 0.0), mean + 3 * sd


def trim(sizes):
    """-> indices kept (split_alleles.py:141-154)"""
    lo, hi = outlier_cutoffs(sizes)
    return [i for i, v in enumerate(sizes) if not (v < lo or v > hi)]


def bootstrap(sizes, error_rate, seed, region):
    """split_alleles.py:82-88 with the repo's generator: sample rep * n + i = x[i] + N(0, error_rate * (10 + x[i]))."""
    n = len(sizes)
    out = np.empty(BOOTSTRAP * n)
    for t in range(BOOTSTRAP * n):
        x = float(sizes[t % n])
        out[t] = x + error_rate * (10.0 + x) * gauss(seed, region, t)
    return out


def start_parameters(x, n, init, seed, region, lloyd=10):
    """The repo's stand-in for sklearn's k-means start: init 0 spreads the means evenly over [min, max], the others take
    hashed sample positions; `lloyd` rounds of 1-D k-means; then sklearn's own first M step on the one-hot labels."""
    x = np.asarray(x, dtype=np.float64)
    if init == 0:
        lo, hi = float(x.min()), float(x.max())
        means = np.array([lo + (j + 0.5) / n * (hi - lo) for j in range(n)])
    else:
        means = np.array([x[key(seed, region, 1000 + 32 * n + init, j) % len(x)] for j in range(n)])
    for _ in range(lloyd):
        lab = np.argmin(np.abs(x[:, None] - means[None, :]), axis=1)
        for j in range(n):
            if np.any(lab == j):
                means[j] = x[lab == j].sum() / np.count_nonzero(lab == j)
    lab = np.argmin(np.abs(x[:, None] - means[None, :]), axis=1)
    resp = np.zeros((len(x), n))
    resp[np.arange(len(x)), lab] = 1.0
    return m_step(x, resp)


def m_step(x, resp, reg_covar=1e-6):
    """sklearn _estimate_gaussian_parameters, 'diag', one feature."""
    nk = resp.sum(axis=0) + 10 * np.finfo(np.float64).eps
    means = resp.T @ x / nk
    var = resp.T @ (x * x) / nk - means ** 2 + reg_covar
    return nk / len(x), means, var


def e_step(x, weights, means, var):
    """-> (mean log-likelihood, log responsibilities)"""
    pc = 1.0 / np.sqrt(var)
    lp = -0.5 * (LOG_2PI + (x[:, None] * pc[None, :] - (means * pc)[None, :]) ** 2) + np.log(pc)[None, :] + np.log(weights)[None, :]
    top = lp.max(axis=1)
    norm = top + np.log(np.exp(lp - top[:, None]).sum(axis=1))
    return float(norm.mean()), lp - norm[:, None]


def em_fit(x, weights, means, var, tol=1e-3, max_iter=100, reg_covar=1e-6):
    """-> (weights, means, variances, lower bound, iterations, converged)"""
    x = np.asarray(x, dtype=np.float64)
    lower, converged, it = -np.inf, False, 0
    for it in range(1, max_iter + 1):
        prev = lower
        lower, log_resp = e_step(x, weights, means, var)
        weights, means, var = m_step(x, np.exp(log_resp), reg_covar)
        if abs(lower - prev) < tol:
            converged = True
            break
    return weights, means, var, lower, it, converged


def best_fit(x, n, seed, region, n_init=10):
    best = None
    for init in range(n_init):
        fit = em_fit(x, *start_parameters(x, n, init, seed, region))
        if best is None or fit[3] > best[3]:
            best = fit
    order = np.argsort(best[1], kind="stable")              # components by ascending mean (the library's output order)
    return (best[0][order], best[1][order], best[2][order]) + tuple(best[3:])


def std_isf(o):
    """scipy.stats.norm.isf(o) by Newton on erfc (what the library does; the test holds it to scipy)."""
    z = 0.0
    for _ in range(60):
        f = 0.5 * math.erfc(z / math.sqrt(2.0)) - o
        z += f / (math.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi))
    return z


def interval_has_overlap(a, b):
    """split_alleles.py:90-96 (touching intervals overlap)"""
    return max(a[0], b[0]) - min(a[1], b[1]) <= 0


def overlap(means, var, o):
    """split_alleles.py:176-195: do any two components' intervals overlap?"""
    z = std_isf(o)
    n = len(means)
    for i in range(n):
        for j in range(i + 1, n):
            si, sj = max(1.0, math.sqrt(var[i])), max(1.0, math.sqrt(var[j]))
            a = (means[i] - z * si, means[i] + z * si)          # (isf(1 - o), isf(o))
            b = (means[j] - z * sj, means[j] + z * sj)
            if interval_has_overlap(a, b):
                return True
    return False


def one_component(x):
    x = np.asarray(x, dtype=np.float64)
    return m_step(x, np.ones((len(x), 1)))


def auto_gmm(x, max_components, o, seed, region, n_init=10):
    """split_alleles.py:171-200 -> (n, weights, means, variances).  (The reference refits n - 1 from fresh random starts
    when n overlaps; the repo keeps the n - 1 fit it already has -- same model family, same data.)"""
    prev = one_component(x)
    for n in range(2, max_components + 1):
        w, m, v = best_fit(x, n, seed, region, n_init)[:3]
        if overlap(m, v, o):
            return (n - 1,) + tuple(prev)
        prev = (w, m, v)
    return (max_components,) + tuple(prev)


def labels(sizes, weights, means, var):
    """predict / predict_proba of the fitted mixture on the trimmed sizes -> (label, probability of that label)"""
    _ll, log_resp = e_step(np.asarray(sizes, dtype=np.float64), np.asarray(weights), np.asarray(means), np.asarray(var))
    lab = np.argmax(log_resp, axis=1)
    return lab, np.exp(log_resp[np.arange(len(lab)), lab])


def phase_1d(sizes, error_rate, max_components, o, seed, region=0, n_init=10):
    """split_allele_using_gmm_1d (nanoRepeat_bam.py:515-575) up to the labels: -> dict(kept, n, weights, means, variances,
    label, proba) with label / proba per kept size."""
    kept = trim(sizes) if len(sizes) >= 2 else []
    if not kept:
        return dict(kept=kept, n=0, weights=[], means=[], variances=[], label=[], proba=[])
    xs = [sizes[i] for i in kept]
    n, w, m, v = auto_gmm(bootstrap(xs, error_rate, seed, region), max_components, o, seed, region, n_init)
    lab, pr = labels(xs, w, m, v)
    return dict(kept=kept, n=n, weights=list(w), means=list(m), variances=list(v), label=list(lab), proba=list(pr))
