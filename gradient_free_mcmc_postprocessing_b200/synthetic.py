"""Seeded synthetic inputs of the BASELINE.json config shapes (numpy only; no jax).

``make_mvn_mixture`` restates the reference's test-input generator
(``code/src/utils/mvn.py:7-40``: ``rvs`` / ``logpdf`` / ``score``) with the same random-number
consumption order, so seed 12345 reproduces the sample of
``code/notebooks/gaussian_mixture/Gaussian_mixture.ipynb#c5-#c10``.  ``config_c1`` .. ``config_c5`` build
the five BASELINE.json configs as described in SURVEY.md section 8d (the real Lotka-Volterra chains
live in a private bucket, so C2 is an emulation with MCMC-like duplicated rows).
"""
import numpy as np
from scipy.special import logsumexp
from scipy.stats import multivariate_normal as mvn


def make_mvn_mixture(weights, means, covs):
    weights, means, covs = np.asarray(weights, float), np.asarray(means, float), np.asarray(covs, float)
    k, d = means.shape
    assert weights.shape == (k,) and covs.shape == (k, d, d)
    covs_inv = np.linalg.inv(covs)

    def rvs(size, random_state):
        draws = [mvn.rvs(mean=means[i], cov=covs[i], size=size, random_state=random_state)
                 for i in range(k)]
        which = random_state.choice(k, size=size, p=weights)
        stacked = np.stack(draws, axis=1).reshape(size, k, d)
        return np.take_along_axis(stacked, which.reshape(size, 1, 1), axis=1).reshape(size, d)

    def _log_components(x):
        return np.stack([mvn.logpdf(x, mean=means[i], cov=covs[i]) for i in range(k)]).reshape(k, -1)

    def logpdf(x):
        return logsumexp(_log_components(x), axis=0, b=weights[:, None])

    def score(x):
        logc = _log_components(x) + np.log(weights)[:, None]
        resp = np.exp(logc - logsumexp(logc, axis=0))                       # (k, n) responsibilities
        centred = x[np.newaxis, :, :] - means[:, np.newaxis, :]               # (k, n, d)
        return -np.einsum('il,ijk,ilk->lj', resp, covs_inv, centred)

    return rvs, logpdf, score


def gaussian_proxy(sample, ddof=1):
    """Moment-matched Gaussian proxy: (log_q, gradient_q, mean, cov) as in Gaussian_mixture.ipynb#c33-#c37."""
    mean = np.mean(sample, axis=0)
    cov = np.atleast_2d(np.cov(sample, rowvar=False, ddof=ddof))
    log_q = mvn.logpdf(sample, mean=mean, cov=cov)
    gradient_q = -np.einsum('ij,kj->ki', np.linalg.inv(cov), sample - mean)
    return log_q, gradient_q, mean, cov


C1_WEIGHTS = np.array([0.3, 0.7])
C1_MEANS = np.array([[-1., -1.], [1., 1.]])
C1_COVS = np.array([[[0.5, 0.25], [0.25, 1.]],
                    [[2.0, -np.sqrt(3.) * 0.8], [-np.sqrt(3.) * 0.8, 1.5]]])


def config_c1(n=10_000, seed=12345):
    """C1: bivariate 2-component mixture, GF thinning with a moment-matched Gaussian proxy, 'med'."""
    rvs, logpdf, score = make_mvn_mixture(C1_WEIGHTS, C1_MEANS, C1_COVS)
    sample = rvs(n, np.random.default_rng(seed))
    log_q, gradient_q, _, _ = gaussian_proxy(sample)
    return dict(kind='gf', sample=sample, log_p=logpdf(sample), log_q=log_q, gradient_q=gradient_q,
                gradient=score(sample), m=100, preconditioner='med')


def _mcmc_like_repeats(n_unique_draw, n, rng, accept=0.23):
    """Indices that repeat each proposal a geometric number of times (random-walk acceptance ~ accept)."""
    reps = rng.geometric(accept, size=n_unique_draw)
    idx = np.repeat(np.arange(n_unique_draw), reps)[:n]
    assert idx.shape[0] == n
    return idx


def config_c2(n=200_000, d=4, seed=12345, m=1000):
    """C2: Lotka-Volterra-shaped emulation (d=4): tight correlated Gaussian posterior, scores of
    magnitude 1e2-1e3, rows duplicated like a random-walk chain (exact ties), score-based 'med'."""
    rng = np.random.default_rng(seed)
    mean = np.array([0.7, 1.1, 0.9, 1.2])[:d]
    a = rng.standard_normal((d, d))
    cov = 2e-4 * (a @ a.T / d + 0.5 * np.eye(d))
    n_unique = int(n * 0.35) + 1000
    base = mvn.rvs(mean=mean, cov=cov, size=n_unique, random_state=rng).reshape(n_unique, d)
    sample = base[_mcmc_like_repeats(n_unique, n, rng)]
    gradient = np.einsum('jk,ik->ij', np.linalg.inv(cov), mean - sample)
    return dict(kind='score', sample=np.ascontiguousarray(sample), gradient=np.ascontiguousarray(gradient),
                m=m, preconditioner='med')


def _random_mixture(d, k, rng, spread=2.0):
    weights = rng.dirichlet(np.full(k, 5.0))
    means = rng.standard_normal((k, d)) * spread
    covs = []
    for _ in range(k):
        a = rng.standard_normal((d, d))
        covs.append(a @ a.T / d + 0.5 * np.eye(d))
    return weights, means, np.array(covs)


def config_c3(n=1_000_000, d=4, seed=12345, m=1000, k=3, proxy='wide-mixture', inflation=1.3):
    """C3: d=4 Gaussian mixture, gradient-free thinning, 'med'.

    proxy='wide-mixture' (default): q is the same mixture with every covariance inflated by `inflation` - a proxy
    with heavier tails than the target, under which the selections are all distinct.  proxy='gaussian': the
    moment-matched Gaussian of SURVEY.md section 8d; at n >= 10^5 it makes gradient-free thinning degenerate (the
    log-weights spread over ~15 and the search keeps returning the same one or two points - the pathology the
    reference shows for its Laplace proxy, Gaussian_mixture.ipynb#c110-#c111), so index parity on it says little.
    The arithmetic per greedy step is the same for both."""
    rng = np.random.default_rng(seed)
    weights, means, covs = _random_mixture(d, k, rng, spread=1.0)
    rvs, logpdf, score = make_mvn_mixture(weights, means, covs)
    sample = rvs(n, rng)
    if proxy == 'gaussian':
        log_q, gradient_q, _, _ = gaussian_proxy(sample)
    elif proxy == 'wide-mixture':
        _, logpdf_q, score_q = make_mvn_mixture(weights, means, covs * inflation)
        log_q, gradient_q = logpdf_q(sample), score_q(sample)
    else:
        raise ValueError("proxy must be 'wide-mixture' or 'gaussian'")
    return dict(kind='gf', sample=sample, log_p=logpdf(sample), log_q=log_q, gradient_q=gradient_q,
                m=m, preconditioner='med')


def config_c4(n=200_000, d=10, seed=12345, k=3):
    """C4: d=10 Gaussian mixture for the full O(n^2) GF-KSD with the median-heuristic preconditioner."""
    rng = np.random.default_rng(seed)
    weights, means, covs = _random_mixture(d, k, rng, spread=0.7)
    rvs, logpdf, score = make_mvn_mixture(weights, means, covs)
    sample = rvs(n, rng)
    log_q, gradient_q, _, _ = gaussian_proxy(sample)
    return dict(kind='gf-ksd', sample=sample, log_p=logpdf(sample), log_q=log_q, gradient_q=gradient_q,
                gradient=score(sample), preconditioner='med')


def config_c5(n=500_000, d=64, seed=12345, m=500):
    """C5: d=64 two-component standard-normal-plus-shift mixture, score-based thinning, 'med'."""
    rng = np.random.default_rng(seed)
    shift = np.zeros(d)
    shift[:4] = 1.5
    weights = np.array([0.5, 0.5])
    means = np.stack([-shift, shift])
    covs = np.stack([np.eye(d), np.eye(d)])
    comp = rng.choice(2, size=n, p=weights)
    sample = rng.standard_normal((n, d)) + means[comp]
    # analytic score of the equal-covariance mixture
    logc = np.stack([-0.5 * np.sum((sample - means[i]) ** 2, axis=1) for i in range(2)]) + np.log(weights)[:, None]
    resp = np.exp(logc - logsumexp(logc, axis=0))
    gradient = -(sample - (resp[0][:, None] * means[0] + resp[1][:, None] * means[1]))
    return dict(kind='score', sample=sample, gradient=gradient, m=m, preconditioner='med')


CONFIGS = {'c1': config_c1, 'c2': config_c2, 'c3': config_c3, 'c4': config_c4, 'c5': config_c5}
