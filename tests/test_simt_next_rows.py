"""SURVEY.md section 8f ("next") kernels - Student-t proxy, Gaussian KDE proxy, cumulative energy distance -
executed by the SIMT emulator (tests/host/simt) and compared with the oracle / scipy.  See tests/test_simt_kernels.py
for what the emulator does and does not prove."""
import numpy as np
import pytest
from scipy import stats
from scipy.stats import gaussian_kde

import simt_harness as H
from oracle import next_rows as nr
from oracle.stein_thinning_oracle import thinning as ot


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ('GFK_PERSISTENT', 'GFK_RESIDENT', 'SIMT_SMS', 'SIMT_WORKERS'):
        monkeypatch.delenv(k, raising=False)
    H.load()
    H.set_schedule(False)
    yield
    H.set_schedule(False)


@pytest.mark.parametrize('d', [1, 2, 4, 9])
def test_student_t_proxy_matches_scipy_and_the_notebook(d):
    rng = np.random.default_rng(40 + d)
    x = rng.standard_normal((1111, d)) * 2 + 1
    a = rng.standard_normal((d, d))
    sigma = a @ a.T / d + 0.5 * np.eye(d)
    mu = rng.standard_normal(d)
    for df in (2.5, 9.0):
        log_q, grad_q = H.student_t_proxy(x, mu, sigma, df)
        np.testing.assert_allclose(log_q, stats.multivariate_t.logpdf(x, loc=mu, shape=sigma, df=df), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(log_q, nr.t_log_pdf(x, mu, sigma, df), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(grad_q, nr.t_grad_log_pdf(x, mu, sigma, df), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize('sms', [1, 4])
def test_kde_proxy_golden_unique_counts(g45, sms, monkeypatch):
    """Gaussian_mixture.ipynb#c67: thinning with the KDE / weighted-KDE proxy selects 481 / 343 unique points.
    sms = 1 / 4 changes how many ways the data range is split and merged."""
    monkeypatch.setenv('SIMT_SMS', str(sms))
    x, log_p = g45['sample'], g45['log_p']
    log_q, grad_q = H.kde_proxy(x)
    kde = gaussian_kde(x.T, bw_method='silverman')
    np.testing.assert_allclose(log_q, kde.logpdf(x.T), rtol=1e-12, atol=1e-12)
    want_lq, want_g = nr.kde_logpdf_and_score(x)
    np.testing.assert_allclose(grad_q, want_g, rtol=1e-10, atol=1e-11)
    idx, _ = H.SimProblem(x, grad_q, log_p, log_q, preconditioner='med').thin(1000)
    assert len(np.unique(idx)) == 481
    assert np.array_equal(idx, ot.thin_gf(x, log_p, want_lq, want_g, 1000, preconditioner='med'))
    w = np.exp(log_p) / np.sum(np.exp(log_p))                              # #c51
    log_q, grad_q = H.kde_proxy(x, weights=w)
    np.testing.assert_allclose(log_q, gaussian_kde(x.T, bw_method='silverman', weights=w).logpdf(x.T), rtol=1e-12, atol=1e-12)
    idx, _ = H.SimProblem(x, grad_q, log_p, log_q, preconditioner='med').thin(1000)
    assert len(np.unique(idx)) == 343


@pytest.mark.parametrize('d', [1, 3, 4, 8])
def test_kde_proxy_at_other_points_and_far_tails(d):
    rng = np.random.default_rng(50 + d)
    n, m = 777, 301                                       # neither a multiple of the 128-row tiles
    data = rng.standard_normal((n, d)) @ (np.eye(d) + 0.4 * rng.standard_normal((d, d)))
    pts = rng.standard_normal((m, d)) * 3
    pts[:5] += 40.0                                       # far tail: every exp(energy) underflows without the running max
    w = rng.random(n)
    for weights in (None, w):
        log_q, grad_q = H.kde_proxy(data, weights=weights, points=pts)
        want_lq, want_g = nr.kde_logpdf_and_score(data, weights=weights, points=pts)
        assert np.all(np.isfinite(log_q)) and np.all(np.isfinite(grad_q))
        np.testing.assert_allclose(log_q, want_lq, rtol=1e-12, atol=1e-11)
        # the oracle's x-space quadratic form (energies up to ~1e4 through an ill-conditioned H^-1) is the less
        # accurate side here: against long-double arithmetic the kernel is within 1e-10, the oracle within 1e-8
        np.testing.assert_allclose(grad_q, want_g, rtol=2e-8, atol=1e-10)
    log_q, _ = H.kde_proxy(data, points=pts, bw_method='scott')
    np.testing.assert_allclose(log_q, gaussian_kde(data.T, bw_method='scott').logpdf(pts.T), rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize('d,n,m', [(1, 50, 7), (2, 1000, 300), (4, 517, 129), (4, 130, 700), (8, 300, 40)])
def test_cumulative_energy_distance_matches_the_prefix_by_prefix_definition(d, n, m):
    """Comparison.ipynb#c19-#c22 evaluates dcor.energy_distance(validation, sample[idx[:t]]) from scratch per t."""
    rng = np.random.default_rng(60 + d)
    x = rng.standard_normal((n, d))
    y = rng.standard_normal((m, d)) * 1.3 + 0.2
    y[m // 2] = y[0]                                      # a repeated selection (thinning selects with replacement)
    e2 = H.energy_cumulative(x, y)
    for t in sorted({0, 1, 2, m // 3, m - 2, m - 1} & set(range(m))):
        want = nr.energy_distance(x, y[:t + 1])
        assert abs(e2[t] - want) <= 1e-12 * max(1.0, abs(want)), (t, e2[t], want)
    assert abs(np.sqrt(e2[-1]) - nr.fit_quality(x, y, validation_sample_step=1)) < 1e-12
    assert abs(H.energy_cumulative(x, x)[-1]) < 1e-12     # identical samples: distance 0


@pytest.mark.parametrize('seed', [1, 2])
def test_next_row_kernels_do_not_depend_on_thread_schedule(seed):
    rng = np.random.default_rng(70)
    data, pts = rng.standard_normal((400, 3)), rng.standard_normal((150, 3))
    base_kde = H.kde_proxy(data, points=pts)
    base_e = H.energy_cumulative(data, pts)
    H.set_schedule(True, seed)
    got = H.kde_proxy(data, points=pts)
    assert np.array_equal(got[0], base_kde[0]) and np.array_equal(got[1], base_kde[1])
    assert np.array_equal(H.energy_cumulative(data, pts), base_e)
