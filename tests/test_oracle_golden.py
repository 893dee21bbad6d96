"""The oracle against every golden vector the reference holds for this path.
CPU only.  These pin the oracle; the GPU tests then compare libgfk with the oracle."""
import warnings

import numpy as np
import pytest

from oracle import repo_glue
from oracle.stein_thinning_oracle import kernel as ok, stein as os_, thinning as ot


def test_g1_inputs_regenerate(g13):
    # Gradient_free_Stein_thinning.ipynb#c17,#c18 (8 printed digits)
    assert np.allclose(g13['sample_mean'], g13['nb_sample_mean'], atol=5e-9)
    assert np.allclose(g13['sample_cov'], g13['nb_sample_cov'], atol=5e-9)


def test_g1_thin(g13):
    idx = ot.thin(g13['sample'], g13['gradient'], 20)
    assert idx.dtype == np.uint32
    assert np.array_equal(idx, g13['idx_thin'])          # #c8 stored output


def test_g2_thin_gf_equals_thin_when_q_is_p(g13):
    idx = ot.thin_gf(g13['sample'], g13['log_p'], g13['log_p'], g13['gradient'], 20)
    assert np.array_equal(idx, g13['idx_thin'])          # #c14 assertion


def test_g3_thin_gf(g13):
    idx = ot.thin_gf(g13['sample'], g13['log_p'], g13['log_q'], g13['gradient_q'], 20)
    assert np.array_equal(idx, g13['idx_thin_gf'])       # #c21 stored output


def test_g1_fails_without_standardisation(g13):
    # pins the mean-absolute-deviation standardisation
    idx = ot.thin(g13['sample'], g13['gradient'], 20, standardize=False)
    assert not np.array_equal(idx, g13['idx_thin'])


def test_g4_inputs_regenerate(g45):
    assert np.allclose(g45['sample_mean'], g45['nb_sample_mean'], atol=5e-9)
    assert np.allclose(g45['sample_cov'], g45['nb_sample_cov'], atol=5e-9)
    assert np.allclose(g45['laplace_mean'], g45['nb_laplace_mean'], atol=1e-6)


def test_g4_unique_counts(g45):
    idx = ot.thin(g45['sample'], g45['gradient'], 1000, preconditioner='med')
    assert len(np.unique(idx)) == int(g45['nb_unique_thin_med'])         # 538, Gaussian_mixture.ipynb#c67
    idx = ot.thin_gf(g45['sample'], g45['log_p'], g45['log_q'], g45['gradient_q'], 1000, preconditioner='med')
    assert len(np.unique(idx)) == int(g45['nb_unique_thin_gf_med'])      # 400


def test_g5_degenerate_laplace_proxy(g45):
    with pytest.warns(UserWarning, match='log_q differs from log_p by more than 10'):
        idx = ot.thin_gf(g45['sample'], g45['log_p'], g45['log_q_laplace'], g45['gradient_q_laplace'], 300,
                         preconditioner='med')
    assert np.array_equal(np.unique(idx), [int(g45['nb_stuck_index'])])  # 494, #c111


def test_g6_loose_scalars(g45):
    lw10 = (g45['log_q_laplace'] - g45['log_p']) / np.log(10)
    assert abs(np.min(lw10) - float(g45['nb_min_log10_q_over_p'])) < 5e-3   # BFGS-version sensitive
    km = os_.kmat(ot._make_stein_integrand(g45['sample'], g45['gradient_q_laplace']), 1000)
    assert abs(np.max(np.abs(km)) / float(g45['nb_max_abs_kmat']) - 1) < 1e-3


def test_reference_test_ksd_contract(ref_test_ksd):
    # code/tests/test_ksd.py:8-28 with the oracle's kmat / reindex_integrand
    mat = ref_test_ksd['mat']
    res = os_.kmat(repo_glue.reindex_integrand(lambda a, b: mat[a, b], ref_test_ksd['indices']), mat.shape[0])
    np.testing.assert_array_equal(res, ref_test_ksd['expected'])


def test_glue_matches_reference_glue(g45):
    # fixtures were produced by the reference's own code/src/thinning.py and code/src/utils/ksd.py
    idx = repo_glue.gaussian_thin(g45['sample'], g45['log_p'], g45['sample_mean'], g45['sample_cov'], 50)
    assert np.array_equal(idx, g45['glue_gaussian_thin_50'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        idx = repo_glue.gaussian_thin(g45['sample'], g45['log_p'], g45['laplace_mean'], g45['laplace_cov'], 30)
    assert np.array_equal(idx, g45['glue_gaussian_thin_laplace_30'])
    ks = repo_glue.calculate_ksd(g45['sample'], g45['gradient'], g45['glue_idx_for_ksd'])
    np.testing.assert_array_equal(ks, g45['glue_calculate_ksd'])


def test_ksd_matches_kmat_definition(g13):
    integrand = ot._make_stein_integrand(g13['sample'], g13['gradient'])
    km = os_.kmat(integrand, 60)
    ks = os_.ksd(integrand, 60)
    expect = np.array([np.sqrt(km[:t + 1, :t + 1].sum()) / (t + 1) for t in range(60)])
    assert np.allclose(ks, expect, rtol=1e-12)


def test_selection_with_replacement_and_validation(g13):
    s, g = g13['sample'][:30], g13['gradient'][:30]
    idx = ot.thin(s, g, 100)                 # m > n: selection is with replacement
    assert idx.shape == (100,) and idx.max() < 30
    with pytest.raises(ValueError):
        ot.thin(s, g[:-1], 5)
    bad = s.copy(); bad[3, 0] = np.nan
    with pytest.raises(ValueError):
        ot.thin(bad, g, 5)
    with pytest.raises(Exception):
        ot.thin(np.ones((10, 2)), np.ones((10, 2)), 3, standardize=False, preconditioner='med')
    with pytest.raises(ValueError):
        ok.make_precon(s, 'nonsense')


def test_preconditioners(g13):
    s = g13['sample']
    for pre, d in (('id', 2), ('med', 2), ('sclmed', 2), ('smpcov', 2), ('2.5', 2)):
        linv = ok.make_precon(s, pre)
        assert linv.shape == (d, d) and np.all(np.linalg.eigvals(linv) > 0)
    assert np.allclose(ok.make_precon(s, '2.5'), np.identity(2) / 2.5)


# ---- the oracle of the steps either side of the path ----------------------------------------------
def test_kde_proxy_golden_unique_counts(g45):
    """Gaussian_mixture.ipynb#c42-#c54,#c67: thin_gf with the KDE / weighted-KDE proxy selects 481 / 343 unique
    points of the seeded mixture sample - pins oracle.next_rows.kde_logpdf_and_score."""
    from oracle import next_rows as nr
    x, log_p = g45['sample'], g45['log_p']
    log_q, gradient_q = nr.kde_logpdf_and_score(x)
    idx = ot.thin_gf(x, log_p, log_q, gradient_q, 1000, preconditioner='med')
    assert len(np.unique(idx)) == 481
    p = np.exp(log_p)
    log_q, gradient_q = nr.kde_logpdf_and_score(x, weights=p / np.sum(p))
    idx = ot.thin_gf(x, log_p, log_q, gradient_q, 1000, preconditioner='med')
    assert len(np.unique(idx)) == 343


def test_student_t_oracle_matches_scipy_like_the_notebook():
    """Gradient_free_Student_t.ipynb#c30: np.testing.assert_allclose(t_log_pdf(...), multivariate_t.logpdf(...))."""
    from scipy import stats
    from oracle import next_rows as nr
    rng = np.random.default_rng(2)
    x = rng.standard_normal((500, 4)) * 0.02 + np.array([-0.38, 0.29, -0.02, -0.018])
    a = rng.standard_normal((4, 4))
    scale = 1e-4 * (a @ a.T / 4 + 0.5 * np.eye(4))
    mu = x.mean(0)
    np.testing.assert_allclose(nr.t_log_pdf(x, mu, scale, 9.0), stats.multivariate_t.logpdf(x, loc=mu, shape=scale, df=9.0))
    eps = 1e-7
    num = np.array([(nr.t_log_pdf(x + eps * np.eye(4)[k], mu, scale, 9.0) - nr.t_log_pdf(x - eps * np.eye(4)[k], mu, scale, 9.0))
                    / (2 * eps) for k in range(4)]).T
    np.testing.assert_allclose(nr.t_grad_log_pdf(x, mu, scale, 9.0), num, rtol=1e-5, atol=1e-3)
