"""SURVEY.md section 8f ("next") rows.  These GPU tests were written in round 1 after GPU access had been closed
for the round, so the code under them has only been checked on the host (formula order against scipy / einsum).
They are non-strict xfail until a GPU run has confirmed them: a pass shows up as XPASS, a failure does not mask
the verified suite.  The file sorts last on purpose (the driver runs pytest with -x).
"""
import warnings

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason='added in round 1 without GPU access; to be confirmed on a B200')]


def test_gaussian_proxy_on_device_matches_scipy(gpu_device, g45):
    from scipy.stats import multivariate_normal as mvn
    from gradient_free_mcmc_postprocessing_b200.thinning import gaussian_proxy_device
    x, mean, cov = g45['sample'], g45['sample_mean'], g45['sample_cov']
    _, log_q, grad_q = gaussian_proxy_device(x, mean, cov, device=gpu_device)
    np.testing.assert_allclose(log_q.cpu().numpy(), mvn.logpdf(x, mean=mean, cov=cov), rtol=1e-13, atol=1e-13)
    want = -np.einsum('ij,kj->ki', np.linalg.inv(cov), x - mean)          # code/src/thinning.py:16
    np.testing.assert_allclose(grad_q.cpu().numpy(), want, rtol=1e-13, atol=1e-14)
    rng = np.random.default_rng(3)
    d = 10
    a = rng.standard_normal((d, d))
    cov = a @ a.T / d + 0.5 * np.eye(d)
    mean = rng.standard_normal(d)
    x = rng.standard_normal((5000, d)) @ np.linalg.cholesky(cov).T + mean
    _, log_q, grad_q = gaussian_proxy_device(x, mean, cov, device=gpu_device)
    np.testing.assert_allclose(log_q.cpu().numpy(), mvn.logpdf(x, mean=mean, cov=cov), rtol=1e-12)
    np.testing.assert_allclose(grad_q.cpu().numpy(), -np.einsum('ij,kj->ki', np.linalg.inv(cov), x - mean),
                               rtol=1e-12, atol=1e-13)


def test_gaussian_thin_with_device_proxy_matches_reference_glue(gpu_device, g45):
    from gradient_free_mcmc_postprocessing_b200.thinning import gaussian_thin
    idx = gaussian_thin(g45['sample'], g45['log_p'], g45['sample_mean'], g45['sample_cov'], 50,
                        proxy_on_device=True, device=gpu_device)
    assert np.array_equal(idx, g45['glue_gaussian_thin_50'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        idx = gaussian_thin(g45['sample'], g45['log_p'], g45['laplace_mean'], g45['laplace_cov'], 30,
                            proxy_on_device=True, device=gpu_device)
    assert np.array_equal(idx, g45['glue_gaussian_thin_laplace_30'])
