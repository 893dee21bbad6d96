"""Regenerate the golden fixtures in this directory.  Run ONLY in the authoring
container (it reads ``/root/reference``); the ``.npz`` files it writes are
committed and are what the tests (here and on the GPU box) read.

    python tests/golden/make_golden.py

What comes from where:

* G1-G3  inputs regenerated from the seeded recipe of
  ``code/notebooks/examples/Gradient_free_Stein_thinning.ipynb#c2-#c6,#c12,#c16-#c19``;
  the expected index vectors are the notebook's STORED OUTPUTS (#c8, #c21) and
  the stored input checks (#c17, #c18).
* G4/G5  mixture sample, log-pdf and score produced by the REFERENCE's own
  ``code/src/utils/mvn.py`` (imported from /root/reference with a stub ``jax``
  module, because jax is not installed and only its unused ``logpdf_jax`` needs
  it); expected unique-index counts / stuck index are the stored outputs of
  ``code/notebooks/gaussian_mixture/Gaussian_mixture.ipynb#c34,#c35,#c67,#c111``.
* glue   ``gaussian_thin`` / ``calculate_ksd`` outputs produced by the REFERENCE's
  own ``code/src/thinning.py`` and ``code/src/utils/ksd.py`` imported from
  /root/reference, with ``stein_thinning`` (not installable) resolved to the
  oracle restatement.  These pin the wrappers, not the arithmetic.
* test_ksd  matrices copied as data from ``code/tests/test_ksd.py:9-27``.
"""
import os
import sys
import types
import warnings

import numpy as np
from scipy.optimize import minimize
from scipy.stats import multivariate_normal as mvn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = '/root/reference/code/src'
sys.path.insert(0, ROOT)

from oracle import stein_thinning_oracle as sto  # noqa: E402


def _stub_modules():
    """Make the reference's glue importable: alias stein_thinning -> oracle, stub jax."""
    for name in ('stein_thinning', 'stein_thinning.thinning', 'stein_thinning.stein',
                 'stein_thinning.kernel'):
        pass
    pkg = types.ModuleType('stein_thinning')
    pkg.thinning, pkg.stein, pkg.kernel = sto.thinning, sto.stein, sto.kernel
    sys.modules['stein_thinning'] = pkg
    sys.modules['stein_thinning.thinning'] = sto.thinning
    sys.modules['stein_thinning.stein'] = sto.stein
    sys.modules['stein_thinning.kernel'] = sto.kernel
    jax = types.ModuleType('jax')
    jnp = types.ModuleType('jax.numpy')
    jsp = types.ModuleType('jax.scipy')
    jst = types.ModuleType('jax.scipy.stats')
    jst.multivariate_normal = None
    jax.numpy, jax.scipy, jsp.stats = jnp, jsp, jst
    sys.modules.update({'jax': jax, 'jax.numpy': jnp, 'jax.scipy': jsp, 'jax.scipy.stats': jst})
    sys.path.insert(0, REF_SRC)


def bivariate_gaussian():
    rng = np.random.default_rng(12345)
    means = np.array([0., 0.])
    covs = np.array([[1., 0.8], [0.8, 1.]])
    sample = mvn.rvs(mean=means, cov=covs, size=1000, random_state=rng)
    gradient = np.einsum('kj,ij->ik', np.linalg.inv(covs), means - sample)
    log_p = mvn.logpdf(sample, mean=means, cov=covs)
    sample_mean = np.mean(sample, axis=0)
    sample_cov = np.cov(sample, rowvar=False, ddof=means.shape[0])
    log_q = mvn.logpdf(sample, mean=sample_mean, cov=sample_cov)
    gradient_q = np.einsum('kj,ij->ik', np.linalg.inv(sample_cov), sample_mean - sample)
    np.savez(
        os.path.join(HERE, 'g1_g3_bivariate_gaussian.npz'),
        sample=sample, gradient=gradient, log_p=log_p, log_q=log_q, gradient_q=gradient_q,
        sample_mean=sample_mean, sample_cov=sample_cov,
        # stored notebook outputs
        nb_sample_mean=np.array([-0.01905537, 0.00423809]),
        nb_sample_cov=np.array([[1.00976763, 0.78753313], [0.78753313, 0.9879006]]),
        idx_thin=np.array([164, 60, 859, 821, 66, 870, 327, 885, 677, 944, 111, 601, 950,
                           249, 584, 795, 174, 317, 792, 637], dtype=np.uint32),
        idx_thin_gf=np.array([302, 995, 914, 931, 889, 918, 65, 714, 885, 46, 601, 88, 111,
                              16, 478, 462, 750, 79, 783, 739], dtype=np.uint32),
    )


def gaussian_mixture():
    from utils.mvn import make_mvn_mixture  # the reference's own generator
    weights = np.array([0.3, 0.7])
    means = np.array([[-1., -1.], [1., 1.]])
    covs = np.array([[[0.5, 0.25], [0.25, 1.]],
                     [[2.0, -np.sqrt(3.) * 0.8], [-np.sqrt(3.) * 0.8, 1.5]]])
    rvs, logpdf, score, _ = make_mvn_mixture(weights, means, covs)
    rng = np.random.default_rng(12345)
    sample = rvs(1000, random_state=rng)
    gradient = score(sample)
    log_p = logpdf(sample)
    sample_mean = np.mean(sample, axis=0)
    sample_cov = np.cov(sample, rowvar=False, ddof=1)
    log_q = mvn.logpdf(sample, mean=sample_mean, cov=sample_cov)
    gradient_q = -np.einsum('ij,kj->ki', np.linalg.inv(sample_cov), sample - sample_mean)
    # Laplace proxy, Gaussian_mixture.ipynb#c56-#c61 (BFGS hess_inv is scipy-version sensitive)
    res = minimize(lambda x: -logpdf(x), sample_mean, method='BFGS')
    assert res.success
    laplace_mean, laplace_cov = res.x, res.hess_inv
    log_q_laplace = mvn.logpdf(sample, mean=laplace_mean, cov=laplace_cov)
    gradient_q_laplace = -np.einsum('ij,kj->ki', np.linalg.inv(laplace_cov), sample - laplace_mean)

    # reference glue, executed from /root/reference, arithmetic = oracle
    import thinning as ref_thinning          # code/src/thinning.py
    from utils import ksd as ref_ksd         # code/src/utils/ksd.py
    idx_glue = ref_thinning.gaussian_thin(sample, log_p, sample_mean, sample_cov, 50)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        idx_glue_laplace = ref_thinning.gaussian_thin(sample, log_p, laplace_mean, laplace_cov, 30)
    idx_st_100 = sto.thinning.thin(sample, gradient, 100, preconditioner='med')
    ksd_glue = ref_ksd.calculate_ksd(sample, gradient, idx_st_100)

    np.savez(
        os.path.join(HERE, 'g4_g5_gaussian_mixture.npz'),
        weights=weights, means=means, covs=covs,
        sample=sample, gradient=gradient, log_p=log_p,
        sample_mean=sample_mean, sample_cov=sample_cov, log_q=log_q, gradient_q=gradient_q,
        laplace_mean=laplace_mean, laplace_cov=laplace_cov,
        log_q_laplace=log_q_laplace, gradient_q_laplace=gradient_q_laplace,
        # stored notebook outputs
        nb_sample_mean=np.array([0.37127232, 0.35813142]),
        nb_sample_cov=np.array([[2.51131846, 0.00893545], [0.00893545, 2.16606407]]),
        nb_laplace_mean=np.array([0.96124667, 1.02187361]),
        nb_laplace_cov=np.array([[2.00636653, -1.3359801], [-1.3359801, 1.43121009]]),
        nb_unique_thin_med=np.array(538), nb_unique_thin_gf_med=np.array(400),
        nb_stuck_index=np.array(494),
        nb_min_log10_q_over_p=np.array(-15.400907840699126),
        nb_max_abs_kmat=np.array(356.09025856685986),
        # reference glue over oracle arithmetic
        glue_gaussian_thin_50=idx_glue, glue_gaussian_thin_laplace_30=idx_glue_laplace,
        glue_idx_for_ksd=idx_st_100, glue_calculate_ksd=ksd_glue,
    )


def reference_test_ksd():
    mat = np.array([[1, 2, 3, 4, 5], [2, 6, 7, 8, 9], [3, 7, 10, 11, 12],
                    [4, 8, 11, 13, 14], [5, 9, 12, 14, 15]])
    indices = np.array([2, 0, 4, 3, 1])
    expected = np.array([[10, 3, 12, 11, 7], [3, 1, 5, 4, 2], [12, 5, 15, 14, 9],
                         [11, 4, 14, 13, 8], [7, 2, 9, 8, 6]])
    np.savez(os.path.join(HERE, 'reference_test_ksd.npz'), mat=mat, indices=indices,
             expected=expected)


if __name__ == '__main__':
    _stub_modules()
    bivariate_gaussian()
    gaussian_mixture()
    reference_test_ksd()
    print('golden fixtures written to', HERE)
