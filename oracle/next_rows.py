"""Oracle restatement of the steps either side of the hot path - SURVEY.md section 8f (TEST INFRASTRUCTURE ONLY).

Each function follows a cell of the reference's notebooks (paths under ``/root/reference/code/notebooks``):

* ``t_log_pdf`` / ``t_grad_log_pdf`` / ``thin_gf_t`` - ``lotka_volterra/Gradient_free_Student_t.ipynb#c28,#c31,#c40``
* ``kde_logpdf_and_score``                           - ``gaussian_mixture/Gaussian_mixture.ipynb#c42,#c46,#c51``
  (the notebook differentiates ``kde.logpdf`` with jax autodiff; the analytic gradient of the same Gaussian
  mixture is used here) and ``lotka_volterra/KDE.ipynb#c7-#c9``
* ``energy_distance`` / ``fit_quality``              - ``lotka_volterra/Comparison.ipynb#c19``; ``dcor`` is not
  installed here, so its default V-statistic (``report/report.tex:451-455``) is restated with scipy ``cdist``

Pinned by reference golden values: thinning with the KDE / weighted-KDE proxy of the Gaussian-mixture sample selects
481 / 343 unique points (``Gaussian_mixture.ipynb#c67``; ``tests/test_oracle_golden.py``).  ``t_log_pdf`` is checked
against ``scipy.stats.multivariate_t`` as the notebook does (``#c30``).  ``energy_distance`` is PARITY UNPINNED
against ``dcor`` itself (absent); it is the report's formula.
"""
import numpy as np
from scipy.spatial.distance import cdist
from scipy.special import gamma
from scipy.stats import gaussian_kde

from .stein_thinning_oracle.thinning import thin_gf


def t_log_pdf(x, mu, sigma, df):
    d = x.shape[1]
    sigma_inv = np.linalg.inv(sigma)
    x_mu = x - mu
    return (
        np.log(gamma((df + d) / 2))
        - np.log(gamma(df / 2))
        - d * (np.log(df) + np.log(np.pi)) / 2
        - np.log(np.linalg.det(sigma)) / 2
        - (df + d) / 2 * np.log(1 + np.einsum('ij,jk,ik->i', x_mu, sigma_inv, x_mu) / df)
    )


def t_grad_log_pdf(x, mu, sigma, df):
    d = x.shape[1]
    sigma_inv = np.linalg.inv(sigma)
    x_mu = x - mu
    direction_scaled = np.einsum('jk,ik->ij', sigma_inv, x_mu)
    mahalanobis_d = np.einsum('ij,jk,ik->i', x_mu, sigma_inv, x_mu)
    return -(df + d) / df / (1 + mahalanobis_d / df).reshape(-1, 1) * direction_scaled


def thin_gf_t(sample, log_p, t_mu, t_scale, t_df, thinned_size):
    from scipy import stats
    log_q = stats.multivariate_t.logpdf(sample, loc=t_mu, shape=t_scale, df=t_df)
    gradient_q = t_grad_log_pdf(sample, t_mu, t_scale, t_df)
    return thin_gf(sample, log_p, log_q, gradient_q, thinned_size, range_cap=200)


def kde_logpdf_and_score(sample, weights=None, points=None, bw_method='silverman', block=256):
    """``kde = gaussian_kde(sample.T, bw_method, weights)``; returns ``kde.logpdf(points.T)`` and its gradient."""
    kde = gaussian_kde(sample.T, bw_method=bw_method, weights=weights)
    points = sample if points is None else points
    h_inv = np.linalg.inv(kde.covariance)
    w = kde.weights
    log_q = kde.logpdf(points.T)
    grad = np.empty_like(points, dtype=float)
    for b in range(0, points.shape[0], block):
        diff = points[b:b + block, None, :] - sample[None, :, :]
        energy = -0.5 * np.einsum('ijk,kl,ijl->ij', diff, h_inv, diff)
        resp = w[None, :] * np.exp(energy - energy.max(axis=1, keepdims=True))
        resp /= resp.sum(axis=1, keepdims=True)                                     # posterior responsibilities
        grad[b:b + block] = -np.einsum('ij,ijk->ik', resp, diff) @ h_inv
    return log_q, grad


def energy_distance(x, y):
    """dcor.energy_distance(x, y) with its defaults: 2 E|X - Y| - E|X - X'| - E|Y - Y'| over all pairs."""
    return 2 * np.mean(cdist(x, y)) - np.mean(cdist(x, x)) - np.mean(cdist(y, y))


def fit_quality(validation_sample, subsample, validation_sample_step=10):
    return np.sqrt(energy_distance(validation_sample[::validation_sample_step], subsample))
