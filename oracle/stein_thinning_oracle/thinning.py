"""Oracle restatement of ``stein_thinning.thinning`` (TEST INFRASTRUCTURE ONLY).

Sources followed (paths under ``/root/reference``):

* ``_greedy_search`` - restated verbatim (minus logging) at
  ``code/notebooks/examples/JAX_Stein_Thinning.ipynb#c22``; algorithm at
  ``report/report.tex:413-426``.
* integrand wiring - ``JAX_Stein_Thinning.ipynb#c16-#c17``; GF weights
  ``report/report.tex:373,390-399``; signature prefixes at
  ``code/src/utils/ksd.py:25`` and ``code/src/thinning.py:17``.
* ``_validate_and_standardize`` - call site ``JAX_Stein_Thinning.ipynb#c16``;
  scale = mean absolute deviation per column (SURVEY.md section 8a row a3;
  the golden indices G1-G3 reproduce with it and fail without it).
* weight-spread warning text - stored output of
  ``code/notebooks/gaussian_mixture/Gaussian_mixture.ipynb#c62``.

PARITY UNPINNED: ``range_cap``.  Definition used here:
``lw = log_q - log_p`` is shifted so its maximum is 0 and floored at
``-range_cap``; with ``range_cap=None`` ``lw`` is used as is.
"""
import warnings

import numpy as np

from .kernel import make_imq

WEIGHT_SCALE_THRESHOLD = 10


def _validate_and_standardize(sample, gradient, standardize):
    sample = np.asarray(sample)
    gradient = np.asarray(gradient)
    if sample.ndim != 2:
        raise ValueError('sample is not two-dimensional.')
    n, d = sample.shape
    if n == 0 or d == 0:
        raise ValueError('sample is empty.')
    if gradient.shape != (n, d):
        raise ValueError('Dimensions of sample and gradient are inconsistent.')
    if np.any(np.isnan(sample)) or np.any(np.isinf(sample)):
        raise ValueError('sample contains NaNs or infs.')
    if np.any(np.isnan(gradient)) or np.any(np.isinf(gradient)):
        raise ValueError('gradient contains NaNs or infs.')
    if standardize:
        loc = np.mean(sample, axis=0)
        scl = np.mean(np.abs(sample - loc), axis=0)
        if not np.min(scl) > 0:
            raise ValueError('Too few unique samples in sample.')
        return sample / scl, gradient * scl
    return sample, gradient


def _make_stein_integrand(sample, gradient, standardize=True, preconditioner='id'):
    s, g = _validate_and_standardize(sample, gradient, standardize)
    vfk0 = make_imq(s, preconditioner)

    def integrand(ind1, ind2):
        return vfk0(s[ind1], s[ind2], g[ind1], g[ind2])
    return integrand


def _log_weights(log_p, log_q, n, range_cap):
    log_p = np.asarray(log_p, dtype=float)
    log_q = np.asarray(log_q, dtype=float)
    if log_p.shape != (n,) or log_q.shape != (n,):
        raise ValueError('Dimensions of sample and log-densities are inconsistent.')
    if np.any(np.isnan(log_p)) or np.any(np.isnan(log_q)):
        raise ValueError('log-densities contain NaNs.')
    lw = log_q - log_p
    if np.max(lw) - np.min(lw) > WEIGHT_SCALE_THRESHOLD:
        warnings.warn(f'log_q differs from log_p by more than {WEIGHT_SCALE_THRESHOLD} - '
                      'consider using q that matches target better')
    if range_cap is not None:
        lw = np.maximum(lw - np.max(lw), -range_cap)
    return lw


def _make_stein_gf_integrand(sample, log_p, log_q, gradient_q, standardize=True,
                             preconditioner='id', range_cap=None):
    s, g = _validate_and_standardize(sample, gradient_q, standardize)
    vfk0 = make_imq(s, preconditioner)
    lw = _log_weights(log_p, log_q, s.shape[0], range_cap)

    def integrand(ind1, ind2):
        return np.exp(lw[ind1] + lw[ind2]) * vfk0(s[ind1], s[ind2], g[ind1], g[ind2])
    return integrand


def _greedy_search(n_points, integrand):
    idx = np.empty(n_points, dtype=np.uint32)
    running = integrand(slice(None), slice(None))      # diagonal k(x_i, x_i)
    idx[0] = np.argmin(running)
    for i in range(1, n_points):
        running += 2 * integrand(slice(None), [idx[i - 1]])
        idx[i] = np.argmin(running)
    return idx


def thin(sample, gradient, n_points, standardize=True, preconditioner='id'):
    integrand = _make_stein_integrand(sample, gradient, standardize, preconditioner)
    return _greedy_search(n_points, integrand)


def thin_gf(sample, log_p, log_q, gradient_q, n_points, standardize=True,
            preconditioner='id', range_cap=None):
    integrand = _make_stein_gf_integrand(sample, log_p, log_q, gradient_q,
                                         standardize, preconditioner, range_cap)
    return _greedy_search(n_points, integrand)
