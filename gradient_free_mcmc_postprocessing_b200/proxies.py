"""The steps either side of the thinning path, on the device ("next" rows).

* ``t_log_pdf`` / ``t_grad_log_pdf`` / ``thin_gf_t``: the multivariate Student-t proxy of
  ``lotka_volterra/Gradient_free_Student_t.ipynb#c28,#c31,#c40`` (same names and argument order)
* ``kde_logpdf_and_score`` / ``kde_thin``: the (weighted) Gaussian KDE proxy of
  ``gaussian_mixture/Gaussian_mixture.ipynb#c42-#c54`` - ``logpdf_and_score(kde, sample)`` there
* ``energy_distance`` / ``cumulative_energy_distance`` / ``fit_quality``: ``dcor.energy_distance`` as used by
  ``lotka_volterra/Comparison.ipynb#c19-#c23``; the cumulative form returns the value for EVERY prefix of the
  selected points from one pass over the pair distances

The d x d work (inverse, determinant, Cholesky factor, bandwidth factor) stays on the host like the reference's own
``np.linalg.inv``; everything per row / per pair is a libgfk call.  Round 1 status: these kernels have been executed
by the CPU emulator of the test suite only (GPU access was closed); their ``-m gpu`` tests are in
``tests/test_zz_next_rows.py``.
"""
import ctypes
from math import lgamma

import numpy as np
import torch

from . import _lib
from ._engine import _bytes, _ptr, _require_cuda, _stream_ptr, _to_device_f64


def _host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


# ---- Student-t proxy -------------------------------------------------------------------------------------------
def _t_proxy_device(x, mu, sigma, df, device=None):
    lib = _lib.load()
    dev = _require_cuda(device)
    mu = np.atleast_1d(np.asarray(mu, dtype=float))
    sigma = np.atleast_2d(np.asarray(sigma, dtype=float))
    d = mu.shape[0]
    if sigma.shape != (d, d) or tuple(x.shape)[1:] != (d,):
        raise ValueError('inconsistent shapes of x, mu and sigma')
    df = float(df)
    sign, logdet = np.linalg.slogdet(sigma)
    if sign <= 0 or not df > 0:
        raise ValueError('sigma must be positive definite and df positive')
    # scipy.stats.multivariate_t._logpdf: A - B - C - D + E
    cst = lgamma(0.5 * (df + d)) - lgamma(0.5 * df) - d / 2. * np.log(df * np.pi) - 0.5 * logdet
    with torch.cuda.device(dev):
        xd = _to_device_f64(x, dev, 'x')
        n = xd.shape[0]
        mean_dev = torch.from_numpy(mu).to(dev)
        prec_dev = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(sigma))).to(dev)
        log_q = torch.empty(n, dtype=torch.float64, device=dev)
        grad_q = torch.empty((n, d), dtype=torch.float64, device=dev)
        _lib.check(lib.gfk_student_t_proxy(_ptr(xd), n, d, _ptr(mean_dev), _ptr(prec_dev), ctypes.c_double(df),
                                           ctypes.c_double(cst), _ptr(log_q), _ptr(grad_q), _stream_ptr(dev)))
    return xd, log_q, grad_q


def t_log_pdf(x, mu, sigma, df, *, device=None):
    return _t_proxy_device(x, mu, sigma, df, device)[1].cpu().numpy()


def t_grad_log_pdf(x, mu, sigma, df, *, device=None):
    return _t_proxy_device(x, mu, sigma, df, device)[2].cpu().numpy()


def thin_gf_t(sample, log_p, t_mu, t_scale, t_df, thinned_size, *, precision=None, device=None):
    """``Gradient_free_Student_t.ipynb#c40``: Student-t proxy, then ``thin_gf(..., range_cap=200)``; the proxy's
    log-density and score never leave the device."""
    from .thinning import thin_gf
    xd, log_q, gradient_q = _t_proxy_device(sample, t_mu, t_scale, t_df, device)
    return thin_gf(xd, log_p, log_q, gradient_q, thinned_size, range_cap=200, precision=precision, device=device)


# ---- Gaussian KDE proxy ------------------------------------------------------------------------------------------
def kde_bandwidth(sample, weights=None, bw_method='silverman'):
    """The kernel covariance H of ``scipy.stats.gaussian_kde(sample.T, bw_method, weights)`` and the normalised
    weights: data covariance (weighted, unbiased) times factor**2 with Silverman's / Scott's factor on the
    effective sample size."""
    x = _host(sample)
    n, d = x.shape
    if weights is None:
        w = np.full(n, 1.0 / n)
    else:
        w = np.asarray(_host(weights), dtype=float)
        if w.shape != (n,):
            raise ValueError('weights must have one entry per row of sample')
        w = w / np.sum(w)
    neff = 1.0 / np.sum(w ** 2)
    if bw_method == 'silverman':
        factor = np.power(neff * (d + 2.0) / 4.0, -1. / (d + 4))
    elif bw_method == 'scott' or bw_method is None:
        factor = np.power(neff, -1. / (d + 4))
    elif np.isscalar(bw_method):
        factor = float(bw_method)
    else:
        raise ValueError("bw_method should be 'scott', 'silverman' or a scalar")
    cov = np.atleast_2d(np.cov(x, rowvar=False, bias=False, aweights=w))
    return cov * factor ** 2, w


def kde_logpdf_and_score(sample, weights=None, points=None, bw_method='silverman', *, device=None,
                         return_device=False):
    """``log q`` and ``grad log q`` of the Gaussian KDE of ``sample`` at ``points`` (default: the sample itself) -
    what ``logpdf_and_score(kde, sample)`` of ``Gaussian_mixture.ipynb#c46`` returns."""
    lib = _lib.load()
    dev = _require_cuda(device)
    h, w = kde_bandwidth(sample, weights, bw_method)
    d = h.shape[0]
    chol = np.linalg.cholesky(h)
    whiten = np.ascontiguousarray(np.linalg.inv(chol))                 # W' W = H^-1
    log_norm = -0.5 * (d * np.log(2 * np.pi) + 2.0 * np.sum(np.log(np.diag(chol))))
    with torch.cuda.device(dev):
        data = _to_device_f64(sample, dev, 'sample')
        pts = data if points is None else _to_device_f64(points, dev, 'points')
        n, m = data.shape[0], pts.shape[0]
        w_dev = None if weights is None else torch.from_numpy(np.ascontiguousarray(w)).to(dev)
        w_mat = torch.from_numpy(whiten).to(dev)
        log_q = torch.empty(m, dtype=torch.float64, device=dev)
        grad_q = torch.empty((m, d), dtype=torch.float64, device=dev)
        ws = _bytes(lib.gfk_kde_ws_bytes(n, m, d), dev)
        _lib.check(lib.gfk_kde_proxy(_ptr(data), _ptr(w_dev), n, _ptr(pts), m, d, _ptr(w_mat),
                                     ctypes.c_double(log_norm), _ptr(log_q), _ptr(grad_q), _ptr(ws), _stream_ptr(dev)))
    if return_device:
        return data, log_q, grad_q
    return log_q.cpu().numpy(), grad_q.cpu().numpy()


def kde_thin(sample, log_p, thinned_size, weights=None, bw_method='silverman', preconditioner='med', *,
             precision=None, device=None):
    """``Gaussian_mixture.ipynb#c47-#c48,#c51-#c54``: KDE proxy, then ``thin_gf(..., preconditioner='med')``."""
    from .thinning import thin_gf
    data, log_q, gradient_q = kde_logpdf_and_score(sample, weights, None, bw_method, device=device, return_device=True)
    return thin_gf(data, log_p, log_q, gradient_q, thinned_size, preconditioner=preconditioner, precision=precision,
                   device=device)


# ---- energy distance ---------------------------------------------------------------------------------------------
def cumulative_energy_distance(x, y, *, device=None):
    """SQUARED energy distance of ``x`` against ``y[:t + 1]`` for every t (float64 ndarray of length ``len(y)``)."""
    lib = _lib.load()
    dev = _require_cuda(device)
    with torch.cuda.device(dev):
        xd, yd = _to_device_f64(x, dev, 'x'), _to_device_f64(y, dev, 'y')
        if xd.dim() != 2 or yd.dim() != 2 or xd.shape[1] != yd.shape[1]:
            raise ValueError('x and y must be two-dimensional with the same number of columns')
        n, m, d = xd.shape[0], yd.shape[0], xd.shape[1]
        if n == 0 or m == 0:
            raise ValueError('x and y must not be empty')
        e2 = torch.empty(m, dtype=torch.float64, device=dev)
        ws = _bytes(lib.gfk_energy_ws_bytes(n, m), dev)
        _lib.check(lib.gfk_energy_distance_cumulative(_ptr(xd), n, _ptr(yd), m, d, _ptr(e2), _ptr(ws), _stream_ptr(dev)))
        return e2.cpu().numpy()


def energy_distance(x, y, *, device=None):
    """``dcor.energy_distance(x, y)`` (defaults): 2 E|X - Y| - E|X - X'| - E|Y - Y'|."""
    return float(cumulative_energy_distance(x, y, device=device)[-1])


def fit_quality(validation_sample, subsample, validation_sample_step=10, *, device=None):
    """``Comparison.ipynb#c19`` with the validation sample passed explicitly."""
    return float(np.sqrt(energy_distance(validation_sample[::validation_sample_step], subsample, device=device)))


def energy_distance_curve(validation_sample, sample, idx, thinned_sizes, validation_sample_step=10, *, device=None):
    """``rw_energy_distance`` of ``Comparison.ipynb#c21-#c22``: rows ``[thinned_size, fit_quality(sample[idx[:size]])]``
    for every requested size - from ONE cumulative pass instead of one energy-distance evaluation per size."""
    idx = np.asarray(idx).astype(np.int64)
    sizes = np.asarray(thinned_sizes, dtype=np.int64)
    if sizes.size and (sizes.min() < 1 or sizes.max() > idx.shape[0]):
        raise ValueError('thinned sizes must lie in [1, len(idx)]')
    if isinstance(sample, torch.Tensor):
        selected = sample[torch.as_tensor(idx, device=sample.device)]
    else:
        selected = np.asarray(sample)[idx]
    e2 = cumulative_energy_distance(validation_sample[::validation_sample_step], selected, device=device)
    return np.stack([sizes.astype(float), np.sqrt(np.maximum(e2[sizes - 1], 0.0))], axis=1)
