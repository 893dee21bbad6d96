"""``stein_thinning.thinning`` + the reference's ``code/src/thinning.py`` on the B200.

Same names, positional order and defaults as the reference's call sites
(``thin``: Gradient_free_Stein_thinning.ipynb#c8, Gaussian_mixture.ipynb#c28;
``thin_gf``: code/src/thinning.py:17, Gaussian_mixture.ipynb#c39;
``_make_stein_integrand``: code/src/utils/ksd.py:25; ``_greedy_search``:
JAX_Stein_Thinning.ipynb#c18,#c22; ``gaussian_thin``: code/src/thinning.py:14-17).
Two keyword-only extensions: ``precision`` ('f64' strict parity mode [default], 'f64-fast',
'f32') and ``device``.
"""
import numpy as np
import torch

from ._engine import WEIGHT_SCALE_THRESHOLD, SteinProblem  # noqa: F401
from .integrand import SteinIntegrand, greedy_search_callable


def _validate_and_standardize(sample, gradient, standardize, *, device=None):
    """Returns the (standardised) sample and gradient as numpy arrays, computed on the device."""
    prob = SteinProblem(sample, gradient, standardize=standardize, preconditioner='id', precision='f64',
                        device=device)
    return prob.x.cpu().numpy(), prob.s.cpu().numpy()


def _make_stein_integrand(sample, gradient, standardize=True, preconditioner='id', *, precision=None,
                          device=None):
    return SteinIntegrand(SteinProblem(sample, gradient, standardize=standardize,
                                       preconditioner=preconditioner, precision=precision, device=device))


def _make_stein_gf_integrand(sample, log_p, log_q, gradient_q, standardize=True, preconditioner='id',
                             range_cap=None, *, precision=None, device=None):
    return SteinIntegrand(SteinProblem(sample, gradient_q, log_p=log_p, log_q=log_q, standardize=standardize,
                                       preconditioner=preconditioner, range_cap=range_cap,
                                       precision=precision, device=device))


def _greedy_search(n_points, integrand):
    """Greedy selection.  A ``SteinIntegrand`` runs fused on the GPU; any other callable obeying the
    integrand protocol is driven on the host exactly like the reference loop (its arithmetic is the
    callable's own, e.g. the fake integrand of code/tests/test_ksd.py)."""
    if isinstance(integrand, SteinIntegrand):
        return integrand.thin(n_points)
    return greedy_search_callable(n_points, integrand)


def thin(sample, gradient, n_points, standardize=True, preconditioner='id', *, precision=None, device=None):
    """Stein thinning: ``n_points`` greedily selected row indices (``np.ndarray[uint32]``, with replacement), exactly
    ``stein_thinning.thinning.thin``.

    Limits relative to the library the reference imports (neither is ever exercised by the reference, which fixes
    c = 1, beta = -1/2 and uses scalar preconditioners, report/report.tex:435): the IMQ exponent is fixed at
    beta = -1/2 (``kernel.vfk0_imq`` raises ``NotImplementedError`` for any other value); a dense preconditioner
    (``'smpcov'``) must be symmetric and runs in the ``'f64-fast'`` formulation even when ``precision='f64'`` is asked
    for (a ``UserWarning`` says so) - the strict, index-exact parity mode covers ``'id'``, ``'med'``, ``'sclmed'`` and
    numeric-string preconditioners."""
    integrand = _make_stein_integrand(sample, gradient, standardize, preconditioner, precision=precision,
                                      device=device)
    return _greedy_search(n_points, integrand)


def thin_gf(sample, log_p, log_q, gradient_q, n_points, standardize=True, preconditioner='id', range_cap=None,
            *, precision=None, device=None):
    """Gradient-free Stein thinning with proxy ``q`` (``stein_thinning.thinning.thin_gf``): the kernel is
    ``exp(lw_i + lw_j) k_Q(x_i, x_j)`` with ``lw = log_q - log_p``; warns like the reference when the spread of ``lw``
    exceeds 10.  Same limits as ``thin`` (beta = -1/2; dense preconditioners symmetric and in the FMA formulation)."""
    integrand = _make_stein_gf_integrand(sample, log_p, log_q, gradient_q, standardize, preconditioner,
                                         range_cap, precision=precision, device=device)
    return _greedy_search(n_points, integrand)


def gaussian_proxy_device(sample, mean, cov, *, device=None):
    """The Gaussian proxy's log-density and score on the device: returns CUDA
    tensors ``(sample_dev, log_q, gradient_q)``.  The d x d inverse and log-determinant stay on the host."""
    import ctypes
    from . import _lib
    from ._engine import _ptr, _require_cuda, _stream_ptr, _to_device_f64
    lib = _lib.load()
    dev = _require_cuda(device)
    cov = np.atleast_2d(np.asarray(cov, dtype=float))
    mean = np.atleast_1d(np.asarray(mean, dtype=float))
    d = mean.shape[0]
    if cov.shape != (d, d) or tuple(sample.shape)[1:] != (d,):
        raise ValueError('inconsistent shapes of sample, mean and cov')
    sign, logdet = np.linalg.slogdet(cov)
    if sign <= 0:
        raise np.linalg.LinAlgError('cov must be symmetric positive definite')
    cst = d * np.log(2 * np.pi) + logdet
    with torch.cuda.device(dev):
        x = _to_device_f64(sample, dev, 'sample')
        n = x.shape[0]
        mean_dev = torch.from_numpy(mean).to(dev)
        prec_dev = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(cov))).to(dev)
        log_q = torch.empty(n, dtype=torch.float64, device=dev)
        grad_q = torch.empty((n, d), dtype=torch.float64, device=dev)
        _lib.check(lib.gfk_gaussian_proxy(_ptr(x), n, d, _ptr(mean_dev), _ptr(prec_dev), ctypes.c_double(cst),
                                          _ptr(log_q), _ptr(grad_q), _stream_ptr(dev)))
    return x, log_q, grad_q


def gaussian_thin(sample, log_p, mean, cov, thinned_size, range_cap=200, *, precision=None, device=None,
                  proxy_on_device=False):
    """Gaussian-proxy gradient-free thinning (reference: code/src/thinning.py:14-17).

    By default the proxy's log-density and score are computed on the host exactly as in the reference (scipy
    ``mvn.logpdf`` + einsum).  ``proxy_on_device=True`` computes them with ``gfk_gaussian_proxy`` instead, which
    also removes the host-to-device copy of ``log_q`` / ``gradient_q`` (added in
    round 1 without GPU access, so it is opt-in until it has been measured).
    """
    if proxy_on_device:
        x_dev, log_q, gradient_q = gaussian_proxy_device(sample, mean, cov, device=device)
        return thin_gf(x_dev, log_p, log_q, gradient_q, thinned_size, range_cap=range_cap, preconditioner='med',
                       precision=precision, device=device)
    from scipy.stats import multivariate_normal as mvn
    sample_np = sample.cpu().numpy() if isinstance(sample, torch.Tensor) else np.asarray(sample)
    log_q = mvn.logpdf(sample_np, mean=mean, cov=cov)
    gradient_q = -np.einsum('ij,kj->ki', np.linalg.inv(cov), sample_np - mean)
    return thin_gf(sample, log_p, log_q, gradient_q, thinned_size, range_cap=range_cap, preconditioner='med',
                   precision=precision, device=device)
