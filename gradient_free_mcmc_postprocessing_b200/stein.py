"""``stein_thinning.stein`` on the B200: ``ksd(integrand, n)`` and ``kmat(integrand, n)``.

Reference call sites: code/src/utils/ksd.py:27, code/tests/test_ksd.py:20,
code/notebooks/gaussian_mixture/Gaussian_mixture.ipynb#c94.  A ``SteinIntegrand`` is evaluated by
the tiled CUDA kernels; any other callable is driven through the protocol on the host.
"""
from .integrand import SteinIntegrand, kmat_callable, ksd_callable


def ksd(integrand, n):
    """Cumulative KSD: ``ks[t] = sqrt(sum_{a,b<=t} k(a,b)) / (t+1)``, length ``n``."""
    if isinstance(integrand, SteinIntegrand):
        return integrand.ksd(int(n))
    return ksd_callable(integrand, int(n))


def kmat(integrand, n):
    """Dense symmetric ``n x n`` Gram matrix of the integrand."""
    if isinstance(integrand, SteinIntegrand):
        return integrand.kmat(int(n))
    return kmat_callable(integrand, int(n))
