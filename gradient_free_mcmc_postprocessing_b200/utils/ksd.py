"""The reference's ``code/src/utils/ksd.py`` (``reindex_integrand``, ``calculate_ksd``) on the B200."""
import numpy as np

from ..integrand import SteinIntegrand
from ..stein import ksd
from ..thinning import _make_stein_integrand


def reindex_integrand(integrand, indices):
    """Integrand with index transformation (reference: code/src/utils/ksd.py:9-16)."""
    if isinstance(integrand, SteinIntegrand):
        return integrand.reindex(indices)
    indices = np.asarray(indices)

    def res(ind1, ind2):
        return integrand(indices[ind1], indices[ind2])
    return res


def calculate_ksd(sample, gradient, idx, *, precision=None, device=None):
    """Cumulative KSD for the provided indices (reference: code/src/utils/ksd.py:19-27)."""
    idx = np.asarray(idx)
    integrand = _make_stein_integrand(sample, gradient, precision=precision, device=device)
    integrand_restricted = reindex_integrand(integrand, idx)
    return ksd(integrand_restricted, idx.shape[0])
