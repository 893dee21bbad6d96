"""Oracle restatement of the reference repo's own wrappers (TEST INFRASTRUCTURE ONLY).

* ``reindex_integrand`` / ``calculate_ksd`` - ``/root/reference/code/src/utils/ksd.py:9-27``
* ``gaussian_thin``                       - ``/root/reference/code/src/thinning.py:14-17``

``laplace_approximation`` (``code/src/thinning.py:8-11``, scipy BFGS on the host)
is out of scope.
"""
import numpy as np
from scipy.stats import multivariate_normal as mvn

from .stein_thinning_oracle.stein import ksd
from .stein_thinning_oracle.thinning import _make_stein_integrand, thin_gf


def reindex_integrand(integrand, indices):
    def reindexed(ind1, ind2):
        return integrand(indices[ind1], indices[ind2])
    return reindexed


def calculate_ksd(sample, gradient, idx):
    full = _make_stein_integrand(sample, gradient)
    return ksd(reindex_integrand(full, idx), idx.shape[0])


def gaussian_thin(sample, log_p, mean, cov, thinned_size, range_cap=200):
    log_q = mvn.logpdf(sample, mean=mean, cov=cov)
    gradient_q = -np.einsum('ij,kj->ki', np.linalg.inv(cov), sample - mean)
    return thin_gf(sample, log_p, log_q, gradient_q, thinned_size,
                   range_cap=range_cap, preconditioner='med')
