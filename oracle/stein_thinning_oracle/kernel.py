"""Oracle restatement of ``stein_thinning.kernel`` (TEST INFRASTRUCTURE ONLY).

Sources followed (paths under ``/root/reference``; ``#cN`` = 0-based notebook
cell):

* ``vfk0_imq``  - ``code/notebooks/examples/JAX_Stein_Thinning.ipynb#c27``
  (the library's expression, restated there with ``jnp`` in place of ``np``)
  and ``code/notebooks/Kernel_Stein_discrepancy.ipynb#c7`` (term split);
  derivation ``report/report.tex:853-868``; ``c = 1``, ``beta = -1/2`` at
  ``report/report.tex:435``.
* ``make_precon`` / ``make_imq`` - call sites
  ``JAX_Stein_Thinning.ipynb#c16,#c26,#c28``; ``'med'`` is the squared median
  pairwise Euclidean distance (``report/report.tex:323,435``).

The ORDER of floating-point operations below is deliberate: it is the order
numpy evaluates the library's expression in, which is what fixes the last bit
of every kernel value ("hard parts").
"""
import numpy as np
from scipy.spatial.distance import pdist

# Upper bound on the number of rows that enter the median heuristic
# (PARITY UNPINNED for n > MEDIAN_SUBSAMPLE).
MEDIAN_SUBSAMPLE = 1000


def vfk0_imq(x, y, sx, sy, linv, c=1.0, beta=-0.5):
    """Row-paired Langevin Stein kernel built on the IMQ base kernel.

    ``x, sx`` are ``(n, d)``; ``y, sy`` are ``(n, d)`` or ``(1, d)`` (broadcast);
    ``linv`` is the ``(d, d)`` inverse preconditioner.  Returns ``(n,)``.
    """
    diff = x.T - y.T                                   # (d, n)
    quad = c + np.sum(np.dot(linv, diff) * diff, axis=0)
    first = (-4 * beta * (beta - 1)
             * np.sum(np.dot(np.dot(linv, linv), diff) * diff, axis=0)
             / (quad ** (-beta + 2)))
    second = (-2 * beta
              * (np.trace(linv) + np.sum(np.dot(linv, sx.T - sy.T) * diff, axis=0))
              / (quad ** (-beta + 1)))
    third = np.sum(sx.T * sy.T, axis=0) / (quad ** (-beta))
    return first + second + third


def _is_number(text):
    try:
        float(text)
    except (TypeError, ValueError):
        return False
    return True


def _median_sq(sample, cap=MEDIAN_SUBSAMPLE):
    """Squared median of pairwise Euclidean distances on <= ``cap`` rows.

    For n > cap the rows are ``sample[np.linspace(0, n - 1, cap, dtype=int)]``
    (SURVEY.md section 8a row a2 - upstream behaviour, PARITY UNPINNED).
    """
    n = sample.shape[0]
    rows = sample[np.linspace(0, n - 1, cap, dtype=int)] if n > cap else sample
    return np.median(pdist(rows)) ** 2


def make_precon(sample, preconditioner='id'):
    """Inverse preconditioner matrix ``linv`` (d x d) for the IMQ kernel."""
    n, d = sample.shape
    if preconditioner == 'id':
        return np.identity(d)
    if preconditioner in ('med', 'sclmed'):
        m2 = _median_sq(sample)
        if m2 == 0:
            raise Exception('Too few unique samples in sample.')
        if preconditioner == 'sclmed':
            m2 = m2 / np.log(np.minimum(MEDIAN_SUBSAMPLE, n))
        return np.linalg.inv(m2 * np.identity(d))
    if preconditioner == 'smpcov':
        cov = np.cov(sample, rowvar=False)
        if not np.all(np.linalg.eigvals(np.atleast_2d(cov)) > 0):
            raise Exception('Too few unique samples in sample.')
        return np.linalg.inv(np.atleast_2d(cov))
    if _is_number(preconditioner):
        return np.linalg.inv(float(preconditioner) * np.identity(d))
    raise ValueError('Incorrect preconditioner type.')


def make_imq(sample, preconditioner='id'):
    """Closure ``vfk0(x, y, sx, sy)`` with the preconditioner baked in."""
    linv = make_precon(sample, preconditioner)

    def vfk0(x, y, sx, sy):
        return vfk0_imq(x, y, sx, sy, linv)
    return vfk0
