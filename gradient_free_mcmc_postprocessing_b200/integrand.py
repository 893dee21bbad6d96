"""The integrand protocol of ``stein_thinning`` (``integrand(ind1, ind2) -> ndarray``), device backed.

Reference: the closures returned by ``_make_stein_integrand`` / ``_make_stein_gf_integrand``
(code/notebooks/examples/JAX_Stein_Thinning.ipynb#c17; code/src/utils/ksd.py:9-16;
code/tests/test_ksd.py:16-20).  Index arguments may be ``slice(None)``, lists or arrays and
broadcast against each other like numpy fancy indexing of the paired rows.
"""
import numpy as np


def _expand(index, n):
    if isinstance(index, slice):
        return np.arange(n)[index]
    return np.asarray(index, dtype=np.int64)


class SteinIntegrand:
    """Callable integrand whose pair evaluations, greedy search, KSD and Gram matrix run in libgfk."""

    def __init__(self, problem, indices=None):
        self.problem = problem
        self.indices = None if indices is None else np.asarray(indices, dtype=np.int64)

    @property
    def n(self):
        return self.problem.n if self.indices is None else len(self.indices)

    def __call__(self, ind1, ind2):
        a, b = np.broadcast_arrays(_expand(ind1, self.n), _expand(ind2, self.n))
        if self.indices is not None:
            a, b = self.indices[a], self.indices[b]
        return self.problem.pairs(a, b)

    def reindex(self, indices):
        indices = np.asarray(indices, dtype=np.int64)
        return SteinIntegrand(self.problem, indices if self.indices is None else self.indices[indices])

    def thin(self, n_points):
        if self.indices is not None:
            raise NotImplementedError('greedy search over a re-indexed integrand is not part of the reference path')
        return self.problem.thin(n_points)

    def ksd(self, n):
        if self.indices is None:
            return self.problem.ksd(None, n)
        return self.problem.ksd(self.indices[:n], n)

    def kmat(self, n):
        if self.indices is None:
            return self.problem.kmat(None, n)
        return self.problem.kmat(self.indices[:n], n)


# ---- foreign callables: the protocol itself, driven on the host ---------------------------------------
# These loops contain no kernel arithmetic: every value comes from the caller's own integrand.

def greedy_search_callable(n_points, integrand):
    idx = np.empty(n_points, dtype=np.uint32)
    running = np.array(integrand(slice(None), slice(None)), dtype=float)
    idx[0] = np.argmin(running)
    for i in range(1, n_points):
        running += 2 * np.asarray(integrand(slice(None), [idx[i - 1]]))
        idx[i] = np.argmin(running)
    return idx


def kmat_callable(integrand, n):
    rows, cols = np.triu_indices(n)
    upper = np.asarray(integrand(rows, cols))
    out = np.zeros((n, n), dtype=upper.dtype)
    out[rows, cols] = upper
    out[cols, rows] = upper
    return out


def _ksd_rows(integrand, n):
    """Row by row, the calling pattern of the reference loop (one integrand call per row)."""
    low, diag = np.zeros(n), np.empty(n)
    for t in range(n):
        diag[t] = np.asarray(integrand([t], [t])).reshape(-1)[0]
        if t > 0:
            low[t] = np.sum(integrand(np.arange(t), [t]))
    return low, diag


def _ksd_batched(integrand, n, max_pairs=1 << 22):
    """The same pair values from a few large element-wise calls (whole blocks of rows of the lower triangle per call)
    - what makes the reference's own ``calculate_ksd`` closure (code/src/utils/ksd.py:9-27) cost ~m^2 / 4M device
    calls instead of m."""
    idx = np.arange(n)
    diag = np.asarray(integrand(idx, idx), dtype=float).reshape(-1)
    if diag.shape != (n,):
        raise ValueError('integrand does not evaluate element-wise over index arrays')
    low = np.zeros(n)
    t0 = 1
    while t0 < n:
        t1 = t0 + 1                                      # rows [t0, t1): t0 + ... + (t1 - 1) pairs
        while t1 < n and (t1 * (t1 + 1) - t0 * (t0 - 1)) // 2 <= max_pairs:
            t1 += 1
        rows = np.repeat(np.arange(t0, t1), np.arange(t0, t1))
        starts = np.cumsum(np.arange(t0, t1)) - np.arange(t0, t1)
        cols = np.arange(rows.size) - np.repeat(starts, np.arange(t0, t1))
        vals = np.asarray(integrand(rows, cols), dtype=float).reshape(-1)
        if vals.shape != rows.shape:
            raise ValueError('integrand does not evaluate element-wise over index arrays')
        low[t0:t1] = np.add.reduceat(vals, starts)
        t0 = t1
    return low, diag


def _evaluates_elementwise(integrand, n):
    """Does the foreign callable accept two equal-length index arrays and pair them element-wise?  Decided by ONE
    two-pair probe against the (array, [t]) form every integrand supports; the answer only selects the calling
    pattern, so whatever the probe raises is not an error yet - a genuine defect of the callable resurfaces,
    unswallowed, in the pattern that is then used."""
    if n < 2:
        return False
    try:
        both = np.asarray(integrand(np.array([0, 1]), np.array([1, 0])), dtype=float)
        one = np.asarray(integrand(np.array([0]), [1]), dtype=float)
    except Exception:   # noqa: BLE001 - a probe; see the docstring
        return False
    return both.shape == (2,) and one.reshape(-1).shape == (1,) and \
        (both[0] == one.reshape(-1)[0] or (np.isnan(both[0]) and np.isnan(one.reshape(-1)[0])))


def ksd_callable(integrand, n):
    if _evaluates_elementwise(integrand, n):
        low, diag = _ksd_batched(integrand, n)           # errors propagate
    else:
        low, diag = _ksd_rows(integrand, n)              # a callable that only supports the (array, [t]) form
    ks = np.empty(n)
    total = 0.0
    for t in range(n):
        total += (2 * low[t] + diag[t]) if t > 0 else diag[t]
        ks[t] = np.sqrt(total) / (t + 1)
    return ks
