"""Oracle restatement of ``stein_thinning.stein`` (TEST INFRASTRUCTURE ONLY).

Sources followed (paths under ``/root/reference``):

* ``kmat(integrand, n)`` - index-calling convention pinned by
  ``code/tests/test_ksd.py:8-28`` (the integrand receives two equal-length
  index arrays and the result is the dense symmetric matrix).
* ``ksd(integrand, n)`` - cumulative KSD, ``ks[t] = sqrt(sum_{a,b<=t} k(a,b)) / (t+1)``:
  docstring at ``code/src/utils/ksd.py:24``, callers
  ``code/notebooks/lotka_volterra/Comparison.ipynb#c28,#c29``, definition
  ``report/report.tex:309-313``.  Row t adds ``2*sum_{b<t} k(b,t) + k(t,t)`` to a
  running sum.  Numeric values of ``ksd`` are
  PARITY UNPINNED (only figures exist in the reference).
"""
import numpy as np


def kmat(integrand, n):
    rows, cols = np.triu_indices(n)
    upper = integrand(rows, cols)
    out = np.zeros((n, n), dtype=np.asarray(upper).dtype)
    out[rows, cols] = upper
    out[cols, rows] = upper
    return out


def ksd(integrand, n):
    ks = np.empty(n)
    total = 0.0
    for t in range(n):
        diag = np.asarray(integrand([t], [t])).reshape(-1)[0]
        if t > 0:
            off = np.sum(integrand(np.arange(t), [t]))
            total += 2 * off + diag
        else:
            total += diag
        ks[t] = np.sqrt(total) / (t + 1)
    return ks
