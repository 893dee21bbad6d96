"""``stein_thinning.kernel`` on the B200: ``vfk0_imq``, ``make_precon``, ``make_imq``.

Reference call sites: code/notebooks/Kernel_Stein_discrepancy.ipynb#c2,
code/notebooks/examples/JAX_Stein_Thinning.ipynb#c15-#c16,#c26-#c28.
"""
import numpy as np

from ._engine import SteinProblem


def make_precon(sample, preconditioner='id', *, device=None):
    """Inverse preconditioner ``linv`` (d x d ndarray) of the (already standardised) sample."""
    sample = np.asarray(sample, dtype=float)
    prob = SteinProblem(sample, np.zeros_like(sample), standardize=False, preconditioner=preconditioner,
                        precision='f64', device=device)
    return prob.linv


def vfk0_imq(x, y, sx, sy, linv, c=1.0, beta=-0.5, *, precision=None, device=None):
    """Row-paired Stein kernel; ``y``/``sy`` may be one row (broadcast).  ``linv`` must be
    ``lambda * I`` (any lambda > 0) or symmetric; ``beta`` is fixed at -1/2 as in the reference."""
    if beta != -0.5:
        raise NotImplementedError('the CUDA kernels implement beta = -1/2 (report/report.tex:435)')
    x, y, sx, sy = (np.atleast_2d(np.asarray(a, dtype=float)) for a in (x, y, sx, sy))
    linv = np.atleast_2d(np.asarray(linv, dtype=float))
    n, d = x.shape
    if y.shape[0] not in (1, n) or sy.shape != y.shape or sx.shape != x.shape or y.shape[1] != d:
        raise ValueError('inconsistent shapes')
    prob = SteinProblem(np.concatenate([x, y]), np.concatenate([sx, sy]), standardize=False,
                        preconditioner=linv, precision=precision, device=device, c=c)
    ia = np.arange(n)
    ib = n + (np.zeros(n, dtype=np.int64) if y.shape[0] == 1 else np.arange(n))
    return prob.pairs(ia, ib)


def make_imq(sample, preconditioner='id', *, precision=None, device=None):
    """Closure ``vfk0(x, y, sx, sy)`` with the preconditioner of ``sample`` baked in."""
    linv = make_precon(sample, preconditioner, device=device)

    def vfk0(x, y, sx, sy):
        return vfk0_imq(x, y, sx, sy, linv, precision=precision, device=device)
    return vfk0
