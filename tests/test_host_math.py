"""The exact arithmetic the CUDA kernels run (csrc/stein_math.cuh, compiled for the host by
tests/host/build.sh) against numpy, on the CPU.  Shows that the strict formulation reproduces
numpy's operation order bit for bit (everything except numpy's own pow rounding)."""
import ctypes
import decimal
import os
import subprocess

import numpy as np
import pytest

from oracle.stein_thinning_oracle.kernel import vfk0_imq

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'host', 'libpair_check.so')
P = ctypes.POINTER(ctypes.c_double)


SO_PERTURBED = os.path.join(HERE, 'host', 'libpair_check_perturbed.so')


@pytest.fixture(scope='module', params=['exact-rsqrt', 'rsqrt-off-by-up-to-4ulp'])
def hostlib(request):
    """The second build perturbs the approximate reciprocal square root by -3..+4 ulp, as a device
    approximation may be off: every result below must be unchanged (the corrections absorb it)."""
    if not (os.path.exists(SO) and os.path.exists(SO_PERTURBED)):
        subprocess.run(['bash', os.path.join(HERE, 'host', 'build.sh')], check=True)
    return ctypes.CDLL(SO if request.param == 'exact-rsqrt' else SO_PERTURBED)


def _p(a):
    return a.ctypes.data_as(P)


def _call(fn, x, y, sx, sy, lam, lam2, tr, c=1.0):
    out = np.empty(x.shape[0])
    fn(_p(x), _p(y), _p(sx), _p(sy), ctypes.c_int64(x.shape[0]), ctypes.c_int64(y.shape[0]),
       ctypes.c_int(x.shape[1]), ctypes.c_double(lam), ctypes.c_double(lam2), ctypes.c_double(tr),
       ctypes.c_double(c), _p(out))
    return out


def _cr_pow(q, half_odd):
    """Correctly rounded q**half_odd (half_odd in 1.5, 2.5) by 70-digit decimal arithmetic."""
    D = decimal.Decimal
    with decimal.localcontext() as ctx:
        ctx.prec = 70
        return np.array([float(D(float(v)) ** int(half_odd - 0.5) * D(float(v)).sqrt()) for v in q])


def _vfk0_cr(x, y, sx, sy, linv):
    """vfk0_imq with numpy's pow replaced by the correctly rounded value; same order otherwise."""
    c, beta = 1.0, -0.5
    diff = x.T - y.T
    quad = c + np.sum(np.dot(linv, diff) * diff, axis=0)
    first = -4 * beta * (beta - 1) * np.sum(np.dot(np.dot(linv, linv), diff) * diff, axis=0) / _cr_pow(quad, 2.5)
    second = -2 * beta * (np.trace(linv) + np.sum(np.dot(linv, sx.T - sy.T) * diff, axis=0)) / _cr_pow(quad, 1.5)
    third = np.sum(sx.T * sy.T, axis=0) / (quad ** (-beta))
    return first + second + third


def test_strict_powers_correctly_rounded(hostlib):
    rng = np.random.default_rng(3)
    q = 1 + rng.random(20000) * np.exp(rng.normal(0, 3, 20000))
    p05, p15, p25 = np.empty_like(q), np.empty_like(q), np.empty_like(q)
    hostlib.host_strict_powers(_p(q), ctypes.c_int64(q.size), _p(p05), _p(p15), _p(p25))
    assert np.array_equal(p05, np.sqrt(q))
    assert np.array_equal(p15, _cr_pow(q, 1.5))
    assert np.array_equal(p25, _cr_pow(q, 2.5))


@pytest.mark.parametrize('d', [1, 2, 3, 4, 7, 8, 10, 16, 64, 130, 300])
@pytest.mark.parametrize('broadcast', [True, False])
def test_strict_pair_is_numpy_order(hostlib, d, broadcast):
    rng = np.random.default_rng(100 + d)
    n = 400
    x, sx = rng.standard_normal((n, d)), rng.standard_normal((n, d)) * 3
    y, sy = (x[[17]].copy(), sx[[17]].copy()) if broadcast else (x[::-1].copy(), sx[::-1].copy())
    linv = np.linalg.inv(1.7234567 * np.identity(d))
    lam, lam2, tr = linv[0, 0], np.dot(linv, linv)[0, 0], np.trace(linv)
    got = _call(hostlib.host_pair_strict, x, y, sx, sy, lam, lam2, tr)
    assert np.array_equal(got, _vfk0_cr(x, y, sx, sy, linv))       # bit for bit
    ref = vfk0_imq(x, y, sx, sy, linv)                             # numpy's own pow: ~95 % correctly rounded
    assert np.mean(got == ref) > 0.9
    fast = _call(hostlib.host_pair_fast, x, y, sx, sy, lam, lam2, tr)
    scale = np.abs(ref) + 1e-3 * np.max(np.abs(ref))
    assert np.max(np.abs(fast - ref) / scale) < 1e-11


def test_markstein_division_is_ieee(hostlib):
    # t3 = s4 / sqrt(qf) with random numerators: strict quotients must equal IEEE division
    rng = np.random.default_rng(77)
    n = 20000
    x = rng.standard_normal((n, 1)) * np.exp(rng.normal(0, 2, (n, 1)))
    y = np.zeros((1, 1))
    sx = rng.standard_normal((n, 1)) * np.exp(rng.normal(0, 4, (n, 1)))
    sy = np.ones((1, 1))
    got = _call(hostlib.host_pair_strict, x, y, sx, sy, 1.0, 1.0, 0.0)
    # direct: with trace = 0 the kernel is (-3 r^2 / q^2.5 + sx*r / q^1.5) + sx / q^0.5 in numpy's order
    r = x[:, 0] - 0.0
    q = 1.0 + (1.0 * r) * r
    t1 = (-3.0 * ((1.0 * r) * r)) / _cr_pow(q, 2.5)
    t2 = (0.0 + (1.0 * (sx[:, 0] - 1.0)) * r) / _cr_pow(q, 1.5)
    t3 = (sx[:, 0] * 1.0) / np.sqrt(q)
    assert np.array_equal(got, (t1 + t2) + t3)


def test_pairwise_dot_is_numpy_pairwise(hostlib):
    rng = np.random.default_rng(9)
    for d in (1, 5, 8, 9, 17, 64, 128, 129, 200, 257, 1000):
        a, b = rng.standard_normal((50, d)), rng.standard_normal((50, d))
        out = np.empty(50)
        hostlib.host_pairwise_dot(_p(a), _p(b), ctypes.c_int64(50), ctypes.c_int(d), _p(out))
        assert np.array_equal(out, np.sum(a.T * b.T, axis=0)), d
