"""The REAL kernel sources of libgfk (csrc/*.cu) executed on the CPU by the SIMT emulator in tests/host/simt and
compared with the oracle - the same parity cases as tests/test_gpu_parity.py at sizes a CPU finishes in seconds.

This is test infrastructure: it proves the kernels' logic (index arithmetic, shared-memory layouts, barriers, warp /
block / grid reductions, tie-breaking, the flag-in-data grid exchange of the single-launch kernel) without a GPU.  It
says nothing about device arithmetic that differs from the host's (CUDA's exp / rsqrt), which the `-m gpu` tests
cover.  With the host's exact reciprocal square root the strict mode must reproduce the oracle's indices exactly.
"""
import ctypes
import os
import warnings

import numpy as np
import pytest

import simt_harness as H
from oracle.stein_thinning_oracle import kernel as ok, stein as os_, thinning as ot


def _mcmc_like(rng, n, d, dup=0.5):
    base = rng.standard_normal((n, d)) @ (np.eye(d) + 0.3 * rng.standard_normal((d, d)))
    keep = rng.random(n) > dup
    keep[0] = True
    src = np.maximum.accumulate(np.where(keep, np.arange(n), 0))
    x = base[src]
    grad = -(x - 0.1) * (1.0 + 0.2 * np.arange(d))
    return np.ascontiguousarray(x), np.ascontiguousarray(grad)


@pytest.fixture(autouse=True)
def _clean_env(monkeypatch):
    for k in ('GFK_PERSISTENT', 'GFK_RESIDENT', 'SIMT_SMS', 'SIMT_WORKERS'):
        monkeypatch.delenv(k, raising=False)
    H.load()
    H.set_schedule(False)
    yield
    H.set_schedule(False)


STRATEGIES = {
    'single-launch': {},
    'single-launch-rows-from-global': {'GFK_RESIDENT': '0'},
    'launch-per-step': {'GFK_PERSISTENT': '0'},
}


# ---- golden vectors through the emulated kernels -----------------------------------------------------------
@pytest.mark.parametrize('strategy', list(STRATEGIES))
@pytest.mark.parametrize('precision', ['f64', 'f64-fast'])
def test_golden_g1_g2_g3(g13, strategy, precision, monkeypatch):
    for k, v in STRATEGIES[strategy].items():
        monkeypatch.setenv(k, v)
    idx, _ = H.SimProblem(g13['sample'], g13['gradient'], precision=precision).thin(20)
    assert np.array_equal(idx, g13['idx_thin'])
    idx, _ = H.SimProblem(g13['sample'], g13['gradient'], g13['log_p'], g13['log_p'], precision=precision).thin(20)
    assert np.array_equal(idx, g13['idx_thin'])
    idx, _ = H.SimProblem(g13['sample'], g13['gradient_q'], g13['log_p'], g13['log_q'], precision=precision).thin(20)
    assert np.array_equal(idx, g13['idx_thin_gf'])


def test_golden_g4_g5(g45):
    idx, _ = H.SimProblem(g45['sample'], g45['gradient'], preconditioner='med').thin(1000)
    assert len(np.unique(idx)) == 538
    idx, _ = H.SimProblem(g45['sample'], g45['gradient_q'], g45['log_p'], g45['log_q'], preconditioner='med').thin(1000)
    assert len(np.unique(idx)) == 400
    p = H.SimProblem(g45['sample'], g45['gradient_q_laplace'], g45['log_p'], g45['log_q_laplace'], preconditioner='med')
    assert p.spread > 10                                     # the caller's UserWarning condition
    idx, _ = p.thin(1500)
    assert np.array_equal(np.unique(idx), [494])


# ---- greedy search: every kernel variant against the oracle ----------------------------------------------------
@pytest.mark.parametrize('d', [1, 2, 3, 4])
@pytest.mark.parametrize('sms', [1, 3, 7])
def test_single_launch_kernel_matches_oracle(d, sms, monkeypatch):
    """thin_persistent_kernel: 1, 3 and 7 concurrently running CTAs, partly filled last CTA, rows per thread > 1."""
    monkeypatch.setenv('SIMT_SMS', str(sms))
    rng = np.random.default_rng(100 + d)
    n = 2500 + 37 * d
    x, g = _mcmc_like(rng, n, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1) + 0.3 * np.sin(x[:, 0])
    log_q = -0.5 * np.sum((x - 0.1) ** 2, axis=1) / 1.3
    m = 40
    expect = ot.thin(x, g, m, preconditioner='med')
    expect_gf = ot.thin_gf(x, log_p, log_q, g, m, preconditioner='med')
    for precision in ('f64', 'f64-fast'):
        idx, _ = H.SimProblem(x, g, preconditioner='med', precision=precision).thin(m)
        assert np.array_equal(idx, expect), (precision, np.flatnonzero(idx != expect)[:3])
        idx, _ = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=precision).thin(m)
        assert np.array_equal(idx, expect_gf), (precision, np.flatnonzero(idx != expect_gf)[:3])


@pytest.mark.parametrize('d', [1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 33, 64, 130])
def test_launch_per_step_kernels_match_oracle(d, monkeypatch):
    """thin_step_small_kernel (d <= 4), thin_step_chunked_kernel (d <= 128), thin_step_tiled_kernel (strict d > 128),
    thin_init_kernel, the last-block-done grid argmin and thin_finish_kernel."""
    monkeypatch.setenv('GFK_PERSISTENT', '0')
    rng = np.random.default_rng(1000 + d)
    x, g = _mcmc_like(rng, 1300, d)
    m = 25
    for pre in ('id', 'med'):
        expect = ot.thin(x, g, m, preconditioner=pre)
        for precision in ('f64', 'f64-fast'):
            idx, _ = H.SimProblem(x, g, preconditioner=pre, precision=precision).thin(m)
            assert np.array_equal(idx, expect), (precision, d, pre, np.flatnonzero(idx != expect)[:3])


@pytest.mark.parametrize('range_cap', [None, 5.0])
@pytest.mark.parametrize('d', [2, 10])
def test_gradient_free_with_range_cap(d, range_cap):
    rng = np.random.default_rng(2000 + d)
    x, gq = _mcmc_like(rng, 1500, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1) + 0.3 * np.sin(x[:, 0])
    log_q = -0.5 * np.sum((x - 0.1) ** 2, axis=1) / 1.3
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        expect = ot.thin_gf(x, log_p, log_q, gq, 30, preconditioner='med', range_cap=range_cap)
    for precision in ('f64', 'f64-fast'):
        idx, _ = H.SimProblem(x, gq, log_p, log_q, preconditioner='med', range_cap=range_cap, precision=precision).thin(30)
        assert np.array_equal(idx, expect), (precision, d, range_cap)


@pytest.mark.parametrize('d', [2, 3, 4])
def test_launch_strategies_agree_bitwise(d, monkeypatch):
    """Same indices AND the same final running objective, bit for bit, from all three launch strategies."""
    rng = np.random.default_rng(500 + d)
    x, g = _mcmc_like(rng, 3001, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = log_p / 1.1
    for precision in ('f64', 'f64-fast', 'f32'):
        results = []
        for env in STRATEGIES.values():
            for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            results.append(H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=precision).thin(60))
        for idx, obj in results[1:]:
            assert np.array_equal(idx, results[0][0]), precision
            assert np.array_equal(obj, results[0][1]), precision


def test_final_objective_is_numpys_bit_for_bit():
    """Without standardisation (its column sums are reduced in a different order) the strict kernels must leave
    exactly the running objective numpy computes, wherever numpy's pow is correctly rounded - demand 90 % equal
    and the rest within 4 ulp-ish."""
    rng = np.random.default_rng(9)
    x, g = _mcmc_like(rng, 2000, 3, dup=0.3)
    integrand = ot._make_stein_integrand(x, g, standardize=False)
    m = 15
    expect = ot._greedy_search(m, integrand)
    a = integrand(slice(None), slice(None)).copy()
    for j in expect[:-1]:
        a += 2 * integrand(slice(None), [j])
    idx, obj = H.SimProblem(x, g, standardize=False).thin(m)
    assert np.array_equal(idx, expect)
    assert np.mean(obj == a) > 0.9
    np.testing.assert_allclose(obj, a, rtol=1e-14)


def test_selection_with_replacement_and_row_shards(monkeypatch):
    rng = np.random.default_rng(5)
    x, g = _mcmc_like(rng, 40, 2, dup=0.0)
    expect = ot.thin(x, g, 300)
    for env in STRATEGIES.values():
        for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        idx, _ = H.SimProblem(x, g).thin(300)
        assert np.array_equal(idx, expect)
    # a rank's shard [r0, r1) alone: greedy search restricted to those candidate rows (global indices)
    x, g = _mcmc_like(rng, 900, 3)
    p = H.SimProblem(x, g, standardize=False)
    integrand = ot._make_stein_integrand(x, g, standardize=False)
    r0, r1 = 311, 777
    a = integrand(slice(None), slice(None))[r0:r1].copy()
    want = []
    for _ in range(12):
        j = r0 + int(np.argmin(a))
        want.append(j)
        a += 2 * integrand(np.arange(r0, r1), [j])
    idx, _ = p.thin(12, row_range=(r0, r1))
    assert np.array_equal(idx, want)


def test_nan_objective_is_selected_first_like_numpy(monkeypatch):
    """np.argmin returns the first NaN; an inf - inf in the weights produces one."""
    rng = np.random.default_rng(3)
    x, g = _mcmc_like(rng, 600, 2, dup=0.0)
    g[417] = 1e200      # overflow to inf in <sx, sy>, inf - inf -> NaN downstream
    g[88] = 1e200
    with np.errstate(all='ignore'):
        expect = ot.thin(x, g, 5, standardize=False)
    for env in ({}, {'GFK_PERSISTENT': '0'}):
        monkeypatch.delenv('GFK_PERSISTENT', raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        idx, _ = H.SimProblem(x, g, standardize=False).thin(5)
        assert np.array_equal(idx, expect)


def test_infinite_objectives_like_numpy(monkeypatch):
    """All objectives +inf: np.argmin returns index 0 every step; some +inf: they never win."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((700, 2))
    g_all = np.full((700, 2), 1e200)
    g_some = g_all.copy()
    g_some[300:] = rng.standard_normal((400, 2))
    for g in (g_all, g_some):
        with np.errstate(all='ignore'):
            expect = ot.thin(x, g, 4, standardize=False)
        for env in STRATEGIES.values():
            for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            for precision in ('f64', 'f64-fast'):
                assert np.array_equal(H.SimProblem(x, g, standardize=False, precision=precision).thin(4)[0], expect)


# ---- kernel values, Gram matrix, KSD ---------------------------------------------------------------------------
@pytest.mark.parametrize('gf', [False, True])
@pytest.mark.parametrize('d', [1, 2, 7, 10, 33, 70])
def test_ksd_rowsum_kernel_matches_oracle(gf, d):
    """ksd_rowsum_kernel (tile staging, the 4 x 4 register blocks with strided columns, triangular mask, row
    reduction), ksd_diag_kernel and the sequential scan - m not a multiple of the tile, several dimension chunks."""
    rng = np.random.default_rng(4000 + d)
    n = 333
    x, g = _mcmc_like(rng, n, d, dup=0.1)
    if gf:
        log_p = -0.5 * np.sum(x ** 2, axis=1)
        log_q = log_p + 2.0 * np.sin(x[:, 0]) - 0.1 * x[:, -1] ** 2       # weights within float32 range for any d
        ref = ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med')
        mk = lambda prec: H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=prec)   # noqa: E731
    else:
        ref = ot._make_stein_integrand(x, g, preconditioner='med')
        mk = lambda prec: H.SimProblem(x, g, preconditioner='med', precision=prec)                 # noqa: E731
    km_ref = os_.kmat(ref, n)
    ks_ref = os_.ksd(ref, n)
    low_ref = np.array([km_ref[t, :t].sum() for t in range(n)])
    scale = np.abs(km_ref) + 1e-3 * np.abs(km_ref).max()
    for prec, tol_k, tol_ks in (('f64', 1e-11, 1e-10), ('f64-fast', 1e-11, 1e-10), ('f32', 5e-4, 1e-5)):
        p = mk(prec)
        km = p.kmat()
        assert np.max(np.abs(km - km_ref) / scale) < tol_k, prec
        low, diag = p.ksd_rowsums()
        np.testing.assert_allclose(diag, np.diag(km_ref), rtol=tol_k * 10, err_msg=prec)
        assert np.max(np.abs(low - low_ref)) <= tol_ks * np.abs(km_ref).sum(), prec
        np.testing.assert_allclose(p.ksd(), ks_ref, rtol=tol_ks, atol=0, err_msg=prec)
    # selections with repeats, split over 1, 2 and 3 ranks
    sel = rng.integers(0, n, 150)
    want = os_.ksd(lambda a, b: ref(sel[a], sel[b]), 150)
    p = mk('f64')
    for n_parts in (1, 2, 3):
        np.testing.assert_allclose(p.ksd(sel=sel, n_parts=n_parts), want, rtol=1e-10, err_msg=str(n_parts))
    parts = [p.ksd_rowsums(sel=sel, part=r, n_parts=3) for r in range(3)]
    owners = (np.arange(150) // 64) % 3
    for r, (low, diag) in enumerate(parts):
        assert np.all(low[owners != r] == 0) and np.all(diag[owners != r] == 0)


def test_ksd_dense_preconditioner_and_pairs():
    rng = np.random.default_rng(77)
    n, d = 200, 5
    x, g = _mcmc_like(rng, n, d, dup=0.0)
    a = rng.standard_normal((d, d))
    dense = a @ a.T / d + np.eye(d)
    s, gs = ot._validate_and_standardize(x, g, True)
    want = lambda i, j: ok.vfk0_imq(s[i], s[j], gs[i], gs[j], dense)   # noqa: E731
    p = H.SimProblem(x, g, preconditioner=dense, precision='f64-fast')
    ia, ib = rng.integers(0, n, 700), rng.integers(0, n, 700)
    np.testing.assert_allclose(p.pairs(ia, ib), want(ia, ib), rtol=1e-9, atol=1e-12)
    km = np.array([want(np.full(n, i), np.arange(n)) for i in range(n)])
    np.testing.assert_allclose(p.ksd(), os_.ksd(lambda i, j: km[i, j], n), rtol=1e-10)
    idx, _ = p.thin(10)                                    # chunked kernel, DENSE instantiation
    a_run = np.diag(km).copy()
    expect = []
    for _ in range(10):
        j = int(np.argmin(a_run))
        expect.append(j)
        a_run += 2 * km[:, j]
    assert np.array_equal(idx, expect)


@pytest.mark.parametrize('d', [2, 4, 10, 64, 200])
def test_pair_values_match_oracle(d):
    rng = np.random.default_rng(3000 + d)
    x, g = _mcmc_like(rng, 500, d, dup=0.2)
    ia, ib = rng.integers(0, 500, 2000), rng.integers(0, 500, 2000)
    raw_ref = ot._make_stein_integrand(x, g, standardize=False, preconditioner='med')(ia, ib)
    raw = H.SimProblem(x, g, standardize=False, preconditioner='med').pairs(ia, ib)
    assert np.mean(raw == raw_ref) > 0.85
    assert np.max(np.abs(raw - raw_ref) / (np.abs(raw_ref) + 1e-3 * np.abs(raw_ref).max())) < 1e-13
    want = ot._make_stein_integrand(x, g, preconditioner='med')(ia, ib)
    scale = np.abs(want) + 1e-3 * np.abs(want).max()
    assert np.max(np.abs(H.SimProblem(x, g, preconditioner='med', precision='f64-fast').pairs(ia, ib) - want) / scale) < 1e-11
    assert np.max(np.abs(H.SimProblem(x, g, preconditioner='med', precision='f32').pairs(ia, ib) - want) / scale) < 2e-4


# ---- the passes in front of the pair kernels ----------------------------------------------------------------------
def test_standardise_median_and_log_weights():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((5000, 6)) * np.array([1, 10, 0.1, 3, 1e-3, 50.]) + 7
    g = rng.standard_normal((5000, 6))
    s_ref, g_ref = ot._validate_and_standardize(x, g, True)
    p = H.SimProblem(x, g, preconditioner='med')
    np.testing.assert_allclose(p.x, s_ref, rtol=1e-11)
    np.testing.assert_allclose(p.s, g_ref, rtol=1e-11)
    # the median heuristic itself is exact: same subsample rule, scipy's operation order, exact order statistics
    for n in (4321, 1000, 1001, 700, 2, 3):
        y = np.ascontiguousarray(rng.standard_normal((n, 5)))
        q = H.SimProblem(y, y, standardize=False, preconditioner='med')
        assert np.array_equal(q.linv, ok.make_precon(y, 'med')), n
    q = H.SimProblem(x[:, :3], g[:, :3], standardize=False, preconditioner='sclmed')
    assert np.array_equal(q.linv, ok.make_precon(x[:, :3], 'sclmed'))
    # log-weights and the range cap
    log_p, log_q = rng.standard_normal(5000) * 30, rng.standard_normal(5000) * 30
    q = H.SimProblem(x, g, log_p, log_q)
    assert np.array_equal(q.lw, log_q - log_p) and q.spread == np.ptp(log_q - log_p)
    q = H.SimProblem(x, g, log_p, log_q, range_cap=20.0)
    lw = log_q - log_p
    assert np.array_equal(q.lw, np.maximum(lw - lw.max(), -20.0))


def test_validation_errors_surface_from_the_kernels():
    rng = np.random.default_rng(12)
    x, g = rng.standard_normal((50, 2)), rng.standard_normal((50, 2))
    bad = x.copy(); bad[7, 1] = np.nan
    with pytest.raises(H.SimError, match='NaNs or infs'):
        H.SimProblem(bad, g)
    badg = g.copy(); badg[3, 0] = np.inf
    with pytest.raises(H.SimError, match='NaNs or infs'):
        H.SimProblem(x, badg)
    with pytest.raises(H.SimError, match='Too few unique samples'):
        H.SimProblem(np.ones((10, 2)), np.ones((10, 2)))
    lp = np.zeros(50); lp[4] = np.nan
    with pytest.raises(H.SimError, match='NaNs'):
        H.SimProblem(x, g, lp, np.zeros(50))


def test_gaussian_proxy_kernel_matches_scipy():
    """SURVEY 8f-1 (code/src/thinning.py:15-16): log q and grad log q of the Gaussian proxy."""
    import ctypes
    from scipy.stats import multivariate_normal as mvn
    lib = H.load()
    rng = np.random.default_rng(21)
    for d in (1, 2, 4, 9):
        n = 1111
        x = np.ascontiguousarray(rng.standard_normal((n, d)) * 2 + 1)
        a = rng.standard_normal((d, d))
        cov = a @ a.T + np.eye(d)
        mean = rng.standard_normal(d)
        prec = np.linalg.inv(cov)
        cst = d * np.log(2 * np.pi) + np.linalg.slogdet(cov)[1]
        xd, md, pd_ = H.dev(x), H.dev(mean), H.dev(prec)
        log_q, grad_q = H.guarded((n,), np.float64, fill=np.nan), H.guarded((n, d), np.float64, fill=np.nan)
        H.check(lib.gfk_gaussian_proxy(H.ptr(xd), n, d, H.ptr(md), H.ptr(pd_), ctypes.c_double(cst), H.ptr(log_q),
                                       H.ptr(grad_q), None))
        np.testing.assert_allclose(log_q, mvn.logpdf(x, mean=mean, cov=cov), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(grad_q, -np.einsum('ij,kj->ki', prec, x - mean), rtol=1e-12, atol=1e-12)


# ---- scheduling-order independence (a missing barrier shows up as a wrong result) -----------------------------------
@pytest.mark.parametrize('seed', [1, 2, 3])
def test_results_do_not_depend_on_thread_schedule(seed, monkeypatch):
    rng = np.random.default_rng(600)
    x, g = _mcmc_like(rng, 1200, 4)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = log_p / 1.2
    base = {}
    for strategy, env in STRATEGIES.items():
        for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        H.set_schedule(False)
        p = H.SimProblem(x, g, log_p, log_q, preconditioner='med')
        base[strategy] = p.thin(12)
        H.set_schedule(True, seed)
        q = H.SimProblem(x, g, log_p, log_q, preconditioner='med')
        assert q.lam == p.lam and np.array_equal(q.x, p.x) and np.array_equal(q.lw, p.lw)
        idx, obj = q.thin(12)
        assert np.array_equal(idx, base[strategy][0]) and np.array_equal(obj, base[strategy][1]), strategy
    H.set_schedule(False)
    x, g = _mcmc_like(rng, 300, 40)
    monkeypatch.setenv('GFK_PERSISTENT', '0')
    p = H.SimProblem(x, g, preconditioner='med', precision='f64-fast')
    want_idx, want_ks, want_km = p.thin(6)[0], p.ksd(), p.kmat()
    H.set_schedule(True, seed)
    assert np.array_equal(p.thin(6)[0], want_idx)
    assert np.array_equal(p.ksd(), want_ks) and np.array_equal(p.kmat(), want_km)


# ---- the emulator itself: the defects it exists to catch are reported, not silently tolerated ---------------------
def test_emulator_reports_deliberate_defects():
    import ctypes
    lib = H.load()
    lib.simt_selftest_error.restype = ctypes.c_char_p
    rc = lib.simt_selftest_divergent_barrier()
    assert rc != 0 and b'divergent __syncthreads' in lib.simt_selftest_error(rc)
    rc = lib.simt_selftest_misaligned()
    assert rc != 0 and b'misaligned address' in lib.simt_selftest_error(rc)
    blocks, T = 3, 128
    want = np.array([[((t + 1) % T) * 3 + b for t in range(T)] for b in range(blocks)], dtype=np.int32).reshape(-1)
    out = np.zeros(blocks * T, dtype=np.int32)
    P = ctypes.POINTER(ctypes.c_int)
    for seed in (1, 2, 3):
        H.set_schedule(True, seed)
        assert lib.simt_selftest_neighbour(out.ctypes.data_as(P), blocks, 1) == 0
        assert np.array_equal(out, want)                                  # with the barrier: any schedule
    wrong = 0
    for seed in (1, 2, 3):
        H.set_schedule(True, seed)
        out[:] = -1
        assert lib.simt_selftest_neighbour(out.ctypes.data_as(P), blocks, 0) == 0
        wrong += int(not np.array_equal(out, want))
    assert wrong == 3                                                     # without it: caught under every seed
    H.set_schedule(False)
    # cp.async is deferred like on the device: data read before cp.async.wait_group is the poison pattern, not data
    src, got = np.arange(64, dtype=np.float64), np.zeros(64)
    PD = ctypes.POINTER(ctypes.c_double)
    assert lib.simt_selftest_async_copy(src.ctypes.data_as(PD), got.ctypes.data_as(PD), 1) == 0
    assert np.array_equal(got, np.roll(src, -1))
    assert lib.simt_selftest_async_copy(src.ctypes.data_as(PD), got.ctypes.data_as(PD), 0) == 0
    assert np.isnan(got).sum() >= 60            # (a thread that has already exited has had its copies completed)
    for n in (64, 40, 33, 7):
        out = np.full(64, -1, dtype=np.int32)
        assert lib.simt_selftest_early_exit(out.ctypes.data_as(P), n) == 0
        for w in range(2):
            lanes = [t for t in range(32 * w, 32 * w + 32) if t < n]
            mask = sum(1 << (t % 32) for t in lanes)
            assert np.all(out[lanes].view(np.uint32) == mask), (n, w)
        assert np.all(out[n:] == -1)


# ---- multi-rank: one process per emulated rank, peer mailboxes in shared memory ----------------------------------
@pytest.mark.parametrize('world', [2, 3])
def test_multi_rank_thinning_and_ksd(world, tmp_path):
    """The MULTI instantiation of the single-launch kernel (gather by CTA 0 -> peer mailboxes -> re-publish), the
    mailbox exchange of the launch-per-step kernels, the all-gather fallback and the sharded KSD, on `world`
    emulated ranks: every rank must return the oracle's index sequence."""
    import subprocess
    import sys
    H.load()                                                  # build once, before the ranks race for it
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'simt_rank_worker.py')
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=600)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and 'SIMT_RANK_OK %d %d' % (r, world) in out, out[-3000:]


def test_strict_weights_double_double_when_safe_device_exp_otherwise(monkeypatch):
    """Strict mode forms exp(fl(lw_i + lw_j)) from per-row double-double exponentials (gfk_exp_weights_dd, 12 IEEE
    operations per pair) when gfk_log_weights_range shows every |lw| <= 350, and keeps exp() per pair otherwise.
    Both must select the oracle's indices; pair values agree with the oracle to the last bit wherever numpy's own
    exp and pow are correctly rounded."""
    rng = np.random.default_rng(81)
    n = 600
    x, g = _mcmc_like(rng, n, 3, dup=0.3)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    for shift, safe in ((0.0, True), (650.0, False)):
        log_q = log_p + np.sin(x[:, 0])
        log_q[::2] += shift            # lw_i = +650 and lw_j = -650 pairs: exp(lw_i) overflows, exp(lw_i + lw_j) does not
        log_q[1::2] -= shift
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ref = ot._make_stein_gf_integrand(x, log_p, log_q, g, standardize=False)
        p = H.SimProblem(x, g, log_p, log_q, standardize=False)
        assert np.allclose(p.lw_range, (np.min(log_q - log_p), np.max(log_q - log_p)), rtol=0, atol=0)
        assert (p.w is not None) == safe
        ia, ib = rng.integers(0, n, 3000), rng.integers(0, n, 3000)
        if not safe:
            ia, ib = 2 * (ia // 2), 2 * (ib // 2) + 1                    # (+650, -650) pairs: finite weights
        with np.errstate(over='ignore', invalid='ignore'):
            want = ref(ia, ib)
        got = p.pairs(ia, ib)
        np.testing.assert_allclose(got, want, rtol=1e-14)
        if safe:
            assert np.mean(got == want) > 0.85
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                expect = ot.thin_gf(x, log_p, log_q, g, 30, standardize=False)
            for env in STRATEGIES.values():
                for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
                    monkeypatch.delenv(k, raising=False)
                for k, v in env.items():
                    monkeypatch.setenv(k, v)
                assert np.array_equal(p.thin(30)[0], expect)
            np.testing.assert_allclose(p.ksd(), os_.ksd(ref, n), rtol=1e-10)
    # the device-exp route (forced) and the double-double route agree to 1 ulp per weight and select the same points
    log_q = log_p + np.sin(x[:, 0])
    a = H.SimProblem(x, g, log_p, log_q, preconditioner='med')
    monkeypatch.setenv('GFK_TEST_NO_EXP_DD', '1')
    b = H.SimProblem(x, g, log_p, log_q, preconditioner='med')
    assert a.w is not None and b.w is None
    assert np.array_equal(a.thin(25)[0], b.thin(25)[0])
    ia, ib = rng.integers(0, n, 2000), rng.integers(0, n, 2000)
    np.testing.assert_allclose(a.pairs(ia, ib), b.pairs(ia, ib), rtol=4e-16)


def test_double_double_exponentials_are_accurate_to_2_pow_minus_90():
    import ctypes
    from decimal import Decimal, getcontext
    lib = H.load()
    rng = np.random.default_rng(5)
    l = np.concatenate([rng.uniform(-350, 350, 600), rng.uniform(-3, 3, 400), [0.0, 350.0, -350.0, 400.0]])
    ld, ew = H.dev(l), H.guarded((l.size, 2), np.float64, fill=0.0)
    H.check(lib.gfk_exp_weights_dd(H.ptr(ld), H.ptr(ew), l.size, None))
    assert np.all(np.isnan(ew[-1]))                                      # beyond the limit: NaN marker, never used
    getcontext().prec = 60
    for v, (hi, lo) in zip(l[:-1], ew[:-1]):
        t = Decimal(float(v)).exp()
        assert abs((Decimal(float(hi)) + Decimal(float(lo)) - t) / t) < Decimal(2) ** -90
        assert float(hi) == float(t)                                     # the high part alone is the correctly rounded exp


# ---- edge sizes and BASELINE-shaped inputs -----------------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 3, 31, 32, 33, 511, 512, 513, 1025])
def test_edge_sizes_all_strategies(n, monkeypatch):
    """Fewer rows than threads, exactly one CTA, one row more than a CTA, a single row; m = 1; every d path."""
    rng = np.random.default_rng(900 + n)
    for d in (1, 2, 3, 4, 5, 9):
        x = rng.standard_normal((n, d))
        g = -x * (1 + 0.1 * np.arange(d))
        if n > 3:
            x[n // 2], g[n // 2] = x[1], g[1]                     # an exact duplicate
        for m in (1, 2, 7):
            expect = ot.thin(x, g, m, standardize=(n > 1))
            for env in STRATEGIES.values():
                for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
                    monkeypatch.delenv(k, raising=False)
                for k, v in env.items():
                    monkeypatch.setenv(k, v)
                idx, obj = H.SimProblem(x, g, standardize=(n > 1)).thin(m)
                assert np.array_equal(idx, expect), (n, d, m, env)
                assert np.all(np.isfinite(obj))
        if n >= 2:
            p = H.SimProblem(x, g, standardize=(n > 1), preconditioner='id', precision='f64-fast')
            ref = ot._make_stein_integrand(x, g, standardize=(n > 1))
            np.testing.assert_allclose(p.ksd(), os_.ksd(ref, n), rtol=1e-10)


def test_config_c1_full_size_and_c2_prefix_through_the_emulator():
    """BASELINE config C1 in full (n = 10^4, d = 2, m = 100, gradient-free, 'med') and a prefix of the C2-shaped
    chain (d = 4, runs of exactly duplicated rows, score magnitudes 1e2-1e3) at 60 000 rows on 5 emulated SMs."""
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    os.environ['SIMT_SMS'] = '5'
    try:
        c = synthetic.config_c1()
        expect = ot.thin_gf(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], c['m'], preconditioner='med')
        idx, _ = H.SimProblem(c['sample'], c['gradient_q'], c['log_p'], c['log_q'], preconditioner='med').thin(c['m'])
        assert np.array_equal(idx, expect)
        c = synthetic.config_c2(n=60_000, m=30)
        expect = ot.thin(c['sample'], c['gradient'], 30, preconditioner='med')
        for precision in ('f64', 'f64-fast'):
            idx, _ = H.SimProblem(c['sample'], c['gradient'], preconditioner='med', precision=precision).thin(30)
            assert np.array_equal(idx, expect), precision
        first_of_run = np.concatenate([[True], np.any(c['sample'][1:] != c['sample'][:-1], axis=1)])
        assert np.all(first_of_run[idx])
    finally:
        os.environ.pop('SIMT_SMS', None)


@pytest.mark.parametrize('d', [1, 4, 10, 32, 40])
def test_ksd_row_sums_do_not_depend_on_the_thread_schedule(d):
    """ksd_rowsum_kernel (float64: 4 x 2 pairs per thread, 512 threads; float32: 4 x 4, 256 threads) double-buffers the
    column tiles with cp.async between two barriers per stage: its row sums must be bit-identical under any thread
    schedule, for the whole triangle and for a selection dealt over two parts, and match the oracle's ksd."""
    rng = np.random.default_rng(8000 + d)
    n = 330                                                # 6 row tiles
    x, g = _mcmc_like(rng, n, d, dup=0.1)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = log_p + np.sin(x[:, 0])
    sel = rng.integers(0, n, 200)
    for prec in ('f64', 'f32'):
        for gf in (False, True):
            p = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=prec) if gf else \
                H.SimProblem(x, g, preconditioner='med', precision=prec)
            H.set_schedule(False)
            base, base_sel = p.ksd_rowsums(), p.ksd_rowsums(sel=sel, part=1, n_parts=2)
            for seed in (1, 2):
                H.set_schedule(True, seed)
                got, got_sel = p.ksd_rowsums(), p.ksd_rowsums(sel=sel, part=1, n_parts=2)
                assert np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1]), (prec, gf, seed)
                assert np.array_equal(got_sel[0], base_sel[0]), (prec, gf, seed)
            H.set_schedule(False)
    ref = ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med')
    np.testing.assert_allclose(H.SimProblem(x, g, log_p, log_q, preconditioner='med').ksd(), os_.ksd(ref, n), rtol=1e-10)


@pytest.mark.parametrize('d', [6, 10, 16, 18, 64, 100])
def test_tma_fed_step_kernel_equals_the_slab_kernel(d, monkeypatch):
    """GFK_STEP_TMA=1: thin_step_tma_kernel (a producer warp keeps three [256 rows] x [128 bytes] slabs of x and s in
    flight with cp.async.bulk.tensor + mbarriers, SWIZZLE_128B tiles, eight consumer warps, thread per row) must select
    the oracle's indices and leave bit-identical objectives to thin_step_chunked_kernel, in every precision, under any
    thread schedule.  The emulator completes a bulk copy only when some thread next tests an mbarrier and poisons its
    destination until then, so a consumer that does not wait for `full`, or a producer that does not wait for `empty`,
    reads NaNs."""
    monkeypatch.setenv('GFK_PERSISTENT', '0')
    monkeypatch.setenv('SIMT_SMS', '2')                          # 2 CTAs x 256 rows: several tiles per CTA, ring wraps
    rng = np.random.default_rng(9000 + d)
    x, g = _mcmc_like(rng, 1100, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1) / d
    log_q = log_p + 0.5 * np.sin(x[:, 0])
    m = 6
    expect = ot.thin_gf(x, log_p, log_q, g, m, preconditioner='med')
    lib = H.load()
    for precision in ('f64', 'f64-fast', 'f32'):
        monkeypatch.delenv('GFK_STEP_TMA', raising=False)
        base = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=precision).thin(m)
        monkeypatch.setenv('GFK_STEP_TMA', '1')
        for random, seed in ((False, 0), (True, 5)):
            H.set_schedule(random, seed)
            lib.simt_tma_loads.restype = ctypes.c_longlong
            before = lib.simt_tma_loads()
            idx, obj = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=precision).thin(m)
            eligible = (d * (4 if precision == 'f32' else 8)) % 16 == 0       # the tensor-map stride rule
            assert (lib.simt_tma_loads() > before) == eligible, 'TMA kernel ran / did not run against its eligibility rule'
            assert np.array_equal(idx, base[0]) and np.array_equal(obj, base[1]), (precision, random)
        H.set_schedule(False)
        if precision != 'f32':
            assert np.array_equal(base[0], expect), precision
    # score-based kernel (no weights) and a shard that starts in the middle of the sample
    monkeypatch.delenv('GFK_STEP_TMA', raising=False)
    p = H.SimProblem(x, g, preconditioner='med')
    want = p.thin(5)
    monkeypatch.setenv('GFK_STEP_TMA', '1')
    got = p.thin(5)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert np.array_equal(got[0], ot.thin(x, g, 5, preconditioner='med'))
    # dense preconditioner and odd d (rows not a multiple of 16 bytes): not eligible, must fall back, not fail
    a = rng.standard_normal((d, d))
    dense = a @ a.T / d + np.eye(d)
    p = H.SimProblem(x[:300], g[:300], preconditioner=dense, precision='f64-fast')
    monkeypatch.delenv('GFK_STEP_TMA', raising=False)
    want = p.thin(4)
    monkeypatch.setenv('GFK_STEP_TMA', '1')
    before = lib.simt_tma_loads()
    got = p.thin(4)
    assert lib.simt_tma_loads() == before
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    xo, go = _mcmc_like(rng, 700, d + 1)
    assert np.array_equal(H.SimProblem(xo, go, preconditioner='med').thin(4)[0], ot.thin(xo, go, 4, preconditioner='med'))
    assert lib.simt_tma_loads() == before


@pytest.mark.parametrize('sms', [3, 7, 12])
def test_batched_chains_in_one_launch_equal_one_job_per_chain(sms, monkeypatch):
    """gfk_thin_batched: several chains of different lengths in ONE cooperative launch,
    one group of CTAs per chain with its own workspace and slots, must return exactly what one gfk_thin per chain
    returns - index vectors and running objectives bit for bit, i.e. the oracle's sequences - for the score-based and
    the gradient-free kernel, in every precision, under random thread schedules."""
    monkeypatch.setenv('SIMT_SMS', str(sms))
    rng = np.random.default_rng(4200 + sms)
    m = 25
    for d, gf in ((4, True), (2, False), (3, True)):
        chains = []
        for n in (700, 1500, 2300):
            x, g = _mcmc_like(rng, n, d)
            log_p = -0.5 * np.sum(x ** 2, axis=1)
            chains.append((x, g, log_p, log_p + 0.5 * np.sin(x[:, 0])))
        for precision in ('f64', 'f64-fast', 'f32'):
            mk = (lambda c: H.SimProblem(c[0], c[1], c[2], c[3], preconditioner='med', precision=precision)) if gf else \
                 (lambda c: H.SimProblem(c[0], c[1], preconditioner='med', precision=precision))
            probs = [mk(c) for c in chains]
            single = [p.thin(m) for p in probs]
            for random, seed in ((False, 0), (True, 3)):
                H.set_schedule(random, seed)
                got = H.thin_batched(probs, m)
                assert got is not None, 'an eligible set was refused'
                for c in range(len(chains)):
                    assert np.array_equal(got[0][c], single[c][0]), (d, gf, precision, c, random)
                    assert np.array_equal(got[1][c], single[c][1]), (d, gf, precision, c, random)
            H.set_schedule(False)
            if precision == 'f64':
                for c, ch in enumerate(chains):
                    want = ot.thin_gf(ch[0], ch[2], ch[3], ch[1], m, preconditioner='med') if gf else \
                        ot.thin(ch[0], ch[1], m, preconditioner='med')
                    assert np.array_equal(single[c][0], want), (d, gf, c)
    # not eligible -> refused without launching anything: d > 4, mixed kernel kinds, more chains than SMs
    x5, g5 = _mcmc_like(rng, 400, 5)
    assert H.thin_batched([H.SimProblem(x5, g5), H.SimProblem(x5, g5)], 5) is None
    x4, g4 = _mcmc_like(rng, 400, 4)
    lp = -0.5 * np.sum(x4 ** 2, axis=1)
    assert H.thin_batched([H.SimProblem(x4, g4), H.SimProblem(x4, g4, lp, lp + 0.1 * x4[:, 0])], 5) is None
    assert H.thin_batched([H.SimProblem(x4, g4) for _ in range(sms + 1)], 5) is None
    one = H.thin_batched([H.SimProblem(x4, g4)], 5)                      # a set of one is just a job
    assert np.array_equal(one[0][0], ot.thin(x4, g4, 5))


@pytest.mark.parametrize('d', [1, 4, 10, 32, 40, 64, 72, 130])
def test_tensor_core_ksd_row_sums(d):
    """gfk_ksd_rowsums_mma (csrc/ksd_mma.cuh): the four Gram tiles of the KSD on tcgen05.mma kind::tf32 with 3xTF32
    operand splits, operands by TMA into SWIZZLE_128B tiles, accumulators in tensor memory, the tile kernel's pair
    epilogue on top.  The emulator decodes the shared-memory and instruction descriptors the kernel builds (bit layouts
    of cute::UMMA), keeps 11 significant bits of every operand and accumulates in float.  float32 mode: KSD values
    within 1e-5 relative of the float64 oracle (BASELINE.json), row sums close to the SIMT float32 kernel's, identical
    under random thread schedules; a selection with repeats; ineligible problems refused."""
    rng = np.random.default_rng(7700 + d)
    n = 330                                                  # 3 row tiles of 128, 6 column tiles of 64
    x, g = _mcmc_like(rng, n, d, dup=0.1)
    x = x + 3.0                                              # an offset the centring has to remove
    g = g / np.sqrt(d)
    log_p = -0.5 * np.sum((x - 3.0) ** 2, axis=1) / d
    log_q = log_p + 0.3 * np.sin(x[:, 0])
    lib = H.load()
    lib.simt_umma_issued.restype = ctypes.c_longlong
    for gf in (False, True):
        p = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision='f32') if gf else \
            H.SimProblem(x, g, preconditioner='med', precision='f32')
        ref = ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med') if gf else \
            ot._make_stein_integrand(x, g, preconditioner='med')
        H.set_schedule(False)
        before = lib.simt_umma_issued()
        low, diag = p.ksd_rowsums_mma()
        assert lib.simt_umma_issued() > before
        simt_low, simt_diag = p.ksd_rowsums()
        assert np.array_equal(diag, simt_diag)
        scale = np.maximum(np.abs(simt_low), np.max(np.abs(simt_low)) * 1e-3)
        assert np.max(np.abs(low - simt_low) / scale) < 2e-4, np.max(np.abs(low - simt_low) / scale)
        np.testing.assert_allclose(p.ksd_from_rowsums(low, diag), os_.ksd(ref, n), rtol=1e-5)
        H.set_schedule(True, 4)
        low2, _ = p.ksd_rowsums_mma()
        assert np.array_equal(low2, low), gf
        H.set_schedule(False)
        # the triangle dealt over three parts (128-row tiles round-robin): the parts' vectors add up to the whole
        parts = [p.ksd_rowsums_mma(part=k, n_parts=3) for k in range(3)]
        assert np.array_equal(sum(q[0] for q in parts), low) and np.array_equal(sum(q[1] for q in parts), diag)
        assert all(np.count_nonzero(q[1]) <= 128 for q in parts)
        sel = rng.integers(0, n, 200)
        lo_s, dg_s = p.ksd_rowsums_mma(sel=sel)
        np.testing.assert_allclose(p.ksd_from_rowsums(lo_s, dg_s), os_.ksd(lambda i, j: ref(sel[i], sel[j]), 200), rtol=1e-5)
    # not eligible: float64 modes, a dense preconditioner
    assert H.SimProblem(x, g, preconditioner='med', precision='f64-fast').ksd_rowsums_mma() is None
    if d > 1:                                                # a 1 x 1 matrix is a scalar preconditioner
        a = rng.standard_normal((d, d))
        assert H.SimProblem(x, g, preconditioner=a @ a.T / d + np.eye(d), precision='f32').ksd_rowsums_mma() is None


def test_a_lost_bulk_copy_ends_in_an_error_not_in_a_hang(monkeypatch):
    """The mbarrier waits of the TMA-fed step kernel and of the tensor-core KSD are bounded (csrc/tma.cuh dev_mbar_wait):
    if the bytes a barrier expects never arrive, the waiter raises the launch's abort flag after its deadline, every
    other wait of the kernel returns at once, the kernel terminates (tensor memory freed, no barrier left hanging) and
    the check call reports GFK_ECUDA.  The emulator loses one chosen bulk copy; run in a subprocess because the
    deadline is read once per process."""
    import subprocess
    import sys
    code = '''
import ctypes, os, sys, time
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import simt_harness as H
rng = np.random.default_rng(3)
x = rng.standard_normal((900, 64)); g = -x
lib = H.load()
lib.simt_drop_tma.argtypes = [ctypes.c_longlong]
os.environ["GFK_PERSISTENT"] = "0"; os.environ["GFK_STEP_TMA"] = "1"; os.environ["SIMT_SMS"] = "2"
p = H.SimProblem(x, g, preconditioner="med")
good = p.thin(4)[0]
for which, after in (("step", 5), ("step", 0)):
    lib.simt_drop_tma(after)
    t0 = time.time()
    try:
        p.thin(4)
    except H.SimError as e:
        assert "deadline" in str(e), str(e)
    else:
        raise AssertionError("a lost copy went unnoticed")
    assert time.time() - t0 < 60
lib.simt_drop_tma(-1)
assert np.array_equal(p.thin(4)[0], good)                 # the next job is unaffected
q = H.SimProblem(x[:400], g[:400], preconditioner="med", precision="f32")
ok = q.ksd_rowsums_mma()
for after in (0, 9, 20):                                  # a row operand, a column tile of the first / a later stage
    lib.simt_drop_tma(after)
    try:
        q.ksd_rowsums_mma()
    except H.SimError as e:
        assert "deadline" in str(e), str(e)
    else:
        raise AssertionError("a lost copy went unnoticed")
lib.simt_drop_tma(-1)
again = q.ksd_rowsums_mma()
assert np.array_equal(again[0], ok[0])
print("LOST_COPY_OK")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    H.load()
    env = dict(os.environ, SIMT_MBAR_LIMIT_MS='1500')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and 'LOST_COPY_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_full_size_c3_on_the_b200_grid_shape(monkeypatch):
    """BASELINE config C3 at full size (n = 10^6, d = 4, gradient-free, 'med') with the launch geometry of a B200:
    148 concurrently running CTAs of 512 threads, 14 rows per thread of which 3 stay in shared memory, every CTA
    polling all 148 flag-in-data slots (5 per lane).  A prefix of the greedy search must equal the oracle's."""
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    monkeypatch.setenv('SIMT_SMS', '148')
    c = synthetic.config_c3()
    m = 8
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        expect = ot.thin_gf(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], m, preconditioner='med')
    assert len(np.unique(expect)) == m                       # a healthy proxy: distinct selections
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'full_size_indices.npz'))
    assert np.array_equal(golden['c3_idx'][:m], expect)      # the committed full-length sequence is the oracle's
    assert np.array_equal(golden['c3_idx'], golden['c3_idx_kernels']) and np.array_equal(golden['c2_idx'], golden['c2_idx_kernels'])
    for precision in ('f64', 'f64-fast'):
        p = H.SimProblem(c['sample'], c['gradient_q'], c['log_p'], c['log_q'], preconditioner='med', precision=precision)
        for tmem in ('0', '1'):                  # 1: 8 of every thread's 14 rows parked in tensor memory (opt-in)
            monkeypatch.setenv('GFK_TMEM', tmem)
            idx, _ = p.thin(m)
            assert np.array_equal(idx, expect), (precision, tmem)


@pytest.mark.parametrize('d', [2, 4])
def test_tensor_memory_row_store_equals_the_default_kernel(d, monkeypatch):
    """GFK_TMEM=1: full passes of rows parked in tensor memory (tcgen05.st once, tcgen05.ld every step; the emulator
    checks the lane-quarter rule, convergence of the .sync.aligned instructions, allocation bounds and the final
    dealloc).  Indices and the final objective must be bit-identical to the default kernel and equal the oracle."""
    monkeypatch.setenv('SIMT_SMS', '2')
    rng = np.random.default_rng(9100 + d)
    n = 11 * 1024 + 333                                         # 12 passes: 8 (11) in TMEM, the rest in shared memory
    x, g = _mcmc_like(rng, n, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = log_p + 0.5 * np.sin(x[:, 0])
    m = 6
    expect = ot.thin_gf(x, log_p, log_q, g, m, preconditioner='med')
    for precision in ('f64', 'f64-fast') + (('f32',) if d == 4 else ()):
        monkeypatch.delenv('GFK_TMEM', raising=False)
        base = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=precision).thin(m)
        monkeypatch.setenv('GFK_TMEM', '1')
        lib = H.load()
        lib.simt_tmem_access_count.restype = ctypes.c_longlong
        for random, seed in ((False, 0), (True, 3)):
            H.set_schedule(random, seed)
            before = lib.simt_tmem_access_count()
            idx, obj = H.SimProblem(x, g, log_p, log_q, preconditioner='med', precision=precision).thin(m)
            assert np.array_equal(idx, base[0]) and np.array_equal(obj, base[1]), (precision, random)
            # (passes in tensor memory) x 1024 threads: one store each, one load per greedy update; of the 12 passes
            # the 11 full ones qualify: 8 fit at d = 4 (16 columns per row), all 11 at d = 2
            passes, rem = divmod(lib.simt_tmem_access_count() - before, 1024 * m)
            assert rem == 0 and passes == (11 if (d == 2 or precision == 'f32') else 8)
        H.set_schedule(False)
        if precision != 'f32':
            assert np.array_equal(base[0], expect), precision
    # score-based kernel and a shard that fits shared memory entirely (no tensor-memory pass: falls back silently)
    monkeypatch.setenv('GFK_TMEM', '1')
    assert np.array_equal(H.SimProblem(x, g, preconditioner='med').thin(m)[0], ot.thin(x, g, m, preconditioner='med'))
    assert np.array_equal(H.SimProblem(x[:900], g[:900]).thin(m)[0], ot.thin(x[:900], g[:900], m))
