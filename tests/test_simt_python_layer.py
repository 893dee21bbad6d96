"""The Python host layer of the product (argument handling, preconditioner selection, warnings, ctypes marshalling,
the proxies / chains modules) exercised on the CPU: for these tests ONLY, the library handle is swapped for the
SIMT-emulator build of libgfk (tests/host/libgfk_sim.so) and "the device" is host memory.  The product itself has no
such switch - `_engine._require_cuda` raises without a CUDA device and `_lib.load()` without libgfk.so; both are
monkeypatched here, nothing else.  Cases follow tests/test_gpu_parity.py / tests/test_zz_next_rows.py at small sizes.
"""
import contextlib
import ctypes
import warnings

import numpy as np
import pytest
import torch

import simt_harness as H
from oracle import next_rows as nr, repo_glue
from oracle.stein_thinning_oracle import kernel as ok, stein as os_, thinning as ot


@pytest.fixture
def api(monkeypatch):
    import types
    from gradient_free_mcmc_postprocessing_b200 import _engine, _lib, chains, kernel, proxies, stein, thinning
    from gradient_free_mcmc_postprocessing_b200.utils import ksd as uksd
    for k in ('GFK_PERSISTENT', 'GFK_RESIDENT', 'SIMT_SMS', 'GFK_PRECISION'):
        monkeypatch.delenv(k, raising=False)
    H.set_schedule(False)
    monkeypatch.setattr(_lib, '_lib', H.load())
    cpu = torch.device('cpu')
    for mod in (_engine, proxies):
        monkeypatch.setattr(mod, '_require_cuda', lambda device=None: cpu)
        monkeypatch.setattr(mod, '_stream_ptr', lambda device=None: ctypes.c_void_p(0))
    monkeypatch.setattr(torch.cuda, 'device', lambda dev: contextlib.nullcontext())
    return types.SimpleNamespace(kernel=kernel, stein=stein, thinning=thinning, uksd=uksd, proxies=proxies,
                                 chains=chains, engine=_engine)


def _mcmc_like(rng, n, d, dup=0.5):
    base = rng.standard_normal((n, d)) @ (np.eye(d) + 0.3 * rng.standard_normal((d, d)))
    keep = rng.random(n) > dup
    keep[0] = True
    src = np.maximum.accumulate(np.where(keep, np.arange(n), 0))
    x = base[src]
    return np.ascontiguousarray(x), np.ascontiguousarray(-(x - 0.1) * (1.0 + 0.2 * np.arange(d)))


def test_golden_vectors_through_the_public_api(api, g13, g45):
    t = api.thinning
    for precision in ('f64', 'f64-fast'):
        idx = t.thin(g13['sample'], g13['gradient'], 20, precision=precision)
        assert idx.dtype == np.uint32 and np.array_equal(idx, g13['idx_thin'])
        idx = t.thin_gf(g13['sample'], g13['log_p'], g13['log_q'], g13['gradient_q'], 20, precision=precision)
        assert np.array_equal(idx, g13['idx_thin_gf'])
    assert len(np.unique(t.thin(g45['sample'], g45['gradient'], 1000, preconditioner='med'))) == 538
    with pytest.warns(UserWarning, match='log_q differs from log_p by more than 10'):
        idx = t.thin_gf(g45['sample'], g45['log_p'], g45['log_q_laplace'], g45['gradient_q_laplace'], 400,
                        preconditioner='med')
    assert np.array_equal(np.unique(idx), [494])
    # the reference's own glue (fixtures produced by code/src/thinning.py and code/src/utils/ksd.py)
    for on_device in (False, True):
        idx = t.gaussian_thin(g45['sample'], g45['log_p'], g45['sample_mean'], g45['sample_cov'], 50,
                              proxy_on_device=on_device)
        assert np.array_equal(idx, g45['glue_gaussian_thin_50']), on_device
    ks = api.uksd.calculate_ksd(g45['sample'], g45['gradient'], g45['glue_idx_for_ksd'])
    np.testing.assert_allclose(ks, g45['glue_calculate_ksd'], rtol=1e-10, atol=0)
    # torch tensors in, as a caller with device-resident data passes them
    assert np.array_equal(t.thin(torch.from_numpy(g13['sample']), torch.from_numpy(g13['gradient']), 20), g13['idx_thin'])


def test_preconditioners_kernel_module_and_integrand_protocol(api, g13):
    rng = np.random.default_rng(6)
    x, g = _mcmc_like(rng, 900, 3)
    t = api.thinning
    for kwargs in (dict(standardize=False), dict(preconditioner='sclmed'), dict(preconditioner='0.7')):
        assert np.array_equal(t.thin(x, g, 25, **kwargs), ot.thin(x, g, 25, **kwargs)), kwargs
    integ = t._make_stein_integrand(x, g, preconditioner='smpcov')
    ref = ot._make_stein_integrand(x, g, preconditioner='smpcov')
    ia, ib = rng.integers(0, 900, 1500), rng.integers(0, 900, 1500)
    a, b = integ(ia, ib), ref(ia, ib)
    assert np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.abs(b).max())) < 1e-10
    y = rng.standard_normal((1300, 5))
    assert np.array_equal(api.kernel.make_precon(y, 'med'), ok.make_precon(y, 'med'))
    np.testing.assert_allclose(api.kernel.make_precon(y, 'smpcov'), ok.make_precon(y, 'smpcov'), rtol=1e-10)
    xs, ys = rng.standard_normal((200, 3)), rng.standard_normal((200, 3))
    sx, sy = rng.standard_normal((200, 3)), rng.standard_normal((200, 3))
    linv = np.identity(3) * 0.37
    np.testing.assert_allclose(api.kernel.vfk0_imq(xs, ys, sx, sy, linv), ok.vfk0_imq(xs, ys, sx, sy, linv), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(api.kernel.vfk0_imq(xs, ys[:1], sx, sy[:1], linv, c=2.0),
                               ok.vfk0_imq(xs, ys[:1], sx, sy[:1], linv, c=2.0), rtol=1e-10, atol=1e-13)
    with pytest.raises(NotImplementedError):
        api.kernel.vfk0_imq(xs, ys, sx, sy, linv, beta=-0.7)
    # code/tests/test_ksd.py's contract on the device integrand, and slice / list indexing as _greedy_search uses it
    integ = t._make_stein_integrand(g13['sample'][:5], g13['gradient'][:5])
    mat = api.stein.kmat(integ, 5)
    perm = np.array([2, 0, 4, 3, 1])
    np.testing.assert_array_equal(api.stein.kmat(api.uksd.reindex_integrand(integ, perm), 5), mat[np.ix_(perm, perm)])
    col = t._make_stein_integrand(x, g)(slice(None), [5])
    np.testing.assert_allclose(col, ot._make_stein_integrand(x, g)(slice(None), [5]), rtol=1e-9, atol=1e-12)
    # a foreign callable is driven through the protocol on the host
    fake = np.arange(25, dtype=float).reshape(5, 5)
    fake = fake + fake.T
    assert np.array_equal(api.stein.kmat(lambda i, j: fake[i, j], 5), fake)

    def fake_integrand(i, j):                       # paired rows, broadcasting like the library's own closures
        a, b = np.broadcast_arrays(np.arange(5)[i] if isinstance(i, slice) else np.asarray(i),
                                   np.arange(5)[j] if isinstance(j, slice) else np.asarray(j))
        return fake[a, b]
    assert np.array_equal(t._greedy_search(4, fake_integrand), ot._greedy_search(4, fake_integrand))
    np.testing.assert_allclose(api.stein.ksd(fake_integrand, 5), os_.ksd(fake_integrand, 5))


@pytest.mark.parametrize('gf', [False, True])
def test_ksd_and_kmat_through_the_public_api(api, gf):
    rng = np.random.default_rng(4000)
    n, d = 300, 10
    x, g = _mcmc_like(rng, n, d, dup=0.1)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = log_p + np.sin(x[:, 0])
    if gf:
        ref = ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med')
        mk = lambda prec: api.thinning._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med', precision=prec)   # noqa: E731
    else:
        ref = ot._make_stein_integrand(x, g, preconditioner='med')
        mk = lambda prec: api.thinning._make_stein_integrand(x, g, preconditioner='med', precision=prec)                  # noqa: E731
    km_ref, ks_ref = os_.kmat(ref, n), os_.ksd(ref, n)
    scale = np.abs(km_ref) + 1e-3 * np.abs(km_ref).max()
    for prec, tol_k, tol_ks in (('f64', 1e-11, 1e-10), ('f64-fast', 1e-11, 1e-10), ('f32', 5e-4, 1e-5)):
        integ = mk(prec)
        assert np.max(np.abs(api.stein.kmat(integ, n) - km_ref) / scale) < tol_k, prec
        np.testing.assert_allclose(api.stein.ksd(integ, n), ks_ref, rtol=tol_ks, atol=0, err_msg=prec)
    sel = rng.integers(0, n, 111)
    ks = api.stein.ksd(api.uksd.reindex_integrand(mk('f64'), sel), 111)
    np.testing.assert_allclose(ks, os_.ksd(repo_glue.reindex_integrand(ref, sel), 111), rtol=1e-10)


def test_tensor_core_ksd_through_the_public_api(api, monkeypatch):
    """GFK_KSD_MMA=1 routes ksd() of a float32, 32 <= d <= 64 problem through gfk_ksd_rowsums_mma; anything else keeps
    the tile kernel.  Values within the float32 bar (1e-5 relative) of the float64 oracle."""
    rng = np.random.default_rng(61)
    x, g = _mcmc_like(rng, 300, 48, dup=0.1)
    g = g / 7.0
    ref = ot._make_stein_integrand(x, g, preconditioner='med')
    want = os_.ksd(ref, 300)
    lib = H.load()
    lib.simt_umma_issued.restype = ctypes.c_longlong
    monkeypatch.setenv('GFK_KSD_MMA', '1')
    integ = api.thinning._make_stein_integrand(x, g, preconditioner='med', precision='f32')
    before = lib.simt_umma_issued()
    np.testing.assert_allclose(api.stein.ksd(integ, 300), want, rtol=1e-5)
    assert lib.simt_umma_issued() > before
    before = lib.simt_umma_issued()
    integ64 = api.thinning._make_stein_integrand(x, g, preconditioner='med', precision='f64')
    np.testing.assert_allclose(api.stein.ksd(integ64, 300), want, rtol=1e-10)
    assert lib.simt_umma_issued() == before                     # float64: the tile kernel
    monkeypatch.delenv('GFK_KSD_MMA')
    np.testing.assert_allclose(api.stein.ksd(integ, 300), want, rtol=1e-5)
    assert lib.simt_umma_issued() == before                     # not opted in: the tile kernel


def test_validation_errors_through_the_public_api(api):
    rng = np.random.default_rng(12)
    x, g = rng.standard_normal((50, 2)), rng.standard_normal((50, 2))
    t = api.thinning
    with pytest.raises(ValueError):
        t.thin(x, g[:-1], 5)
    bad = x.copy(); bad[7, 1] = np.nan
    with pytest.raises(ValueError):
        t.thin(bad, g, 5)
    with pytest.raises(ValueError):
        t.thin(np.zeros((0, 2)), np.zeros((0, 2)), 5)
    with pytest.raises(ValueError):
        t.thin(np.ones((10, 2)), np.ones((10, 2)), 3)
    with pytest.raises(Exception):
        t.thin(np.ones((10, 2)), np.ones((10, 2)), 3, standardize=False, preconditioner='med')
    with pytest.raises(ValueError):
        t.thin(x, g, 5, preconditioner='nonsense')
    with pytest.raises(ValueError):
        t.thin_gf(x, np.zeros(49), np.zeros(50), g, 5)
    lp = np.zeros(50); lp[4] = np.nan
    with pytest.raises(ValueError):
        t.thin_gf(x, lp, np.zeros(50), g, 5)
    with pytest.raises(ValueError):
        t.thin(x, g, 0)


def test_next_rows_through_the_public_api(api, g13, g45, tmp_path):
    from scipy import stats
    from scipy.stats import gaussian_kde
    p = api.proxies
    x, log_p = g45['sample'], g45['log_p']
    # KDE proxy: the reference's unique-point counts (Gaussian_mixture.ipynb#c67)
    log_q, grad_q = p.kde_logpdf_and_score(x)
    np.testing.assert_allclose(log_q, gaussian_kde(x.T, bw_method='silverman').logpdf(x.T), rtol=1e-12, atol=1e-12)
    assert len(np.unique(p.kde_thin(x, log_p, 1000))) == 481
    w = np.exp(log_p) / np.sum(np.exp(log_p))
    assert len(np.unique(p.kde_thin(x, log_p, 1000, weights=w))) == 343
    # Student-t proxy
    rng = np.random.default_rng(44)
    d = 4
    y = rng.standard_normal((3000, d)) * 2 + 1
    a = rng.standard_normal((d, d))
    sigma = a @ a.T / d + 0.5 * np.eye(d)
    mu = rng.standard_normal(d)
    np.testing.assert_allclose(p.t_log_pdf(y, mu, sigma, 9.0), stats.multivariate_t.logpdf(y, loc=mu, shape=sigma, df=9.0),
                               rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(p.t_grad_log_pdf(y, mu, sigma, 9.0), nr.t_grad_log_pdf(y, mu, sigma, 9.0), rtol=1e-12, atol=1e-13)
    lp = -0.5 * np.sum(y ** 2, axis=1)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert np.array_equal(p.thin_gf_t(y, lp, mu, sigma, 9.0, 20), nr.thin_gf_t(y, lp, mu, sigma, 9.0, 20))
    # energy distance
    xv, yv = rng.standard_normal((700, 3)), rng.standard_normal((90, 3)) * 1.2
    e2 = p.cumulative_energy_distance(xv, yv)
    for t in (0, 5, 89):
        assert abs(e2[t] - nr.energy_distance(xv, yv[:t + 1])) < 1e-12
    assert abs(p.energy_distance(xv, yv) - nr.energy_distance(xv, yv)) < 1e-12
    assert abs(p.fit_quality(xv, yv, 10) - nr.fit_quality(xv, yv, 10)) < 1e-12
    with pytest.raises(ValueError):
        p.cumulative_energy_distance(xv, yv[:, :2])
    sel = rng.integers(0, 700, 60)
    sizes = np.array([5, 17, 60])
    curve = p.energy_distance_curve(xv, xv, sel, sizes, validation_sample_step=7)      # Comparison.ipynb#c22
    assert curve.shape == (3, 2) and np.array_equal(curve[:, 0], sizes)
    for size, val in curve:
        assert abs(val - nr.fit_quality(xv, xv[sel[:int(size)]], 7)) < 1e-11
    with pytest.raises(ValueError):
        p.energy_distance_curve(xv, xv, sel, [0, 3])
    # chains + .npy layer
    paths = []
    for i in range(3):
        sp, gp = tmp_path / f'rw_samples_{i}.npy', tmp_path / f'rw_grads_{i}.npy'
        np.save(sp, g13['sample'][i * 100:])
        np.save(gp, g13['gradient'][i * 100:])
        paths.append((sp, gp))
    outs = [tmp_path / f'rw_thinned_idx_{i}.npy' for i in range(3)]
    res = api.chains.thin_chains([q[0] for q in paths], [q[1] for q in paths], 20, out_paths=outs)
    assert np.array_equal(res[0], g13['idx_thin'])
    for i in range(3):
        assert np.array_equal(np.load(outs[i]), ot.thin(g13['sample'][i * 100:], g13['gradient'][i * 100:], 20))
    mine = api.chains.thin_chains([q[0] for q in paths], [q[1] for q in paths], 20, rank=1, world=2)
    assert mine[0] is None and mine[2] is None and np.array_equal(mine[1], res[1])
    gf = api.chains.thin_gf_chains([g13['sample']] * 2, [g13['log_p']] * 2, [g13['log_q']] * 2, [g13['gradient_q']] * 2, 20)
    assert all(np.array_equal(r, g13['idx_thin_gf']) for r in gf)


def test_product_still_refuses_to_run_without_cuda():
    """Outside the fixture nothing is patched: the real entry checks must raise on this GPU-less machine."""
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from gradient_free_mcmc_postprocessing_b200 import _engine
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _engine._require_cuda(None)


def test_driver_smoke_entry_runs_over_the_emulator(api, monkeypatch, capsys):
    """__graft_entry__.smoke() is GPU-only; its Python (fixture names, API calls, tolerances) is checked here by
    running it over the emulator build - the CUDA availability assertion is the only thing patched away."""
    import __graft_entry__ as entry
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    entry.smoke()
    assert 'smoke ok' in capsys.readouterr().out


@pytest.mark.parametrize('world', [2, 3])
def test_public_api_under_a_process_group(world):
    """`distributed.init()` + thin / thin_gf / ksd as called under torchrun, on `world` emulated ranks (gloo for the
    plumbing, fake NCCL + emulated cudaIpc for the library): every rank returns the oracle's sequence, through the
    peer mailboxes (single-launch and launch-per-step kernels) and through the all-gather fallback."""
    import socket
    import subprocess
    import sys
    H.load()
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = str(s.getsockname()[1])
    import os
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'simt_dist_worker.py')
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), port], stdout=subprocess.PIPE,
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
        assert p.returncode == 0 and 'SIMT_DIST_OK %d %d' % (r, world) in out, out[-3000:]
