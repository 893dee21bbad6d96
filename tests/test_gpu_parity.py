"""GPU parity tests: libgfk (through the public Python API, i.e. through the C ABI) against the
oracle on identical inputs, against the committed golden fixtures, and - at BASELINE.json's full
sizes - through size-independent properties.

Bar (BASELINE.json north_star): float64 selected-index sequences are bit-exact; KSD values agree
within 1e-10 relative; the float32 mode is held to 1e-5 relative on KSD values.
"""
import warnings

import numpy as np
import pytest

from oracle import repo_glue
from oracle.stein_thinning_oracle import kernel as ok, stein as os_, thinning as ot

pytestmark = pytest.mark.gpu

KSD_RTOL_F64 = 1e-10      # north_star tolerance, float64
KSD_RTOL_F32 = 1e-5       # north_star tolerance, float32 fast mode


@pytest.fixture(scope='module')
def api(gpu_device):
    from gradient_free_mcmc_postprocessing_b200 import kernel, stein, thinning
    from gradient_free_mcmc_postprocessing_b200.utils import ksd as uksd
    import types
    return types.SimpleNamespace(kernel=kernel, stein=stein, thinning=thinning, uksd=uksd, device=gpu_device)


def _mcmc_like(rng, n, d, dup=0.5):
    """Random sample with runs of exactly duplicated rows (exact ties)."""
    base = rng.standard_normal((n, d)) @ (np.eye(d) + 0.3 * rng.standard_normal((d, d)))
    keep = rng.random(n) > dup
    keep[0] = True
    src = np.maximum.accumulate(np.where(keep, np.arange(n), 0))
    x = base[src]
    grad = -(x - 0.1) * (1.0 + 0.2 * np.arange(d))
    return np.ascontiguousarray(x), np.ascontiguousarray(grad)


# ---- golden vectors ---------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['f64', 'f64-fast'])
def test_golden_g1_g2_g3(api, g13, precision):
    t = api.thinning
    idx = t.thin(g13['sample'], g13['gradient'], 20, precision=precision)
    assert idx.dtype == np.uint32 and np.array_equal(idx, g13['idx_thin'])
    idx = t.thin_gf(g13['sample'], g13['log_p'], g13['log_p'], g13['gradient'], 20, precision=precision)
    assert np.array_equal(idx, g13['idx_thin'])
    idx = t.thin_gf(g13['sample'], g13['log_p'], g13['log_q'], g13['gradient_q'], 20, precision=precision)
    assert np.array_equal(idx, g13['idx_thin_gf'])


def test_golden_g4_unique_counts(api, g45):
    t = api.thinning
    idx = t.thin(g45['sample'], g45['gradient'], 1000, preconditioner='med')
    assert len(np.unique(idx)) == 538
    assert np.array_equal(idx, ot.thin(g45['sample'], g45['gradient'], 1000, preconditioner='med'))
    idx = t.thin_gf(g45['sample'], g45['log_p'], g45['log_q'], g45['gradient_q'], 1000, preconditioner='med')
    assert len(np.unique(idx)) == 400


def test_golden_g5_stuck_index_and_warning(api, g45):
    with pytest.warns(UserWarning, match='log_q differs from log_p by more than 10'):
        idx = api.thinning.thin_gf(g45['sample'], g45['log_p'], g45['log_q_laplace'], g45['gradient_q_laplace'],
                                   10_000, preconditioner='med')
    assert idx.shape == (10_000,) and np.array_equal(np.unique(idx), [494])


def test_reference_glue_fixtures(api, g45):
    idx = api.thinning.gaussian_thin(g45['sample'], g45['log_p'], g45['sample_mean'], g45['sample_cov'], 50)
    assert np.array_equal(idx, g45['glue_gaussian_thin_50'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        idx = api.thinning.gaussian_thin(g45['sample'], g45['log_p'], g45['laplace_mean'], g45['laplace_cov'], 30)
    assert np.array_equal(idx, g45['glue_gaussian_thin_laplace_30'])
    ks = api.uksd.calculate_ksd(g45['sample'], g45['gradient'], g45['glue_idx_for_ksd'])
    np.testing.assert_allclose(ks, g45['glue_calculate_ksd'], rtol=KSD_RTOL_F64, atol=0)


# ---- seeded parity against the oracle -----------------------------------------------------------------
@pytest.mark.parametrize('d', [1, 2, 3, 4, 5, 7, 8, 9, 10, 12, 16, 20, 33, 64, 128, 130, 200])
@pytest.mark.parametrize('pre', ['id', 'med'])
def test_thin_indices_match_oracle(api, d, pre):
    rng = np.random.default_rng(1000 + d)
    x, g = _mcmc_like(rng, 3000, d)
    m = 60
    expect = ot.thin(x, g, m, preconditioner=pre)
    for precision in ('f64', 'f64-fast'):
        got = api.thinning.thin(x, g, m, preconditioner=pre, precision=precision)
        assert np.array_equal(got, expect), (precision, d, pre, np.flatnonzero(got != expect)[:3])


@pytest.mark.parametrize('d', [2, 4, 10, 64])
@pytest.mark.parametrize('range_cap', [None, 5.0])
def test_thin_gf_indices_match_oracle(api, d, range_cap):
    rng = np.random.default_rng(2000 + d)
    x, gq = _mcmc_like(rng, 2500, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1) + 0.3 * np.sin(x[:, 0])
    log_q = -0.5 * np.sum((x - 0.1) ** 2, axis=1) / 1.3
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        expect = ot.thin_gf(x, log_p, log_q, gq, 50, preconditioner='med', range_cap=range_cap)
        for precision in ('f64', 'f64-fast'):
            got = api.thinning.thin_gf(x, log_p, log_q, gq, 50, preconditioner='med', range_cap=range_cap,
                                       precision=precision)
            assert np.array_equal(got, expect), (precision, d, range_cap)


@pytest.mark.parametrize('d', [1, 2, 3, 4])
def test_launch_strategies_agree(api, d, monkeypatch):
    """persistent cooperative kernel (rows resident in shared memory / streamed through L2) and the
    one-launch-per-step path must select identical indices and leave the same objective."""
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    rng = np.random.default_rng(500 + d)
    x, g = _mcmc_like(rng, 70_001, d)
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = -0.5 * np.sum(x ** 2, axis=1) / 1.1
    expect = ot.thin_gf(x, log_p, log_q, g, 25, preconditioner='med')
    for precision in ('f64', 'f64-fast', 'f32'):
        results = []
        for env in ({}, {'GFK_RESIDENT': '0'}, {'GFK_PERSISTENT': '0'}):
            monkeypatch.delenv('GFK_RESIDENT', raising=False)
            monkeypatch.delenv('GFK_PERSISTENT', raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            prob = SteinProblem(x, g, log_p=log_p, log_q=log_q, preconditioner='med', precision=precision)
            idx, obj = prob.thin(200, return_objective=True)
            results.append((idx, obj.cpu().numpy()))
        for idx, obj in results[1:]:
            assert np.array_equal(idx, results[0][0]), precision
            assert np.array_equal(obj, results[0][1]), precision
        if precision != 'f32':
            assert np.array_equal(results[0][0][:25], expect)


def test_selection_with_replacement_m_greater_than_n(api):
    rng = np.random.default_rng(5)
    x, g = _mcmc_like(rng, 40, 2, dup=0.0)
    expect = ot.thin(x, g, 300)
    got = api.thinning.thin(x, g, 300)
    assert np.array_equal(got, expect)


def test_unstandardised_and_other_preconditioners(api):
    rng = np.random.default_rng(6)
    x, g = _mcmc_like(rng, 1500, 3)
    for kwargs in (dict(standardize=False), dict(preconditioner='sclmed'), dict(preconditioner='0.7')):
        assert np.array_equal(api.thinning.thin(x, g, 40, **kwargs), ot.thin(x, g, 40, **kwargs)), kwargs
    # dense preconditioner: numpy's BLAS order is not reproducible -> compare kernel values, not indices
    integ = api.thinning._make_stein_integrand(x, g, preconditioner='smpcov')
    ref = ot._make_stein_integrand(x, g, preconditioner='smpcov')
    ia, ib = rng.integers(0, 1500, 4000), rng.integers(0, 1500, 4000)
    a, b = integ(ia, ib), ref(ia, ib)
    assert np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.abs(b).max())) < 1e-10
    got = api.thinning.thin(x, g, 40, preconditioner='smpcov')
    assert got.shape == (40,) and got[0] == ot.thin(x, g, 40, preconditioner='smpcov')[0]


def test_median_subsample_rule_above_1000_rows(api):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((4321, 5))
    assert np.array_equal(api.kernel.make_precon(x, 'med'), ok.make_precon(x, 'med'))       # bit-exact
    assert np.array_equal(api.kernel.make_precon(x[:700], 'med'), ok.make_precon(x[:700], 'med'))
    assert np.array_equal(api.kernel.make_precon(x, 'sclmed'), ok.make_precon(x, 'sclmed'))
    np.testing.assert_allclose(api.kernel.make_precon(x, 'smpcov'), ok.make_precon(x, 'smpcov'), rtol=1e-10)


def test_standardisation_matches_numpy(api):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((5000, 6)) * np.array([1, 10, 0.1, 3, 1e-3, 50.]) + 7
    g = rng.standard_normal((5000, 6))
    s_ref, g_ref = ot._validate_and_standardize(x, g, True)
    s_got, g_got = api.thinning._validate_and_standardize(x, g, True)
    # column 4 is 7 + 1e-3 * N(0,1): its mean absolute deviation is conditioned 1e4 : 1, and numpy's strictly
    # sequential column sum carries ~sqrt(n) ulp of rounding that the chunked device reduction does not
    np.testing.assert_allclose(s_got, s_ref, rtol=1e-11)
    np.testing.assert_allclose(g_got, g_ref, rtol=1e-11)
    x2 = rng.standard_normal((3000, 3))
    s_ref, g_ref = ot._validate_and_standardize(x2, g[:3000, :3], True)
    s_got, g_got = api.thinning._validate_and_standardize(x2, g[:3000, :3], True)
    np.testing.assert_allclose(s_got, s_ref, rtol=1e-14)
    np.testing.assert_allclose(g_got, g_ref, rtol=1e-14)


# ---- kernel values, Gram matrix, KSD ----------------------------------------------------------------
@pytest.mark.parametrize('d', [2, 4, 10, 64, 200])
def test_pair_values_match_oracle(api, d):
    rng = np.random.default_rng(3000 + d)
    x, g = _mcmc_like(rng, 800, d, dup=0.2)
    ref = ot._make_stein_integrand(x, g, preconditioner='med')
    ia, ib = rng.integers(0, 800, 5000), rng.integers(0, 800, 5000)
    want = ref(ia, ib)
    scale = np.abs(want) + 1e-3 * np.abs(want).max()
    strict = api.thinning._make_stein_integrand(x, g, preconditioner='med', precision='f64')(ia, ib)
    assert np.max(np.abs(strict - want) / scale) < 1e-12
    # with identical kernel inputs (no standardisation: its column sums are reduced in a different order
    # from numpy's) the strict mode is bit-equal wherever numpy's own pow is correctly rounded
    raw_ref = ot._make_stein_integrand(x, g, standardize=False, preconditioner='med')(ia, ib)
    raw = api.thinning._make_stein_integrand(x, g, standardize=False, preconditioner='med', precision='f64')(ia, ib)
    assert np.mean(raw == raw_ref) > 0.85
    assert np.max(np.abs(raw - raw_ref) / (np.abs(raw_ref) + 1e-3 * np.abs(raw_ref).max())) < 1e-13
    fast = api.thinning._make_stein_integrand(x, g, preconditioner='med', precision='f64-fast')(ia, ib)
    assert np.max(np.abs(fast - want) / scale) < 1e-11
    f32 = api.thinning._make_stein_integrand(x, g, preconditioner='med', precision='f32')(ia, ib)
    assert np.max(np.abs(f32 - want) / scale) < 2e-4
    # slice / list protocol, as _greedy_search uses it
    col = api.thinning._make_stein_integrand(x, g)(slice(None), [5])
    np.testing.assert_allclose(col, ot._make_stein_integrand(x, g)(slice(None), [5]), rtol=1e-9, atol=1e-12)


def test_vfk0_imq_direct(api):
    rng = np.random.default_rng(11)
    x, y = rng.standard_normal((300, 3)), rng.standard_normal((300, 3))
    sx, sy = rng.standard_normal((300, 3)), rng.standard_normal((300, 3))
    linv = np.identity(3) * 0.37
    np.testing.assert_allclose(api.kernel.vfk0_imq(x, y, sx, sy, linv), ok.vfk0_imq(x, y, sx, sy, linv),
                               rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(api.kernel.vfk0_imq(x, y[:1], sx, sy[:1], linv, c=2.0),
                               ok.vfk0_imq(x, y[:1], sx, sy[:1], linv, c=2.0), rtol=1e-10, atol=1e-13)
    a = rng.standard_normal((3, 3))
    dense = a @ a.T + np.eye(3)
    np.testing.assert_allclose(api.kernel.vfk0_imq(x, y, sx, sy, dense), ok.vfk0_imq(x, y, sx, sy, dense),
                               rtol=1e-9, atol=1e-12)
    vf = api.kernel.make_imq(x, 'med')
    np.testing.assert_allclose(vf(x, y, sx, sy), ok.make_imq(x, 'med')(x, y, sx, sy), rtol=1e-10, atol=1e-13)


def test_reference_test_ksd_on_device_integrand(api, g13):
    """code/tests/test_ksd.py's contract (kmat of a re-indexed integrand = permuted matrix), with the
    device integrand in place of the fake one."""
    integ = api.thinning._make_stein_integrand(g13['sample'][:5], g13['gradient'][:5])
    mat = api.stein.kmat(integ, 5)
    perm = np.array([2, 0, 4, 3, 1])
    res = api.stein.kmat(api.uksd.reindex_integrand(integ, perm), 5)
    np.testing.assert_array_equal(res, mat[np.ix_(perm, perm)])
    np.testing.assert_array_equal(mat, mat.T)


@pytest.mark.parametrize('gf', [False, True])
@pytest.mark.parametrize('d', [2, 10, 40])
def test_kmat_and_ksd_match_oracle(api, gf, d):
    rng = np.random.default_rng(4000 + d)
    n = 700
    x, g = _mcmc_like(rng, n, d, dup=0.1)
    if gf:
        log_p = -0.5 * np.sum(x ** 2, axis=1)
        log_q = -0.5 * np.sum(x ** 2, axis=1) / 1.2
        ref = ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med')
        mk = lambda prec: api.thinning._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner='med',  # noqa: E731
                                                                precision=prec)
    else:
        ref = ot._make_stein_integrand(x, g, preconditioner='med')
        mk = lambda prec: api.thinning._make_stein_integrand(x, g, preconditioner='med', precision=prec)  # noqa: E731
    km_ref = os_.kmat(ref, n)
    ks_ref = os_.ksd(ref, n)
    scale = np.abs(km_ref) + 1e-3 * np.abs(km_ref).max()
    for prec, tol_k, tol_ks in (('f64', 1e-11, KSD_RTOL_F64), ('f64-fast', 1e-11, KSD_RTOL_F64),
                                ('f32', 5e-4, KSD_RTOL_F32)):
        integ = mk(prec)
        km = api.stein.kmat(integ, n)
        assert np.max(np.abs(km - km_ref) / scale) < tol_k, prec
        ks = api.stein.ksd(integ, n)
        np.testing.assert_allclose(ks, ks_ref, rtol=tol_ks, atol=0, err_msg=prec)
    sel = rng.integers(0, n, 333)
    ks = api.stein.ksd(api.uksd.reindex_integrand(mk('f64'), sel), 333)
    np.testing.assert_allclose(ks, os_.ksd(repo_glue.reindex_integrand(ref, sel), 333), rtol=KSD_RTOL_F64)


# ---- error behaviour -----------------------------------------------------------------------------------
def test_validation_errors(api):
    rng = np.random.default_rng(12)
    x, g = rng.standard_normal((50, 2)), rng.standard_normal((50, 2))
    t = api.thinning
    with pytest.raises(ValueError):
        t.thin(x, g[:-1], 5)
    bad = x.copy(); bad[7, 1] = np.nan
    with pytest.raises(ValueError):
        t.thin(bad, g, 5)
    badg = g.copy(); badg[3, 0] = np.inf
    with pytest.raises(ValueError):
        t.thin(x, badg, 5)
    with pytest.raises(ValueError):
        t.thin(np.zeros((0, 2)), np.zeros((0, 2)), 5)
    with pytest.raises(ValueError):                       # zero scale: constant column
        t.thin(np.ones((10, 2)), np.ones((10, 2)), 3)
    with pytest.raises(Exception):                        # zero median
        t.thin(np.ones((10, 2)), np.ones((10, 2)), 3, standardize=False, preconditioner='med')
    with pytest.raises(ValueError):
        t.thin(x, g, 5, preconditioner='nonsense')
    with pytest.raises(ValueError):
        t.thin_gf(x, np.zeros(49), np.zeros(50), g, 5)
    lp = np.zeros(50); lp[4] = np.nan
    with pytest.raises(ValueError):
        t.thin_gf(x, lp, np.zeros(50), g, 5)


def test_device_tensor_inputs(api, g13):
    import torch
    xs = torch.from_numpy(g13['sample']).to(api.device)
    gs = torch.from_numpy(g13['gradient']).to(api.device)
    assert np.array_equal(api.thinning.thin(xs, gs, 20), g13['idx_thin'])
    pinned = torch.from_numpy(g13['sample']).pin_memory()
    assert np.array_equal(api.thinning.thin(pinned, gs, 20), g13['idx_thin'])


# ---- BASELINE.json sizes: size-independent properties ----------------------------------------------------
def _objective_property_check(api, prob, idx, objective, rng, rtol):
    """A_i after m selections must equal k(i,i) + 2 sum_{t<m-1} k(i, idx_t) for any row i, and
    the last selection must be the first minimiser of the objective it was picked from."""
    from gradient_free_mcmc_postprocessing_b200.integrand import SteinIntegrand
    integ = SteinIntegrand(prob)
    rows = np.unique(np.concatenate([rng.integers(0, prob.n, 64), idx[-3:].astype(np.int64)]))
    a = objective.cpu().numpy()
    for i in rows:
        want = integ([i], [i])[0] + 2 * np.sum(integ(np.full(len(idx) - 1, i), idx[:-1].astype(np.int64)))
        assert abs(a[i] - want) <= rtol * max(1.0, abs(want)), (i, a[i], want)
    assert int(np.argmin(a)) == int(idx[-1])


def test_config_c1_full_size_matches_oracle(api):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    c = synthetic.config_c1()
    expect = ot.thin_gf(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], c['m'], preconditioner='med')
    got = api.thinning.thin_gf(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], c['m'], preconditioner='med')
    assert np.array_equal(got, expect)


def test_config_c2_prefix_matches_oracle_and_properties(api):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    c = synthetic.config_c2()
    expect = ot.thin(c['sample'], c['gradient'], 40, preconditioner='med')        # oracle: seconds
    prob = SteinProblem(c['sample'], c['gradient'], preconditioner='med', precision='f64')
    idx, objective = prob.thin(c['m'], return_objective=True)
    assert np.array_equal(idx[:40], expect)
    again = prob.thin(c['m'])
    assert np.array_equal(idx, again)                                                # deterministic
    _objective_property_check(api, prob, idx, objective, np.random.default_rng(0), 1e-9)
    # duplicated rows: the first of every run of identical rows is the one ever selected
    first_of_run = np.concatenate([[True], np.any(c['sample'][1:] != c['sample'][:-1], axis=1)])
    assert np.all(first_of_run[idx])


@pytest.mark.parametrize('precision', ['f64', 'f64-fast'])
def test_config_c3_full_size_properties(api, precision):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
    c = synthetic.config_c3()
    prob = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med',
                        precision=precision)
    idx, objective = prob.thin(c['m'], return_objective=True)
    assert idx.shape == (1000,) and idx.max() < 10 ** 6
    expect = ot.thin_gf(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], 12, preconditioner='med')
    assert np.array_equal(idx[:12], expect)
    _objective_property_check(api, prob, idx, objective, np.random.default_rng(1), 1e-9)


def test_config_c5_prefix_matches_oracle(api):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    c = synthetic.config_c5(n=100_000, m=30)
    expect = ot.thin(c['sample'], c['gradient'], 30, preconditioner='med')
    got = api.thinning.thin(c['sample'], c['gradient'], 30, preconditioner='med')
    assert np.array_equal(got, expect)


def test_config_c4_ksd_subsample_and_linearity(api):
    from gradient_free_mcmc_postprocessing_b200 import synthetic
    c = synthetic.config_c4(n=20_000)
    integ = api.thinning._make_stein_gf_integrand(c['sample'], c['log_p'], c['log_q'], c['gradient_q'],
                                                  preconditioner='med', precision='f64-fast')
    ref = ot._make_stein_gf_integrand(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], preconditioner='med')
    ks = api.stein.ksd(integ, 20_000)
    np.testing.assert_allclose(ks[:1500], os_.ksd(ref, 1500), rtol=KSD_RTOL_F64)
    # checksum: (n * ks[-1])^2 is the full Gram sum; compare with a blocked numpy evaluation of the oracle integrand
    total = 0.0
    n = 20_000
    for b in range(0, n, 50):
        rows = np.repeat(np.arange(b, min(b + 50, n)), n)
        cols = np.tile(np.arange(n), min(b + 50, n) - b)
        if b < 200:       # 200 full rows of the Gram matrix: seconds on the CPU
            total += np.sum(ref(rows, cols))
    km_rows = integ(np.repeat(np.arange(200), n), np.tile(np.arange(n), 200))
    assert abs(np.sum(km_rows) - total) <= 1e-10 * abs(total)
    assert np.isfinite(ks).all() and ks[-1] > 0
