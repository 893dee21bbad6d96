"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/gfk.h
declares, the product fails loudly without a GPU, the integrand protocol (code/tests/test_ksd.py)
works for foreign callables, and the synthetic generators reproduce the reference fixture."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from gradient_free_mcmc_postprocessing_b200 import _lib, synthetic
from gradient_free_mcmc_postprocessing_b200.stein import kmat, ksd
from gradient_free_mcmc_postprocessing_b200.thinning import _greedy_search, thin
from gradient_free_mcmc_postprocessing_b200.utils.ksd import reindex_integrand

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'gfk.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gfk_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f'{name} declared in include/gfk.h but not exported'
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)
    assert lib.gfk_version() == 100
    assert lib.gfk_thin_ws_bytes() > 0 and lib.gfk_standardize_ws_bytes(10 ** 6, 4) > 0
    assert lib.gfk_comm_size() == 1 and lib.gfk_comm_rank() == 0


def test_problem_struct_layout_matches_header():
    # 5 pointers, int64, 2 x int32, 4 doubles
    assert ctypes.sizeof(_lib.Problem) == 5 * 8 + 8 + 4 + 4 + 4 * 8


def test_argument_errors_without_gpu():
    lib = _lib.load()
    rc = lib.gfk_thin(None, 0, 0, 1, None, None, None, None)
    assert rc == _lib.GFK_EINVAL and b'NULL' in lib.gfk_last_error()
    with pytest.raises(ValueError):
        _lib.mode_code('f16')


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_fails_loudly_without_gpu(g13):
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        thin(g13['sample'], g13['gradient'], 5)


def test_reference_test_ksd_through_product_protocol(ref_test_ksd):
    # port of code/tests/test_ksd.py:8-28: a foreign integrand driven through kmat + reindex_integrand
    mat = ref_test_ksd['mat']
    new_integrand = reindex_integrand(lambda a, b: mat[a, b], ref_test_ksd['indices'])
    np.testing.assert_array_equal(kmat(new_integrand, mat.shape[0]), ref_test_ksd['expected'])


def test_foreign_callable_greedy_and_ksd():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((40, 40))
    gram = a @ a.T
    integrand = lambda i, j: gram[i, j] if not isinstance(i, slice) else np.diag(gram)  # noqa: E731
    idx = _greedy_search(7, lambda i, j: (np.diag(gram) if isinstance(j, slice) else gram[:, j[0]]))
    running = np.diag(gram).copy()
    expect = [int(np.argmin(running))]
    for _ in range(6):
        running += 2 * gram[:, expect[-1]]
        expect.append(int(np.argmin(running)))
    assert idx.dtype == np.uint32 and list(idx) == expect
    ks = ksd(lambda i, j: gram[np.asarray(i), np.asarray(j)], 12)
    assert np.allclose(ks, [np.sqrt(gram[:t + 1, :t + 1].sum()) / (t + 1) for t in range(12)])
    del integrand


def test_synthetic_generator_reproduces_reference_sample(g45):
    rvs, logpdf, score = synthetic.make_mvn_mixture(synthetic.C1_WEIGHTS, synthetic.C1_MEANS, synthetic.C1_COVS)
    sample = rvs(1000, np.random.default_rng(12345))
    assert np.array_equal(sample, g45['sample'])
    assert np.allclose(logpdf(sample), g45['log_p'], rtol=0, atol=1e-12)
    assert np.allclose(score(sample), g45['gradient'], rtol=0, atol=1e-11)


def test_synthetic_configs_shapes():
    c1 = synthetic.config_c1()
    assert c1['sample'].shape == (10_000, 2) and c1['m'] == 100
    c2 = synthetic.config_c2(n=5000)
    assert c2['sample'].shape == (5000, 4)
    assert len(np.unique(c2['sample'], axis=0)) < 2500           # MCMC-like duplicated rows


def test_install_as_stein_thinning_aliases():
    import sys
    import gradient_free_mcmc_postprocessing_b200 as pkg
    saved = {k: sys.modules.get(k) for k in ('stein_thinning', 'stein_thinning.thinning', 'stein_thinning.kernel',
                                             'stein_thinning.stein')}
    try:
        pkg.install_as_stein_thinning()
        from stein_thinning.kernel import make_imq, make_precon, vfk0_imq  # noqa: F401
        from stein_thinning.stein import kmat as k2, ksd as s2  # noqa: F401
        from stein_thinning.thinning import (_greedy_search, _make_stein_gf_integrand, _make_stein_integrand,  # noqa: F401
                                             _validate_and_standardize, thin as t2, thin_gf)  # noqa: F401
        import inspect
        assert list(inspect.signature(thin_gf).parameters)[:8] == [
            'sample', 'log_p', 'log_q', 'gradient_q', 'n_points', 'standardize', 'preconditioner', 'range_cap']
        assert list(inspect.signature(t2).parameters)[:5] == ['sample', 'gradient', 'n_points', 'standardize',
                                                               'preconditioner']
        assert inspect.signature(t2).parameters['preconditioner'].default == 'id'
        assert inspect.signature(thin_gf).parameters['range_cap'].default is None
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


class _FakeProblem:
    """Stands in for a device-resident SteinProblem: k(i, j) = table[i, j], so the integrand's index plumbing
    (slices, lists, broadcasting, re-indexing) can be checked on the CPU."""

    def __init__(self, n, seed=0):
        rng = np.random.default_rng(seed)
        a = rng.standard_normal((n, n))
        self.table, self.n = a + a.T, n
        self.calls = []

    def pairs(self, ia, ib):
        self.calls.append((np.shape(ia), np.shape(ib)))
        return self.table[np.asarray(ia), np.asarray(ib)]

    def ksd(self, sel, m):
        sel = np.arange(m) if sel is None else np.asarray(sel)
        sub = self.table[np.ix_(sel, sel)]
        return np.array([np.sqrt(abs(sub[:t + 1, :t + 1].sum())) / (t + 1) for t in range(m)])

    def kmat(self, sel, m):
        sel = np.arange(m) if sel is None else np.asarray(sel)
        return self.table[np.ix_(sel, sel)]


def test_stein_integrand_index_protocol_on_fake_problem():
    from gradient_free_mcmc_postprocessing_b200.integrand import SteinIntegrand
    prob = _FakeProblem(12)
    integ = SteinIntegrand(prob)
    t = prob.table
    # the three call shapes _greedy_search / ksd / kmat use (JAX_Stein_Thinning.ipynb#c17,#c22; test_ksd.py:16-20)
    np.testing.assert_array_equal(integ(slice(None), slice(None)), np.diag(t))
    np.testing.assert_array_equal(integ(slice(None), [5]), t[:, 5])
    np.testing.assert_array_equal(integ(np.arange(4), [7]), t[:4, 7])
    rows, cols = np.triu_indices(12)
    np.testing.assert_array_equal(integ(rows, cols), t[rows, cols])
    np.testing.assert_array_equal(integ([3], [3]), t[[3], [3]])
    # re-indexing composes (code/src/utils/ksd.py:9-16) and is honoured by kmat / ksd
    perm = np.array([2, 0, 4, 3, 1, 11, 7])
    sub = reindex_integrand(integ, perm)
    np.testing.assert_array_equal(sub(slice(None), [1]), t[perm, perm[1]])
    np.testing.assert_array_equal(kmat(sub, 5), t[np.ix_(perm[:5], perm[:5])])
    again = reindex_integrand(sub, np.array([6, 0, 3]))
    np.testing.assert_array_equal(kmat(again, 3), t[np.ix_(perm[[6, 0, 3]], perm[[6, 0, 3]])])
    assert ksd(sub, 4).shape == (4,)
    with pytest.raises(NotImplementedError):
        sub.thin(3)


def test_pick_winner_matches_device_comparator_semantics():
    """distributed.pick_winner is the host mirror of gfk::cand_better (stein_math.cuh)."""
    from gradient_free_mcmc_postprocessing_b200.distributed import pick_winner
    assert pick_winner([1.0, 1.0, 2.0], [9, 4, 1]) == 4            # ties -> lowest index
    assert pick_winner([1.0, np.nan, 0.0], [0, 5, 9]) == 5         # NaN beats every number (np.argmin)
    assert pick_winner([np.nan, np.nan], [8, 3]) == 3
    assert pick_winner([np.inf, np.inf], [2, 1]) == 1
    assert pick_winner([-np.inf, 0.0], [7, 0]) == 7


def test_next_row_host_logic(tmp_path):
    """kde_bandwidth = scipy's gaussian_kde covariance and weights; the chain layer's dealing and .npy round trip."""
    from scipy.stats import gaussian_kde
    from gradient_free_mcmc_postprocessing_b200 import chains, proxies
    rng = np.random.default_rng(0)
    x = rng.standard_normal((300, 3))
    w = rng.random(300)
    for weights in (None, w):
        for bw in ('silverman', 'scott', 0.3):
            kde = gaussian_kde(x.T, bw_method=bw, weights=weights)
            h, wn = proxies.kde_bandwidth(x, weights, bw)
            np.testing.assert_allclose(h, kde.covariance, rtol=1e-13)
            np.testing.assert_allclose(wn, kde.weights, rtol=1e-15)
    with pytest.raises(ValueError):
        proxies.kde_bandwidth(x, w[:-1])
    with pytest.raises(ValueError):
        proxies.kde_bandwidth(x, None, 'nonsense')
    assert chains._my_chains(5, 0, 2) == [0, 2, 4] and chains._my_chains(5, 1, 2) == [1, 3]
    p = chains.save_indices(tmp_path / 'idx.npy', np.array([3, 1, 2]))
    got = chains.load_array(p)
    assert got.dtype == np.uint32 and np.array_equal(got, [3, 1, 2])
    assert chains.load_array(x) is x
    with pytest.raises(ValueError):
        chains.thin_chains([x], [x, x], 3)
