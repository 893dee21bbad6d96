"""Drop-in check: the reference's OWN modules (code/src/thinning.py, code/src/utils/ksd.py) and its own unit test
(code/tests/test_ksd.py) run unchanged on top of this package after ``install_as_stein_thinning()``.

CPU only and only where the reference checkout exists (this authoring container): the reference tree is not on the
GPU box, so the test skips there.  The library handle is the SIMT-emulator build (see tests/test_simt_python_layer.py).
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import pytest

from test_simt_python_layer import api  # noqa: F401  (fixture: emulator build behind the public API)

REF = '/root/reference/code'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'src')), reason='reference checkout not present')


@pytest.fixture
def reference_modules(api, monkeypatch):   # noqa: F811
    import gradient_free_mcmc_postprocessing_b200 as gfk
    saved = {k: sys.modules.get(k) for k in ('stein_thinning', 'stein_thinning.thinning', 'stein_thinning.kernel',
                                             'stein_thinning.stein', 'thinning', 'utils', 'utils.ksd')}
    gfk.install_as_stein_thinning()
    monkeypatch.syspath_prepend(os.path.join(REF, 'src'))
    for k in ('thinning', 'utils', 'utils.ksd'):
        sys.modules.pop(k, None)
    import thinning as ref_thinning        # /root/reference/code/src/thinning.py
    from utils import ksd as ref_ksd       # /root/reference/code/src/utils/ksd.py
    assert ref_thinning.__file__.startswith(REF) and ref_ksd.__file__.startswith(REF)
    yield ref_thinning, ref_ksd
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_reference_glue_runs_unchanged_on_this_package(reference_modules, g45):
    ref_thinning, ref_ksd = reference_modules
    idx = ref_thinning.gaussian_thin(g45['sample'], g45['log_p'], g45['sample_mean'], g45['sample_cov'], 50)
    assert idx.dtype == np.uint32 and np.array_equal(idx, g45['glue_gaussian_thin_50'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        idx = ref_thinning.gaussian_thin(g45['sample'], g45['log_p'], g45['laplace_mean'], g45['laplace_cov'], 30)
    assert np.array_equal(idx, g45['glue_gaussian_thin_laplace_30'])
    ks = ref_ksd.calculate_ksd(g45['sample'], g45['gradient'], g45['glue_idx_for_ksd'])
    np.testing.assert_allclose(ks, g45['glue_calculate_ksd'], rtol=1e-10, atol=0)


def test_reference_unit_test_passes_on_this_package(reference_modules):
    spec = importlib.util.spec_from_file_location('reference_test_ksd', os.path.join(REF, 'tests', 'test_ksd.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)           # imports stein_thinning.stein.kmat and utils.ksd.reindex_integrand
    mod.test_reindex_integrand()
