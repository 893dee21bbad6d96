"""B200-native Stein thinning and gradient-free kernel Stein discrepancy.

Drop-in for the hot path of aglebov/gradient-free-mcmc-postprocessing: the modules mirror the
third-party ``stein_thinning`` package the reference imports (``thinning``, ``kernel``, ``stein``)
and the reference's own wrappers (``thinning.gaussian_thin``, ``utils.ksd``).  All arithmetic
runs in ``libgfk.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/gfk.h``).

    from gradient_free_mcmc_postprocessing_b200.thinning import thin, thin_gf, gaussian_thin
    from gradient_free_mcmc_postprocessing_b200.stein import ksd, kmat
    from gradient_free_mcmc_postprocessing_b200.utils.ksd import calculate_ksd

``install_as_stein_thinning()`` registers these modules under the name ``stein_thinning`` so
existing reference code (``from stein_thinning.thinning import thin_gf``) runs unchanged.
"""
import sys

from . import _lib  # noqa: F401  (does not load the library until first use)

__all__ = ['install_as_stein_thinning']


def install_as_stein_thinning():
    """Alias this package's modules as ``stein_thinning[.thinning|.kernel|.stein]``."""
    import types
    from . import kernel, stein, thinning
    pkg = types.ModuleType('stein_thinning')
    pkg.thinning, pkg.kernel, pkg.stein = thinning, kernel, stein
    pkg.__path__ = []
    sys.modules['stein_thinning'] = pkg
    sys.modules['stein_thinning.thinning'] = thinning
    sys.modules['stein_thinning.kernel'] = kernel
    sys.modules['stein_thinning.stein'] = stein
    return pkg
