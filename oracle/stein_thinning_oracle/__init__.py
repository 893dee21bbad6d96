"""numpy restatement of the un-vendored ``stein_thinning`` library (FP64).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Why a restatement: the reference repo imports all hot-path arithmetic from the
third-party package ``stein_thinning`` (``/root/reference/pyproject.toml:31``,
unpinned; used at ``code/src/thinning.py:5,17`` and
``code/src/utils/ksd.py:5-6,25,27``).  That package is neither vendored under
``/root/reference`` nor installable here (no network).  The reference
notebooks restate its functions inline and print golden outputs from seeded
inputs; this oracle follows those restatements expression for expression and
is pinned by the golden vectors in ``tests/golden/`` (G1-G5 of SURVEY.md
section 8c).

Parity status:
  * pinned by reference golden vectors: ``vfk0_imq``, ``make_precon('id'|'med')``
    for n <= 1000, ``_validate_and_standardize``, ``_make_stein_integrand``,
    ``_make_stein_gf_integrand`` (range_cap=None), ``_greedy_search``, ``thin``,
    ``thin_gf``; index contract of ``kmat`` / ``reindex_integrand``
    (``code/tests/test_ksd.py``).
  * PARITY UNPINNED (no reference test, fixture or regenerable output
    constrains them; the oracle IS the definition): the ``'med'`` subsample
    rule for n > 1000, ``range_cap`` semantics, numeric values of ``ksd()``,
    and the ``'sclmed'`` / ``'smpcov'`` / numeric-string preconditioners.
"""
from . import kernel, stein, thinning  # noqa: F401
