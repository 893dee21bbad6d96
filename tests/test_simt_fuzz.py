"""Property test: for arbitrary small problems the emulated kernels select exactly the oracle's indices (strict and
FMA formulations, every launch strategy, any number of concurrently running CTAs).  Derandomised, so the CPU suite
is reproducible; raise max_examples locally for a longer hunt."""
import os
import warnings

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import simt_harness as H
from oracle.stein_thinning_oracle import stein as os_, thinning as ot

STRATEGY_ENVS = [{}, {'GFK_RESIDENT': '0'}, {'GFK_PERSISTENT': '0'}]


@settings(max_examples=int(os.environ.get('GFK_FUZZ_EXAMPLES', '40')), derandomize='GFK_FUZZ_EXAMPLES' not in os.environ, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(2, 1400), d=st.integers(1, 12), m=st.integers(1, 24),
       dup=st.sampled_from([0.0, 0.3, 0.8]), pre=st.sampled_from(['id', 'med', 'sclmed', '0.7']), gf=st.booleans(),
       strategy=st.integers(0, 2), sms=st.integers(1, 6), fast=st.booleans(), scale=st.sampled_from([1e-3, 1.0, 50.0]))
def test_thinning_matches_oracle(seed, n, d, m, dup, pre, gf, strategy, sms, fast, scale):
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n, d)) * scale + rng.standard_normal(d)
    keep = rng.random(n) > dup
    keep[0] = True
    x = np.ascontiguousarray(base[np.maximum.accumulate(np.where(keep, np.arange(n), 0))])
    g = np.ascontiguousarray(-(x - x.mean(0)) / scale ** 2 * (1.0 + 0.2 * np.arange(d)))
    if np.any(np.mean(np.abs(x - x.mean(0)), axis=0) == 0) or len(np.unique(x, axis=0)) < 2:
        return                                                   # degenerate sample: both sides raise
    log_p = -0.5 * np.sum(((x - x.mean(0)) / scale) ** 2, axis=1)
    log_q = log_p + 0.5 * np.sin(x[:, 0] / scale)
    for k in ('GFK_PERSISTENT', 'GFK_RESIDENT'):
        os.environ.pop(k, None)
    os.environ.update(STRATEGY_ENVS[strategy])
    os.environ['SIMT_SMS'] = str(sms)
    H.set_schedule(False)
    precision = 'f64-fast' if fast else 'f64'
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            try:
                expect = (ot.thin_gf(x, log_p, log_q, g, m, preconditioner=pre) if gf
                          else ot.thin(x, g, m, preconditioner=pre))
            except Exception as e:                               # zero median: the reference refuses the sample
                assert 'Too few unique samples' in str(e)
                return
            p = (H.SimProblem(x, g, log_p, log_q, preconditioner=pre, precision=precision) if gf
                 else H.SimProblem(x, g, preconditioner=pre, precision=precision))
        idx, _ = p.thin(m)
        if not np.array_equal(idx, expect):
            # Classify the first divergence ("index-exact parity").  The strict mode may part from
            # numpy only where the best and the runner-up objective agree to a few ulp - numpy's own pow is not
            # correctly rounded, so such ties are machine-dependent in the reference itself
            # (mirror-symmetric samples, e.g. n = 2, produce them); the FMA formulation where they agree to 1e-11.
            t = int(np.flatnonzero(idx != expect)[0])
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                integ = (ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner=pre) if gf
                         else ot._make_stein_integrand(x, g, preconditioner=pre))
            a = integ(slice(None), slice(None)).copy()
            for j in expect[:t]:
                a += 2 * integ(slice(None), [j])
            gap = abs(a[idx[t]] - a[expect[t]])
            tol = 1e-11 if fast else 16 * np.finfo(float).eps
            assert gap <= tol * max(1.0, np.max(np.abs(a))), ('not a near-tie', precision, t, idx[t], expect[t], gap)
            return
        if n <= 300:
            ref = (ot._make_stein_gf_integrand(x, log_p, log_q, g, preconditioner=pre) if gf
                   else ot._make_stein_integrand(x, g, preconditioner=pre))
            np.testing.assert_allclose(p.ksd(), os_.ksd(ref, n), rtol=1e-10)
    finally:
        for k in ('GFK_PERSISTENT', 'GFK_RESIDENT', 'SIMT_SMS'):
            os.environ.pop(k, None)
