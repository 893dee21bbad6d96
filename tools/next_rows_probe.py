"""Timings of everything that has not been measured on a GPU yet (round 1 ended with GPU access closed):
full-size C4 KSD, the strict-mode KSD with pre-exponentiated weights, the section-8f kernels (proxies, KDE, energy
distance).  Writes gpurun_out/next_rows_probe.json.  Development tool; run under gpurun on one B200."""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradient_free_mcmc_postprocessing_b200 import proxies, synthetic  # noqa: E402
from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem  # noqa: E402

warnings.simplefilter('ignore')
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def note(key, **kw):
    out[key] = kw
    print(key, kw, flush=True)


# C4: full O(n^2) gradient-free KSD, d = 10, 'med' (BASELINE.json configs[3])
for n in (50_000, 200_000):
    c = synthetic.config_c4(n=n)
    for prec in ('f64', 'f64-fast', 'f32'):
        p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med',
                         precision=prec)
        pairs = n * (n + 1) / 2
        ks = p.ksd(None, min(n, 20_000))
        t = timed(lambda: p.ksd(None, n), reps=2)
        note(f'ksd_c4:{n}:{prec}', seconds=t, pair_evals_per_s=pairs / t,
             fp64_tflops_algorithmic=pairs * 99 / t / 1e12, ks_last=float(ks[-1]))
    del p

# C5-shaped step kernel (d = 64): slab kernel vs the opt-in TMA-fed kernel (cp.async.bulk.tensor + mbarrier ring)
c = synthetic.config_c5(n=500_000, m=40)
for prec in ('f64', 'f64-fast', 'f32'):
    for pipe in ('0', '1'):
        os.environ['GFK_STEP_TMA'] = pipe
        p = SteinProblem(c['sample'], c['gradient'], preconditioner='med', precision=prec)
        idx = p.thin(40)
        t = timed(lambda: p.thin(40), reps=2)
        note(f'c5_step:{prec}:pipe{pipe}', us_per_step=t / 40 * 1e6, hbm_floor_us=500_000 * (1040 if prec != 'f32' else 528) / 6547.2e3,
             idx_head=[int(v) for v in idx[:6]])
    os.environ.pop('GFK_STEP_TMA', None)
del p

# large-d KSD in the float32 mode: SIMT tile kernel vs the tensor-core kernel (opt-in GFK_KSD_MMA=1), d = 64
rng5 = np.random.default_rng(5)
for n5 in (50_000, 200_000):
    x5 = torch.from_numpy(rng5.standard_normal((n5, 64)) + 1.0).cuda()
    g5 = -(x5 - 1.0) / 8.0
    p5 = SteinProblem(x5, g5, preconditioner='med', precision='f32')
    pairs5 = n5 * (n5 + 1) / 2
    for mma in ('0', '1'):
        os.environ['GFK_KSD_MMA'] = mma
        ks5 = p5.ksd(None, 10_000)
        t = timed(lambda: p5.ksd(None, n5), reps=2)
        note(f'ksd_d64_f32:{n5}:mma{mma}', seconds=t, pair_evals_per_s=pairs5 / t, ks_10k=float(ks5[-1]))
    os.environ.pop('GFK_KSD_MMA', None)
    del p5

# section 8f kernels
rng = np.random.default_rng(0)
for n, d in ((100_000, 4), (200_000, 2)):
    x = torch.from_numpy(rng.standard_normal((n, d))).cuda()
    t = timed(lambda: proxies.kde_logpdf_and_score(x, return_device=True), reps=2)
    note(f'kde:{n}:{d}', seconds=t, pairs_per_s=n * n / t)
xv = torch.from_numpy(rng.standard_normal((50_000, 4))).cuda()
yv = torch.from_numpy(rng.standard_normal((10_000, 4))).cuda()
t = timed(lambda: proxies.cumulative_energy_distance(xv, yv), reps=2)
note('energy:50000x10000:4', seconds=t, pairs_per_s=(50_000 ** 2 / 2 + 50_000 * 10_000 + 10_000 ** 2 / 2) / t)
x = torch.from_numpy(rng.standard_normal((1_000_000, 4))).cuda()
cov = np.eye(4) + 0.1
t = timed(lambda: proxies._t_proxy_device(x, np.zeros(4), cov, 9.0))
note('student_t:1e6:4', seconds=t, gb_per_s=1_000_000 * (32 + 8 + 32) / t / 1e9)
# five chains per call (lotka_volterra/Stein_thinning.ipynb#c17): one batched cooperative launch vs one job per chain
from gradient_free_mcmc_postprocessing_b200._engine import thin_batched  # noqa: E402
for n_chain, m_chain in ((40_000, 1000), (200_000, 1000), (500_000, 1000)):
    cs = [synthetic.config_c2(n=n_chain, m=m_chain, seed=12345 + i) for i in range(5)]
    probs = [SteinProblem(torch.from_numpy(c['sample']).cuda(), torch.from_numpy(c['gradient']).cuda(), preconditioner='med')
             for c in cs]
    seq = timed(lambda: [p.thin(m_chain) for p in probs], reps=2)
    bat_idx = thin_batched(probs, m_chain)
    if bat_idx is None:
        note(f'chains5:{n_chain}', sequential_s=seq, batched='not eligible')
        continue
    same = all(np.array_equal(a, p.thin(m_chain)) for a, p in zip(bat_idx, probs))
    bat = timed(lambda: thin_batched(probs, m_chain), reps=2)
    note(f'chains5:{n_chain}', sequential_s=seq, batched_s=bat, speedup=seq / bat, same_indices=bool(same))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'next_rows_probe.json'), 'w'), indent=1)
