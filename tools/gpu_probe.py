"""First-contact probe for a B200 box: device info, measured FP64/FP32 FMA peaks, golden parity in
every mode with diagnostics, and timings of the config shapes.  Writes gpurun_out/probe.json."""
import ctypes
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradient_free_mcmc_postprocessing_b200 import _lib, synthetic  # noqa: E402
from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem  # noqa: E402

out = {}
lib = _lib.load()
cc, sms = ctypes.c_int(), ctypes.c_int()
name = ctypes.create_string_buffer(128)
_lib.check(lib.gfk_device_info(ctypes.byref(cc), ctypes.byref(sms), name, 128))
out['device'] = dict(name=name.value.decode(), cc=cc.value, sms=sms.value)
for label, f32 in (('fp64_fma_tflops', 0), ('fp32_fma_tflops', 1)):
    v = ctypes.c_double()
    _lib.check(lib.gfk_fma_peak(f32, ctypes.byref(v), None))
    out[label] = v.value
print(json.dumps(out), flush=True)

g = np.load(os.path.join(ROOT, 'tests/golden/g1_g3_bivariate_gaussian.npz'))
for prec in ('f64', 'f64-fast', 'f32'):
    p = SteinProblem(g['sample'], g['gradient'], precision=prec)
    idx = p.thin(20)
    p2 = SteinProblem(g['sample'], g['gradient_q'], log_p=g['log_p'], log_q=g['log_q'], precision=prec)
    idx2 = p2.thin(20)
    print(prec, 'G1', np.array_equal(idx, g['idx_thin']), 'G3', np.array_equal(idx2, g['idx_thin_gf']),
          idx[:8], idx2[:8], flush=True)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts)


warnings.simplefilter('ignore')
timings = {}
variants = [('persist', {}), ('multilaunch', {'GFK_PERSISTENT': '0'})]
for cname, var in [(c, v) for c in ('c1', 'c2', 'c3', 'c5') for v in variants]:
    if cname == 'c5' and var[0] != 'multilaunch':
        continue
    for k in ('GFK_RESIDENT', 'GFK_PERSISTENT', 'GFK_PREFETCH'):
        os.environ.pop(k, None)
    os.environ.update(var[1])
    c = synthetic.CONFIGS[cname]()
    n, d = c['sample'].shape
    for prec in ('f64', 'f64-fast', 'f32'):
        t0 = time.perf_counter()
        if c['kind'] == 'gf':
            p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'],
                             preconditioner='med', precision=prec)
        else:
            p = SteinProblem(c['sample'], c['gradient'], preconditioner='med', precision=prec)
        torch.cuda.synchronize()
        t_setup = time.perf_counter() - t0
        t = timed(lambda: p.thin(c['m']))
        key = f'{cname}:{var[0]}:{prec}'
        timings[key] = dict(n=n, d=d, m=c['m'], setup_s=t_setup, thin_s=t,
                            pair_evals_per_s=n * c['m'] / t, us_per_step=t / c['m'] * 1e6)
        print(key, timings[key], flush=True)
        del p
out['thin'] = timings
for k in ('GFK_RESIDENT', 'GFK_PERSISTENT', 'GFK_PREFETCH'):
    os.environ.pop(k, None)
c = synthetic.config_c4(n=50_000)
for prec in ('f64-fast', 'f32'):
    p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med',
                     precision=prec)
    t = timed(lambda: p.ksd(None, 50_000), reps=2)
    pairs = 50_000 * 50_001 / 2
    out[f'ksd_c4_50k:{prec}'] = dict(seconds=t, pair_evals_per_s=pairs / t)
    print('ksd', prec, out[f'ksd_c4_50k:{prec}'], flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'probe.json'), 'w'), indent=1)
