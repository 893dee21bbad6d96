"""A/B timing of libgfk builds (GFK_LIB=<path>) or switches (GFK_TMEM=1, GFK_EXP_DD=0, ...): C1/C2/C3 thinning
us/step.  Development tool."""
import os, sys, time, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradient_free_mcmc_postprocessing_b200 import synthetic
from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
warnings.simplefilter('ignore')
tag = os.environ.get('AB_TAG', 'head')
for cname in ('c1', 'c2', 'c3'):
    c = synthetic.CONFIGS[cname]()
    for prec in ('f64', 'f64-fast', 'f32'):
        if c['kind'] == 'gf':
            p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med', precision=prec)
        else:
            p = SteinProblem(c['sample'], c['gradient'], preconditioner='med', precision=prec)
        ts = []
        for rep in range(4):
            t = {}
            p.thin(c['m'], timing=t)
            ts.append(t['loop_ms'])
        print(tag, cname, prec, 'us/step(event)', round(min(ts[1:]) / c['m'] * 1e3, 3), flush=True)
if os.environ.get('AB_KSD'):
    c = synthetic.config_c4(n=50_000)
    for prec in ('f64-fast', 'f32'):
        p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med', precision=prec)
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter(); p.ksd(None, 50_000); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print(tag, 'ksd50k d=10', prec, 'Gpairs/s', round(50_000 * 50_001 / 2 / best / 1e9, 2), flush=True)
