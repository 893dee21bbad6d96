"""Runs the EMULATOR build of the kernels (tests/host/libgfk_sim.so) on full-size BASELINE configs at the B200 launch
geometry against the committed oracle sequences (tests/golden/full_size_indices.npz).  No GPU involved; a prediction of
what the device returns, not a measurement.

    python tools/emulated_full_size_check.py c5 [GFK_STEP_TMA=1 ...]      # c2 | c3 | c5 | chains5, then environment settings
"""
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['SIMT_SMS'] = '148'
for kv in sys.argv[2:]:
    k, v = kv.split('=', 1)
    os.environ[k] = v

import simt_harness as H  # noqa: E402
from gradient_free_mcmc_postprocessing_b200 import synthetic  # noqa: E402

warnings.simplefilter('ignore')
which = sys.argv[1]
if which == 'chains5':
    # five C2-shaped chains of 2 x 10^5 rows in ONE batched launch at the B200 geometry (29 CTAs per chain) against one
    # job per chain (148 CTAs each) - the first 120 greedy steps
    m = 120
    probs = []
    for i in range(5):
        c = synthetic.config_c2(seed=12345 + i, m=m)
        probs.append(H.SimProblem(c['sample'], c['gradient'], preconditioner='med'))
    t = time.time()
    single = [p.thin(m) for p in probs]
    t1 = time.time() - t
    t = time.time()
    got = H.thin_batched(probs, m)
    same = got is not None and all(np.array_equal(got[0][i], single[i][0]) and np.array_equal(got[1][i], single[i][1])
                                   for i in range(5))
    print('chains5 emulated: one job per chain %.1f s, batched %.1f s, indices and objectives identical: %s' %
          (t1, time.time() - t, same), flush=True)
    sys.exit(0 if same else 1)
golden = np.load(os.path.join(ROOT, 'tests', 'golden', 'full_size_indices.npz'))
c = synthetic.CONFIGS[which]()
t = time.time()
if c['kind'] == 'gf':
    p = H.SimProblem(c['sample'], c['gradient_q'], c['log_p'], c['log_q'], preconditioner='med')
else:
    p = H.SimProblem(c['sample'], c['gradient'], preconditioner='med')
idx = p.thin(c['m'])[0]
want = golden[which + '_idx']
same = bool(np.array_equal(idx, want))
print('%s emulated kernels %s: %.1f s, all %d indices identical to the numpy oracle: %s' %
      (which, ' '.join(sys.argv[2:]), time.time() - t, len(want), same), flush=True)
if not same:
    print('   first divergence at step', int(np.flatnonzero(idx != want)[0]))
    sys.exit(1)
