"""Full-length index sequences of BASELINE configs C2, C3 and C5 from the numpy ORACLE (oracle/stein_thinning_oracle),
for the `-m gpu` tests to compare all m selections with (the oracle needs minutes at these sizes, so it cannot run
inside the tests).  Also runs the emulator build of the kernels on the same inputs at the B200 launch geometry and
reports whether - and where - the two sequences part.

    python tests/golden/make_golden_full.py        # writes tests/golden/full_size_indices.npz  (~8 KB)
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ['SIMT_SMS'] = '148'

import simt_harness as H  # noqa: E402
from gradient_free_mcmc_postprocessing_b200 import synthetic  # noqa: E402
from oracle.stein_thinning_oracle import thinning as ot  # noqa: E402

warnings.simplefilter('ignore')
out = {}

c = synthetic.config_c2()
t = time.time()
out['c2_idx'] = ot.thin(c['sample'], c['gradient'], c['m'], preconditioner='med')
print('C2 oracle: %.1f s, %d unique of %d' % (time.time() - t, len(np.unique(out['c2_idx'])), c['m']), flush=True)
t = time.time()
sim = H.SimProblem(c['sample'], c['gradient'], preconditioner='med').thin(c['m'])[0]
print('C2 emulated kernels: %.1f s, identical: %s' % (time.time() - t, np.array_equal(sim, out['c2_idx'])), flush=True)
if not np.array_equal(sim, out['c2_idx']):
    print('   first divergence at step', int(np.flatnonzero(sim != out['c2_idx'])[0]))
out['c2_idx_kernels'] = sim

c = synthetic.config_c3()
t = time.time()
out['c3_idx'] = ot.thin_gf(c['sample'], c['log_p'], c['log_q'], c['gradient_q'], c['m'], preconditioner='med')
print('C3 oracle: %.1f s, %d unique of %d' % (time.time() - t, len(np.unique(out['c3_idx'])), c['m']), flush=True)
t = time.time()
sim = H.SimProblem(c['sample'], c['gradient_q'], c['log_p'], c['log_q'], preconditioner='med').thin(c['m'])[0]
print('C3 emulated kernels: %.1f s, identical: %s' % (time.time() - t, np.array_equal(sim, out['c3_idx'])), flush=True)
if not np.array_equal(sim, out['c3_idx']):
    print('   first divergence at step', int(np.flatnonzero(sim != out['c3_idx'])[0]))
out['c3_idx_kernels'] = sim
c = synthetic.config_c5()          # d = 64: the oracle needs ~20 minutes for the 500 steps
t = time.time()
out['c5_idx'] = ot.thin(c['sample'], c['gradient'], c['m'], preconditioner='med')
print('C5 oracle: %.1f s, %d unique of %d' % (time.time() - t, len(np.unique(out['c5_idx'])), c['m']), flush=True)
t = time.time()
sim = H.SimProblem(c['sample'], c['gradient'], preconditioner='med').thin(c['m'])[0]
print('C5 emulated kernels: %.1f s, identical: %s' % (time.time() - t, np.array_equal(sim, out['c5_idx'])), flush=True)
out['c5_idx_kernels'] = sim
np.savez_compressed(os.path.join(HERE, 'full_size_indices.npz'), **out)
print('wrote', os.path.join(HERE, 'full_size_indices.npz'))
