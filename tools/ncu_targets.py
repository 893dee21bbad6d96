"""Small single-purpose workloads for `ncu -k regex:<kernel>` captures (tools/gpu_first_contact.sh).
usage: ncu_targets.py ksd | c5 | next"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradient_free_mcmc_postprocessing_b200 import proxies, synthetic  # noqa: E402
from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem  # noqa: E402

warnings.simplefilter('ignore')
what = sys.argv[1]
if what == 'ksd':
    c = synthetic.config_c4(n=30_000)
    for prec in ('f64-fast', 'f32'):
        p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med',
                         precision=prec)
        p.ksd(None, 30_000)
elif what == 'c5':
    c = synthetic.config_c5(n=500_000, m=12)
    SteinProblem(c['sample'], c['gradient'], preconditioner='med', precision='f64').thin(12)
elif what == 'next':
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((60_000, 4))).cuda()
    proxies.kde_logpdf_and_score(x, return_device=True)
    proxies.cumulative_energy_distance(x, x[:5000].contiguous())
torch.cuda.synchronize()
