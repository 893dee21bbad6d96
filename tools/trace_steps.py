"""Development tool: per-phase SM-clock stamps of CTA 0 / thread 0 in the persistent thinning kernel.
Needs a -DGFK_TRACE build: GFK_LIB=$PWD/libgfk_trace.so python tools/trace_steps.py [c1|c2|c3] [precision]"""
import os, sys, warnings
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gradient_free_mcmc_postprocessing_b200 import synthetic
from gradient_free_mcmc_postprocessing_b200._engine import SteinProblem
warnings.simplefilter('ignore')
multi = 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1
if multi:
    import torch.distributed as dist
    from gradient_free_mcmc_postprocessing_b200 import distributed as gdist
    torch.cuda.set_device(int(os.environ['LOCAL_RANK'])); dist.init_process_group('nccl'); gdist.init()
cname = sys.argv[1] if len(sys.argv) > 1 else 'c3'
prec = sys.argv[2] if len(sys.argv) > 2 else 'f64-fast'
c = synthetic.CONFIGS[cname]()
if c['kind'] == 'gf':
    p = SteinProblem(c['sample'], c['gradient_q'], log_p=c['log_p'], log_q=c['log_q'], preconditioner='med', precision=prec)
else:
    p = SteinProblem(c['sample'], c['gradient'], preconditioner='med', precision=prec)
for rep in range(3):
    t = {}
    p.thin(c['m'], timing=t)
ws = t['ws'].cpu().numpy()
n1, n2 = 128 * 8 * 8, 64 * 160 * 8
pad = int(os.environ.get('TRACE_PAD', '8'))                 # tail padding of struct ThinWs (16-byte aligned)
tr = ws[-(n1 + n2 + pad):-(n2 + pad)].view(np.int64).reshape(128, 8)
cta = ws[-(n2 + pad):-pad].view(np.int64).reshape(64, 160)
names = ['t0->t1 update(thread 0)', 't1->t2 block_argmin+CTA wait', 't2->t3 local exchange', 't3->t4 cross-GPU', 't4->t5 sync out',
         't5->t0 y load + setup']
if not multi or int(os.environ['RANK']) == 0:
    print(cname, prec, 'us/step', t['loop_ms'] / c['m'] * 1e3)
    d = np.diff(tr[10:90, :6], axis=1).astype(float)
    nxt = (tr[11:91, 0] - tr[10:90, 5]).astype(float)
    allp = np.concatenate([d, nxt[:, None]], axis=1) / 1e3     # ns -> us
    for k, nme in enumerate(names):
        print(f'  {nme:30s} mean {allp[:, k].mean():6.2f} us   p90 {np.percentile(allp[:, k], 90):6.2f}')
    print('  total', allp.sum(axis=1).mean())
    ncta = int((cta[5] > 0).sum())
    rel = (cta[4:60, :ncta] - tr[36:92, 0][:, None]) / 1e3     # CTA publish time relative to CTA 0's step start
    print(f'  CTA publish time after step start over {ncta} CTAs: min {rel.min(axis=1).mean():.2f} median '
          f'{np.median(rel, axis=1).mean():.2f} max {rel.max(axis=1).mean():.2f} us; slowest CTA ids', np.bincount(rel.argmax(axis=1)).argsort()[-3:])
if multi:
    gdist.finalize(); dist.destroy_process_group()
