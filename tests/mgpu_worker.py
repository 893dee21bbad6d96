"""Worker for tests/test_multigpu.py: run under torchrun with >= 2 GPUs (one rank per GPU, NCCL)."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gradient_free_mcmc_postprocessing_b200 import distributed as gdist, synthetic  # noqa: E402
from gradient_free_mcmc_postprocessing_b200.stein import ksd  # noqa: E402
from gradient_free_mcmc_postprocessing_b200.thinning import _make_stein_gf_integrand, thin, thin_gf  # noqa: E402
from oracle.stein_thinning_oracle import stein as os_, thinning as ot  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
    dist.init_process_group('nccl')
    # the NCCL all-gather exchange is the default; the peer-mailbox modes run only when asked for (GFK_TEST_MAILBOX=1,
    # tools/r02_contact_n2.sh) - they are opt-in in the product too
    gdist.init(p2p=os.environ.get('GFK_TEST_MAILBOX') == '1')
    warnings.simplefilter('ignore')
    from gradient_free_mcmc_postprocessing_b200 import _lib
    p2p = bool(_lib.load().gfk_p2p_enabled())
    if rank == 0:
        print('p2p mailboxes:', p2p, flush=True)
    # duplicated rows -> exact ties that straddle shard boundaries
    c = synthetic.config_c2(n=30_011, m=80)
    expect = ot.thin(c['sample'], c['gradient'], 80, preconditioner='med')
    c1 = synthetic.config_c1()
    expect1 = ot.thin_gf(c1['sample'], c1['log_p'], c1['log_q'], c1['gradient_q'], 100, preconditioner='med')
    rng = np.random.default_rng(3)
    x10 = rng.standard_normal((20_000, 10))
    g10 = -x10 * (1 + 0.1 * np.arange(10))
    expect10 = ot.thin(x10, g10, 40, preconditioner='med')
    # exchange through peer mailboxes (persistent + multi-launch kernels) and through NCCL all-gather
    modes = [dict(GFK_P2P='1'), dict(GFK_P2P='1', GFK_PERSISTENT='0'), dict(GFK_P2P='0')] if p2p else [dict(GFK_P2P='0')]
    for env in modes:
        for k in ('GFK_PERSISTENT', 'GFK_P2P'):
            os.environ.pop(k, None)
        os.environ.update(env)
        for precision in ('f64', 'f64-fast'):
            got = thin(c['sample'], c['gradient'], 80, preconditioner='med', precision=precision)
            assert np.array_equal(got, expect), (rank, env, precision, got[:10], expect[:10])
            got = thin_gf(c1['sample'], c1['log_p'], c1['log_q'], c1['gradient_q'], 100, preconditioner='med',
                          precision=precision)
            assert np.array_equal(got, expect1), (rank, env, precision)
            got10 = thin(x10, g10, 40, preconditioner='med', precision=precision)
            assert np.array_equal(got10, expect10), (rank, env, precision)
        dist.barrier()
        if rank == 0:
            print('mode ok:', env, flush=True)
    os.environ.pop('GFK_PERSISTENT', None)
    os.environ['GFK_P2P'] = '0'
    got = expect1
    # every rank returns the same sequence
    t = torch.from_numpy(got.astype(np.int64)).cuda()
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert all(torch.equal(g, t) for g in gathered)
    # sharded KSD triangle + all-reduce
    n = 3000
    integ = _make_stein_gf_integrand(c1['sample'][:n], c1['log_p'][:n], c1['log_q'][:n], c1['gradient_q'][:n],
                                     preconditioner='med')
    ref = ot._make_stein_gf_integrand(c1['sample'][:n], c1['log_p'][:n], c1['log_q'][:n], c1['gradient_q'][:n],
                                      preconditioner='med')
    np.testing.assert_allclose(ksd(integ, n), os_.ksd(ref, n), rtol=1e-10)
    dist.barrier()
    gdist.finalize()
    dist.destroy_process_group()
    if rank == 0:
        print('MGPU_OK world', world)


if __name__ == '__main__':
    main()
