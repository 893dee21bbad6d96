"""One rank of a world-size-N run of the PUBLIC Python API over the emulator build (TEST INFRASTRUCTURE ONLY):
torch.distributed (gloo) for the plumbing, gradient_free_mcmc_postprocessing_b200.distributed.init() for the library
communicator (fake NCCL) and the peer mailboxes (emulated cudaIpc), then thin / thin_gf / ksd exactly as a user calls
them under torchrun.  usage: simt_dist_worker.py RANK WORLD PORT"""
import contextlib
import ctypes
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import simt_harness as H  # noqa: E402
from gradient_free_mcmc_postprocessing_b200 import _engine, _lib, distributed as gdist, proxies  # noqa: E402
from gradient_free_mcmc_postprocessing_b200.stein import ksd  # noqa: E402
from gradient_free_mcmc_postprocessing_b200.thinning import _make_stein_gf_integrand, thin, thin_gf  # noqa: E402
from oracle.stein_thinning_oracle import stein as os_, thinning as ot  # noqa: E402


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    warnings.simplefilter('ignore')
    os.environ.setdefault('SIMT_SMS', '3')
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=port)
    # the same two patches as tests/test_simt_python_layer.py: library handle and "device"
    _lib._lib = H.load()
    cpu = torch.device('cpu')
    for mod in (_engine, proxies):
        mod._require_cuda = lambda device=None: cpu
        mod._stream_ptr = lambda device=None: ctypes.c_void_p(0)
    torch.cuda.device = lambda dev: contextlib.nullcontext()
    gdist.nccl_library_path = lambda: os.path.join(HERE, 'host', 'libfake_nccl.so')
    dist.init_process_group('gloo', rank=rank, world_size=world)
    assert gdist.init(p2p=True) == (world, rank)        # mailboxes are opt-in; the default is covered by the last mode
    assert _lib.load().gfk_p2p_enabled() == 1

    rng = np.random.default_rng(77)
    n, d = 2600, 4
    base = rng.standard_normal((n, d))
    keep = rng.random(n) > 0.5
    keep[0] = True
    x = np.ascontiguousarray(base[np.maximum.accumulate(np.where(keep, np.arange(n), 0))])   # ties across shards
    g = np.ascontiguousarray(-(x - 0.1) * (1.0 + 0.2 * np.arange(d)))
    log_p = -0.5 * np.sum(x ** 2, axis=1)
    log_q = log_p + 0.5 * np.sin(x[:, 0])
    for env in (dict(GFK_P2P='1'), dict(GFK_P2P='1', GFK_PERSISTENT='0'), dict(GFK_P2P='0')):
        for k in ('GFK_PERSISTENT', 'GFK_P2P'):
            os.environ.pop(k, None)
        os.environ.update(env)
        for precision in ('f64', 'f64-fast'):
            assert np.array_equal(thin(x, g, 30, preconditioner='med', precision=precision),
                                  ot.thin(x, g, 30, preconditioner='med')), (rank, env, precision)
            assert np.array_equal(thin_gf(x, log_p, log_q, g, 30, preconditioner='med', precision=precision),
                                  ot.thin_gf(x, log_p, log_q, g, 30, preconditioner='med')), (rank, env, precision)
        dist.barrier()
    for k in ('GFK_PERSISTENT', 'GFK_P2P'):
        os.environ.pop(k, None)
    # independent chains are dealt over the ranks and thinned by the rank-local batched launch - no exchange, although a
    # multi-rank communicator is active (thin() itself would shard ONE chain over all ranks)
    from gradient_free_mcmc_postprocessing_b200.chains import thin_chains
    cs = [(x[i * 300:i * 300 + 900], g[i * 300:i * 300 + 900]) for i in range(5)]
    got = thin_chains([c[0] for c in cs], [c[1] for c in cs], 12, preconditioner='med', rank=rank, world=world)
    for i, c in enumerate(cs):
        if i % world == rank:
            assert np.array_equal(got[i], ot.thin(c[0], c[1], 12, preconditioner='med')), (rank, i)
        else:
            assert got[i] is None
    dist.barrier()
    m = 400
    integ = _make_stein_gf_integrand(x[:m], log_p[:m], log_q[:m], g[:m], preconditioner='med')
    ref = ot._make_stein_gf_integrand(x[:m], log_p[:m], log_q[:m], g[:m], preconditioner='med')
    np.testing.assert_allclose(ksd(integ, m), os_.ksd(ref, m), rtol=1e-10)       # tiles dealt over ranks + all-reduce
    dist.barrier()
    gdist.finalize()
    dist.destroy_process_group()
    print('SIMT_DIST_OK', rank, world, flush=True)


if __name__ == '__main__':
    main()
