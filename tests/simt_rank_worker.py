"""One emulated rank of a multi-rank run of tests/host/libgfk_sim.so (TEST INFRASTRUCTURE ONLY).

usage: simt_rank_worker.py RANK WORLD RENDEZVOUS_DIR
One process per rank, like one process per GPU: the communicator comes from tests/host/libfake_nccl.so, the peer
mailboxes from the emulator's shared-memory cudaIpc* calls, so the multi-rank instantiations of the thinning kernels
(cross-rank mailbox stage of the single-launch kernel, mailbox exchange of the launch-per-step kernels, all-gather
fallback) and the sharded KSD run exactly the code the GPUs run.
"""
import ctypes
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import simt_harness as H  # noqa: E402
from gradient_free_mcmc_postprocessing_b200.distributed import shard_bounds  # noqa: E402
from oracle.stein_thinning_oracle import stein as os_, thinning as ot  # noqa: E402

FAKE_NCCL = os.path.join(HERE, 'host', 'libfake_nccl.so').encode()


def _wait_for(path, timeout=60.0):
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise TimeoutError(path)
        time.sleep(0.005)
    while True:
        with open(path, 'rb') as f:
            data = f.read()
        if data:
            return data
        time.sleep(0.005)


def _publish(path, data):
    with open(path + '.tmp', 'wb') as f:
        f.write(data)
    os.rename(path + '.tmp', path)


def _mcmc_like(rng, n, d, dup=0.5):
    base = rng.standard_normal((n, d)) @ (np.eye(d) + 0.3 * rng.standard_normal((d, d)))
    keep = rng.random(n) > dup
    keep[0] = True
    src = np.maximum.accumulate(np.where(keep, np.arange(n), 0))
    x = base[src]
    return np.ascontiguousarray(x), np.ascontiguousarray(-(x - 0.1) * (1.0 + 0.2 * np.arange(d)))


def main():
    rank, world, rdv = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    warnings.simplefilter('ignore')
    os.environ.setdefault('SIMT_SMS', '3')
    lib = H.load()
    uid_path = os.path.join(rdv, 'uid')
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        H.check(lib.gfk_comm_unique_id(buf, FAKE_NCCL))
        _publish(uid_path, buf.raw)
    uid = _wait_for(uid_path)
    H.check(lib.gfk_comm_init(world, rank, uid, FAKE_NCCL))
    assert lib.gfk_comm_size() == world and lib.gfk_comm_rank() == rank
    hbuf = (ctypes.c_char * 64)()
    H.check(lib.gfk_p2p_local_handle(hbuf))
    _publish(os.path.join(rdv, 'handle%d' % rank), hbuf.raw)
    handles = b''.join(_wait_for(os.path.join(rdv, 'handle%d' % r)) for r in range(world))
    H.check(lib.gfk_p2p_open(handles, world, rank))
    assert lib.gfk_p2p_enabled() == 1
    us = ctypes.c_double(0.0)
    H.check(lib.gfk_p2p_selftest(200, ctypes.byref(us), None))

    rng = np.random.default_rng(31)
    x4, g4 = _mcmc_like(rng, 3011, 4)            # runs of duplicated rows straddle the shard boundaries
    log_p = -0.5 * np.sum(x4 ** 2, axis=1)
    log_q = log_p + 0.5 * np.sin(x4[:, 0])
    expect4 = ot.thin_gf(x4, log_p, log_q, g4, 40, preconditioner='med')
    x2, g2 = _mcmc_like(rng, 1500, 2)
    expect2 = ot.thin(x2, g2, 30)
    x10, g10 = _mcmc_like(rng, 1200, 10)
    expect10 = ot.thin(x10, g10, 20, preconditioner='med')
    # peer mailboxes are opt-in (GFK_P2P=1); the default multi-rank exchange is the NCCL all-gather
    for env in (dict(GFK_P2P='1'), dict(GFK_P2P='1', GFK_PERSISTENT='0'), dict()):
        for k in ('GFK_PERSISTENT', 'GFK_P2P'):
            os.environ.pop(k, None)
        os.environ.update(env)
        for precision in ('f64', 'f64-fast'):
            p = H.SimProblem(x4, g4, log_p, log_q, preconditioner='med', precision=precision)
            idx, _ = p.thin(40, row_range=shard_bounds(p.n, world, rank))
            assert np.array_equal(idx, expect4), (rank, env, precision, idx[:8], expect4[:8])
            p = H.SimProblem(x2, g2, precision=precision)
            idx, _ = p.thin(30, row_range=shard_bounds(p.n, world, rank))
            assert np.array_equal(idx, expect2), (rank, env, precision)
            p = H.SimProblem(x10, g10, preconditioner='med', precision=precision)
            idx, _ = p.thin(20, row_range=shard_bounds(p.n, world, rank))
            assert np.array_equal(idx, expect10), (rank, env, precision)
        # more ranks than rows: a rank without rows offers (+inf, no index) in every exchange
        for nt in (1, 2):
            xt, gt = np.ascontiguousarray(x2[5:5 + nt]), np.ascontiguousarray(g2[5:5 + nt])
            p = H.SimProblem(xt, gt, standardize=False)
            idx, _ = p.thin(6, row_range=shard_bounds(nt, world, rank))
            assert np.array_equal(idx, ot.thin(xt, gt, 6, standardize=False)), (rank, env, nt, idx)
    for k in ('GFK_PERSISTENT', 'GFK_P2P'):
        os.environ.pop(k, None)

    # KSD: row tiles dealt round-robin, one sum all-reduce of the two vectors, replicated scan
    n = 450
    p = H.SimProblem(x4[:n], g4[:n], log_p[:n], log_q[:n], preconditioner='med')
    low, diag = p.ksd_rowsums(part=rank, n_parts=world)
    low_d, diag_d = H.dev(low), H.dev(diag)
    H.check(lib.gfk_allreduce_sum(H.ptr(low_d), n, None))
    H.check(lib.gfk_allreduce_sum(H.ptr(diag_d), n, None))
    ks = H.guarded((n,), np.float64)
    H.check(lib.gfk_ksd_finish(H.ptr(low_d), H.ptr(diag_d), n, H.ptr(ks), None))
    ref = ot._make_stein_gf_integrand(x4[:n], log_p[:n], log_q[:n], g4[:n], preconditioner='med')
    np.testing.assert_allclose(ks, os_.ksd(ref, n), rtol=1e-10)
    H.check(lib.gfk_comm_finalize())
    print('SIMT_RANK_OK', rank, world, flush=True)


if __name__ == '__main__':
    main()
