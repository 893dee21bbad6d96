"""One emulated rank of a FAILING multi-rank thinning job (TEST INFRASTRUCTURE ONLY; see simt_rank_worker.py).

usage: simt_fault_worker.py SCENARIO STRATEGY RANK WORLD RENDEZVOUS_DIR
  SCENARIO  absent : the last rank opens its mailbox and then never launches the job
            dies   : the last rank leaves after 17 of 40 greedy steps (it runs the job with m = 17)
            forged : the last rank is absent and its slot of exchange 0 holds a record with the right tag and an
                     index past the end of the sample (a stale record of an earlier, larger job)
  STRATEGY  persistent | steps  (single cooperative launch / one launch per greedy step)
Every surviving rank must get GFK_ECUDA from gfk_thin_check within a bounded time; inputs sit in front of PROT_NONE
pages (simt_harness.guarded), so an out-of-range row read would kill the process instead.
"""
import ctypes
import os
import struct
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import simt_harness as H  # noqa: E402
from simt_rank_worker import FAKE_NCCL, _mcmc_like, _publish, _wait_for  # noqa: E402
from gradient_free_mcmc_postprocessing_b200.distributed import shard_bounds  # noqa: E402

M_FULL, M_DIES = 40, 17


def main():
    scenario, strategy = sys.argv[1], sys.argv[2]
    rank, world, rdv = int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    warnings.simplefilter('ignore')
    os.environ.setdefault('SIMT_SMS', '3')
    os.environ['GFK_P2P'] = '1'
    os.environ['GFK_PERSISTENT'] = '1' if strategy == 'persistent' else '0'
    # rank 0 has the short deadline; the others would wait for a minute and must be released by rank 0's poison word
    os.environ['GFK_SPIN_MS'] = '1500' if rank == 0 else '60000'
    os.environ['GFK_SPIN_FIRST_MS'] = os.environ['GFK_SPIN_MS']
    lib = H.load()
    uid_path = os.path.join(rdv, 'uid')
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        H.check(lib.gfk_comm_unique_id(buf, FAKE_NCCL))
        _publish(uid_path, buf.raw)
    H.check(lib.gfk_comm_init(world, rank, _wait_for(uid_path), FAKE_NCCL))
    hbuf = (ctypes.c_char * 64)()
    H.check(lib.gfk_p2p_local_handle(hbuf))
    _publish(os.path.join(rdv, 'handle%d' % rank), hbuf.raw)
    handles = b''.join(_wait_for(os.path.join(rdv, 'handle%d' % r)) for r in range(world))
    H.check(lib.gfk_p2p_open(handles, world, rank))

    rng = np.random.default_rng(5)
    x, g = _mcmc_like(rng, 2000, 4)
    faulty = world - 1
    p = H.SimProblem(x, g, preconditioner='med')
    bounds = shard_bounds(p.n, world, rank)
    done = os.path.join(rdv, 'done%d' % rank)
    if rank == faulty:
        if scenario == 'dies':
            idx, _ = p.thin(M_DIES, row_range=bounds)         # its own 17 steps complete normally
            assert len(idx) == M_DIES
        # absent / forged: never launch.  Keep the mailbox mapped until the survivors are through.
        for r in range(world - 1):
            _wait_for(os.path.join(rdv, 'done%d' % r), timeout=300.0)
        print('SIMT_FAULT_OK', rank, flush=True)
        return
    if scenario == 'forged' and rank == 0:
        lib.gfk_test_mailbox.restype = ctypes.POINTER(ctypes.c_ulonglong)
        lib.gfk_test_mailbox.argtypes = [ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_int)]
        tag_base, parity = ctypes.c_uint(0), ctypes.c_int(0)
        mail = lib.gfk_test_mailbox(ctypes.byref(tag_base), ctypes.byref(parity))
        tag = (tag_base.value + 0 + 1) << 32                  # exchange 0 of the next job
        vb = struct.unpack('<Q', struct.pack('<d', -1e300))[0]     # it wins every comparison
        off = (((parity.value * 2 + 0) * 16) + faulty) * 4          # csrc/thin.cu mail_offset
        mail[off + 0] = tag | (vb & 0xffffffff)
        mail[off + 1] = tag | (vb >> 32)
        mail[off + 2] = tag | (p.n + 5)                             # past the end of x / s
    t0 = time.time()
    try:
        p.thin(M_FULL, row_range=bounds)
    except H.SimError as e:
        msg = str(e)
    else:
        raise AssertionError('the job of rank %d succeeded although rank %d failed' % (rank, faulty))
    took = time.time() - t0
    assert 'error 2' in msg and 'job aborted on the device' in msg, msg
    if scenario == 'forged' and rank == 0:
        assert 'out of range' in msg, msg
    elif rank == 0:
        assert 'deadline' in msg, msg
    else:
        assert 'peer rank aborted' in msg, msg                      # released by rank 0, not by its own 60 s
    assert took < 40.0, took
    _publish(done, b'1')
    print('SIMT_FAULT_OK', rank, '%.1fs' % took, msg.split('device (')[1].split(')')[0], flush=True)


if __name__ == '__main__':
    main()
