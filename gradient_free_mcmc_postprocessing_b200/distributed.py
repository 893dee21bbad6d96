"""Multi-GPU plumbing: one process per GPU, candidates (rows) block-partitioned across ranks.

``init()`` builds libgfk's NCCL communicator from an existing ``torch.distributed`` process group
(any backend: the 128-byte NCCL unique id is broadcast as a byte tensor).  After it, ``thin`` /
``thin_gf`` shard the candidate rows and exchange one (value, index) pair per rank per greedy step,
and ``ksd`` deals the row tiles round-robin and all-reduces the row sums.
"""
import ctypes
import glob
import os

import numpy as np
import torch

from . import _lib


def shard_bounds(n, n_ranks, rank):
    """Rows ``[begin, end)`` owned by ``rank``: contiguous blocks, sizes differing by at most one."""
    base, extra = divmod(int(n), int(n_ranks))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def pick_winner(values, indices):
    """Host mirror of the device-side exchange reduction: numpy.argmin semantics over per-rank
    (value, global index) candidates - lowest value, NaN first, ties to the lowest index."""
    values = np.asarray(values, dtype=float)
    indices = np.asarray(indices, dtype=np.int64)
    nan = np.isnan(values)
    if nan.any():
        return int(indices[nan].min())
    best = values.min()
    return int(indices[values == best].min())


def ksd_tile_owner(tile, n_parts):
    """Rank that computes 64-row tile ``tile`` of the KSD triangle (round-robin)."""
    return int(tile) % int(n_parts)


def nccl_library_path():
    """The libnccl.so.2 torch itself uses (bundled nvidia-nccl wheel), if present."""
    for root in {os.path.dirname(os.path.dirname(torch.__file__))}:
        hits = glob.glob(os.path.join(root, 'nvidia', 'nccl', 'lib', 'libnccl.so.2'))
        if hits:
            return hits[0]
    return None


def init(group=None, p2p=None):
    """Collective.  Creates the libgfk communicator mirroring the torch.distributed group.

    The per-step exchange of the greedy search defaults to a 16-byte ``ncclAllGather``.  The peer-memory mailboxes
    (CUDA IPC over NVLink, exchange fused into the thinning kernels) are opt-in: ``p2p=True`` or ``GFK_P2P=1`` in
    the environment of EVERY rank.  A failed mailbox exchange ends in ``GfkError`` on all ranks within
    ``GFK_SPIN_MS`` (default 200 ms; csrc/thin.cu "failure protocol"), never in a hang."""
    import torch.distributed as dist
    lib = _lib.load()
    if not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised')
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    path = nccl_library_path()
    path_b = path.encode() if path else None
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0 and world > 1:
        buf = (ctypes.c_char * 128)()
        _lib.check(lib.gfk_comm_unique_id(buf, path_b))
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if world > 1:
        backend = dist.get_backend(group)
        if backend == 'nccl':
            dev_uid = uid.cuda()
            dist.broadcast(dev_uid, src=0, group=group)
            uid = dev_uid.cpu()
        else:
            dist.broadcast(uid, src=0, group=group)
    raw = uid.numpy().tobytes()
    _lib.check(lib.gfk_comm_init(world, rank, raw, path_b))
    if p2p is None:
        p2p = os.environ.get('GFK_P2P', '0') == '1'
    if world > 1 and p2p:
        os.environ['GFK_P2P'] = '1'              # the library reads it per job
        _open_mailboxes(group, world, rank)
    else:
        os.environ['GFK_P2P'] = '0'
    return world, rank


def _open_mailboxes(group, world, rank):
    """Exchange CUDA IPC handles of the per-rank mailboxes so the thinning kernels can store their
    per-step (value, index) candidates straight into peer memory over NVLink.  Collective."""
    import warnings
    import torch.distributed as dist
    lib = _lib.load()
    buf = (ctypes.c_char * 64)()
    rc = lib.gfk_p2p_local_handle(buf)
    mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32)
    nccl = dist.get_backend(group) == 'nccl'
    if nccl:
        mine, ok = mine.cuda(), ok.cuda()
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        handles = b''.join(g.cpu().numpy().tobytes() for g in gathered)
        rc = lib.gfk_p2p_open(handles, world, rank)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32)
        if nccl:
            ok = ok.cuda()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) != 1:
        # every rank must agree, otherwise some would wait on mailboxes nobody writes
        os.environ['GFK_P2P'] = '0'
        warnings.warn('CUDA IPC mailboxes unavailable (%s); using the NCCL all-gather exchange'
                      % lib.gfk_last_error().decode('utf-8', 'replace'))


def finalize():
    _lib.check(_lib.load().gfk_comm_finalize())
