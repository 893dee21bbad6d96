"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard bounds, the (value, index)
exchange reduction with numpy.argmin semantics, KSD tile ownership and the unique-id broadcast."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gradient_free_mcmc_postprocessing_b200.distributed import ksd_tile_owner, pick_winner, shard_bounds


def test_shard_bounds_partition():
    for n in (1, 7, 10, 1000, 10 ** 6 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pick_winner_is_argmin():
    rng = np.random.default_rng(1)
    for _ in range(200):
        v = rng.integers(0, 4, size=12).astype(float)          # many exact ties
        if rng.random() < 0.2:
            v[rng.integers(0, 12)] = np.nan
        spans = [shard_bounds(12, 3, r) for r in range(3)]
        cand_v, cand_i = [], []
        for b, e in spans:
            j = int(np.argmin(v[b:e]))
            cand_v.append(v[b + j]); cand_i.append(b + j)
        assert pick_winner(cand_v, cand_i) == int(np.argmin(v))


def test_ksd_tiles_round_robin():
    owners = [ksd_tile_owner(t, 4) for t in range(16)]
    assert owners == [0, 1, 2, 3] * 4


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, seed):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)                       # same data on every rank (replica)
        objective = rng.integers(0, 50, size=n).astype(float)   # exact ties on purpose
        begin, end = shard_bounds(n, world, rank)
        expect = []
        got = []
        for step in range(20):
            local = int(np.argmin(objective[begin:end]))
            mine = torch.tensor([objective[begin + local], float(begin + local)], dtype=torch.float64)
            gathered = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(gathered, mine)                     # the per-step 16-byte exchange
            vals = [float(g[0]) for g in gathered]
            idxs = [int(g[1]) for g in gathered]
            winner = pick_winner(vals, idxs)
            got.append(winner)
            expect.append(int(np.argmin(objective)))
            objective = objective + rng.integers(0, 5, size=n)  # next "column", identical on all ranks
            objective[winner] += 3
        assert got == expect, (rank, got, expect)
        # the 128-byte id broadcast used by distributed.init (non-NCCL backend branch)
        uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(uid, src=0)
        assert uid.tolist() == list(range(128))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_exchange_world2_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 1001, 7), nprocs=2, join=True)


class _FakeComm:
    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def gfk_comm_size(self):
        return self.world

    def gfk_comm_rank(self):
        return self.rank


def _replicate_worker(rank, world, port):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gradient_free_mcmc_postprocessing_b200._engine import _replicate_from_host
        rng = np.random.default_rng(5)
        for n in (7, 8, 1001):                     # ragged and even splits
            for shape in ((n, 4), (n,)):
                a = rng.standard_normal(shape)
                full = _replicate_from_host(a, torch.device('cpu'), _FakeComm(world, rank))
                assert tuple(full.shape) == shape and full.dtype == torch.float64
                assert np.array_equal(full.numpy(), a), (rank, n, shape)
                t = torch.from_numpy(a.astype(np.float32))          # dtype conversion path
                full = _replicate_from_host(t, torch.device('cpu'), _FakeComm(world, rank))
                assert np.array_equal(full.numpy(), a.astype(np.float32).astype(np.float64))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_host_upload_reassembles_replica_world2_and_3():
    """The shard-upload + all-gather used for host inputs on multi-GPU runs (here on CPU tensors over gloo)."""
    mp.spawn(_replicate_worker, args=(2, _free_port()), nprocs=2, join=True)
    mp.spawn(_replicate_worker, args=(3, _free_port()), nprocs=3, join=True)
