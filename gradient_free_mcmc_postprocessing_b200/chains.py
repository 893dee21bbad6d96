"""Chain-batched thinning and the ``.npy`` result layer ("next").

The reference thins five independent chains per experiment and keeps inputs and results as ``.npy`` files written by
its ``@cached`` decorator (``code/src/utils/caching.py:72-73,100-101``: ``np.save`` / ``np.load`` of
``<function>_<i>.npy``; ``code/notebooks/lotka_volterra/Stein_thinning.ipynb#c15-#c17``: ``thin`` mapped over chains
with joblib).  Chains are independent, so with a ``torch.distributed`` group they are dealt round-robin to the ranks
with no data-path collective - unlike ``thin`` itself, which shards the ROWS of one chain.

The chains a rank owns are thinned in ONE cooperative launch (``gfk_thin_batched``: the grid is split into one group
of CTAs per chain, each with its own workspace and exchange slots), so a set of small chains pays the per-step
latency floor of the greedy search once, not once per chain; sets the batched kernel does not cover (d > 4, a dense
preconditioner, a chain too long for its CTA group) fall back to one job per chain.
"""
import os

import numpy as np

from ._engine import SteinProblem, thin_batched


def load_array(a):
    """ndarray / tensor unchanged; a path is read with ``np.load`` (the reference's cache files are plain .npy)."""
    if isinstance(a, (str, os.PathLike)):
        return np.load(a, allow_pickle=False)
    return a


def save_indices(path, idx):
    """Same on-disk form as the reference's cached index vectors: ``np.save`` of the uint32 array."""
    np.save(path, np.asarray(idx, dtype=np.uint32))
    return path


def _my_chains(n_chains, rank, world):
    return [i for i in range(n_chains) if i % world == rank]


def _thin_set(problems, n_points):
    """Index vectors of the problems: one batched launch when eligible, otherwise one job after the other."""
    from . import _lib
    local_only = _lib.load().gfk_comm_size() > 1     # a multi-rank communicator is active: SteinProblem.thin would shard
    out = thin_batched(problems, n_points) if (len(problems) > 1 or local_only) else None
    if out is None:
        if local_only and problems:
            raise NotImplementedError(
                'these chains are not eligible for the rank-local batched kernel (d > 4, a dense preconditioner or a '
                'chain too long for shared memory) and a multi-rank communicator is active, under which thin() shards '
                'ONE chain over all ranks: thin such chains with thin() / thin_gf(), or finalize the communicator first')
        out = [p.thin(n_points) for p in problems]
    return out


def thin_chains(samples, gradients, n_points, standardize=True, preconditioner='id', *, out_paths=None, rank=0,
                world=1, precision=None, device=None):
    """``[thin(samples[i], gradients[i], n_points, ...) for i in chains]``; entries may be arrays or .npy paths.
    Returns a list with ``None`` for chains owned by other ranks."""
    if len(samples) != len(gradients):
        raise ValueError('samples and gradients must list the same chains')
    out = [None] * len(samples)
    mine = _my_chains(len(samples), rank, world)
    problems = [SteinProblem(load_array(samples[i]), load_array(gradients[i]), standardize=standardize,
                             preconditioner=preconditioner, precision=precision, device=device) for i in mine]
    for i, idx in zip(mine, _thin_set(problems, n_points)):
        out[i] = idx
        if out_paths is not None:
            save_indices(out_paths[i], idx)
    return out


def thin_gf_chains(samples, log_ps, log_qs, gradient_qs, n_points, standardize=True, preconditioner='id',
                   range_cap=None, *, out_paths=None, rank=0, world=1, precision=None, device=None):
    """Gradient-free counterpart of ``thin_chains``."""
    if not (len(samples) == len(log_ps) == len(log_qs) == len(gradient_qs)):
        raise ValueError('all inputs must list the same chains')
    out = [None] * len(samples)
    mine = _my_chains(len(samples), rank, world)
    problems = [SteinProblem(load_array(samples[i]), load_array(gradient_qs[i]), log_p=load_array(log_ps[i]),
                             log_q=load_array(log_qs[i]), standardize=standardize, preconditioner=preconditioner,
                             range_cap=range_cap, precision=precision, device=device) for i in mine]
    for i, idx in zip(mine, _thin_set(problems, n_points)):
        out[i] = idx
        if out_paths is not None:
            save_indices(out_paths[i], idx)
    return out
