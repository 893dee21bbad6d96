"""Chain-batched thinning and the ``.npy`` result layer ("next").

The reference thins five independent chains per experiment and keeps inputs and results as ``.npy`` files written by
its ``@cached`` decorator (``code/src/utils/caching.py:72-73,100-101``: ``np.save`` / ``np.load`` of
``<function>_<i>.npy``; ``code/notebooks/lotka_volterra/Stein_thinning.ipynb#c15-#c17``: ``thin`` mapped over chains
with joblib).  Chains are independent, so with a ``torch.distributed`` group they are dealt round-robin to the ranks
with no data-path collective - unlike ``thin`` itself, which shards the ROWS of one chain.
"""
import os

import numpy as np

from .thinning import thin, thin_gf


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


def thin_chains(samples, gradients, n_points, standardize=True, preconditioner='id', *, out_paths=None, rank=0,
                world=1, precision=None, device=None):
    """``[thin(samples[i], gradients[i], n_points, ...) for i in chains]``; entries may be arrays or .npy paths.
    Returns a list with ``None`` for chains owned by other ranks."""
    if len(samples) != len(gradients):
        raise ValueError('samples and gradients must list the same chains')
    out = [None] * len(samples)
    for i in _my_chains(len(samples), rank, world):
        out[i] = thin(load_array(samples[i]), load_array(gradients[i]), n_points, standardize, preconditioner,
                      precision=precision, device=device)
        if out_paths is not None:
            save_indices(out_paths[i], out[i])
    return out


def thin_gf_chains(samples, log_ps, log_qs, gradient_qs, n_points, standardize=True, preconditioner='id',
                   range_cap=None, *, out_paths=None, rank=0, world=1, precision=None, device=None):
    """Gradient-free counterpart of ``thin_chains``."""
    if not (len(samples) == len(log_ps) == len(log_qs) == len(gradient_qs)):
        raise ValueError('all inputs must list the same chains')
    out = [None] * len(samples)
    for i in _my_chains(len(samples), rank, world):
        out[i] = thin_gf(load_array(samples[i]), load_array(log_ps[i]), load_array(log_qs[i]),
                         load_array(gradient_qs[i]), n_points, standardize, preconditioner, range_cap,
                         precision=precision, device=device)
        if out_paths is not None:
            save_indices(out_paths[i], out[i])
    return out
