"""Device-resident Stein problem: host-side orchestration above the C ABI.

PyTorch is used only to own device buffers and streams; every arithmetic step of the path
(standardisation, median heuristic, log-weights, pair kernels, greedy search, KSD sums)
is a libgfk call.  Nothing here falls back to the CPU.
"""
import ctypes
import os
import warnings

import numpy as np
import torch

from . import _lib
from ._lib import Problem, check

WEIGHT_SCALE_THRESHOLD = 10     # stein_thinning.thinning.WEIGHT_SCALE_THRESHOLD
MEDIAN_SUBSAMPLE = 1000         # rows entering the median heuristic
EXP_DD_LIMIT = 350.0            # kExpDdLimit of csrc/stein_math.cuh: |lw| up to which the strict mode pre-exponentiates


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError('gradient_free_mcmc_postprocessing_b200 needs a CUDA device (sm_100a); '
                           'there is no CPU fallback')
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    if device.type != 'cuda':
        raise ValueError(f'device must be a CUDA device, got {device}')
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    return device


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _to_device_f64(a, device, name):
    """numpy array / CPU tensor (pinned or not) / CUDA tensor -> contiguous float64 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        if t.device != device:
            t = t.to(device, non_blocking=True)
        return t.contiguous()
    arr = np.ascontiguousarray(np.asarray(a), dtype=np.float64)
    if arr.size == 0:
        return torch.empty(arr.shape, dtype=torch.float64, device=device)
    return torch.from_numpy(arr).to(device, non_blocking=False)


def _replicate_from_host(a, device, lib):
    """Multi-GPU: every rank needs the full replica, but only 1/R of it has to cross PCIe on each rank -
    upload this rank's block of rows and all-gather the rest over NVLink (torch.distributed plumbing)."""
    import torch.distributed as dist
    world, rank = lib.gfk_comm_size(), lib.gfk_comm_rank()
    host = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float64))
    if host.dtype != torch.float64:
        host = host.to(torch.float64)
    host = host.contiguous()
    n = host.shape[0]
    per = -(-n // world)
    begin, end = min(rank * per, n), min((rank + 1) * per, n)
    tail = tuple(host.shape[1:])
    mine = torch.zeros((per,) + tail, dtype=torch.float64, device=device)
    if end > begin:
        mine[:end - begin].copy_(host[begin:end], non_blocking=True)
    full = torch.empty((per * world,) + tail, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, mine)
    return full[:n] if per * world != n else full


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _bytes(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def scalar_linv(value, d):
    """``inv(value * I)`` exactly as ``stein_thinning.kernel.make_precon`` forms it."""
    return np.linalg.inv(value * np.identity(d))


def median_sq_device(x_dev, device):
    """Squared median pairwise distance of the (<= 1000-row subsample of the) device sample."""
    lib = _lib.load()
    n, d = x_dev.shape
    if n < 2:
        raise Exception('Too few unique samples in sample.')
    ws = _bytes(lib.gfk_median_ws_bytes(n, MEDIAN_SUBSAMPLE), device)
    mid = (ctypes.c_double * 2)()
    check(lib.gfk_median_pdist(_ptr(x_dev), n, d, MEDIAN_SUBSAMPLE, mid, _ptr(ws), _stream_ptr(device)))
    # np.median of an even count is np.mean of the two middle order statistics
    med = np.mean(np.array([mid[0], mid[1]]))
    return med ** 2


class SteinProblem:
    """Standardised sample, scores, optional GF log-weights and preconditioner, resident in HBM.

    Mirrors what ``_make_stein_integrand`` / ``_make_stein_gf_integrand`` capture in their
    closures (reference: code/notebooks/examples/JAX_Stein_Thinning.ipynb#c16-#c17,
    code/src/utils/ksd.py:25).
    """

    def __init__(self, sample, gradient, log_p=None, log_q=None, standardize=True, preconditioner='id',
                 range_cap=None, precision=None, device=None, c=1.0):
        lib = _lib.load()
        self.c = float(c)
        self.device = device = _require_cuda(device)
        self.mode = _lib.mode_code(precision)
        shape = tuple(sample.shape)
        if len(shape) != 2:
            raise ValueError('sample is not two-dimensional.')
        n, d = shape
        if n == 0 or d == 0:
            raise ValueError('sample is empty.')
        if tuple(gradient.shape) != (n, d):
            raise ValueError('Dimensions of sample and gradient are inconsistent.')
        if n >= 2 ** 32 - 1:
            raise ValueError('n must be below 2**32 - 1 (uint32 indices).')
        self.gf = log_p is not None
        if self.gf and (tuple(np.shape(log_p)) != (n,) or tuple(np.shape(log_q)) != (n,)):
            raise ValueError('Dimensions of sample and log-densities are inconsistent.')
        self.n, self.d = n, d

        with torch.cuda.device(device):
            load = self._loader(sample, device, lib)
            x_raw = load(sample, device, 'sample')
            g_raw = load(gradient, device, 'gradient')
            st = _stream_ptr(device)
            # a3: validation + standardisation
            self.x = torch.empty_like(x_raw)
            self.s = torch.empty_like(g_raw)
            ws = _bytes(lib.gfk_standardize_ws_bytes(n, d), device)
            scl = (ctypes.c_double * d)()
            rc = lib.gfk_standardize(_ptr(x_raw), _ptr(g_raw), n, d, int(bool(standardize)), _ptr(self.x),
                                     _ptr(self.s), scl, _ptr(ws), st)
            if rc == _lib.GFK_ENUMERIC:
                raise ValueError(lib.gfk_last_error().decode())
            check(rc)
            self.scale = np.array(scl[:])
            # a2: preconditioner
            self.xl = None
            self._set_preconditioner(preconditioner)
            # a4: gradient-free log-weights
            self.lw = self.w = None
            if self.gf:
                lp = load(log_p, device, 'log_p')
                lq = load(log_q, device, 'log_q')
                self.lw = torch.empty(n, dtype=torch.float64, device=device)
                spread = ctypes.c_double(0.0)
                lws = _bytes(lib.gfk_log_weights_ws_bytes(), device)
                rc = lib.gfk_log_weights(_ptr(lp), _ptr(lq), n, int(range_cap is not None),
                                         float(range_cap if range_cap is not None else 0.0), _ptr(self.lw),
                                         ctypes.byref(spread), _ptr(lws), st)
                if rc == _lib.GFK_ENUMERIC:
                    raise ValueError(lib.gfk_last_error().decode())
                check(rc)
                if spread.value > WEIGHT_SCALE_THRESHOLD:
                    warnings.warn(f'log_q differs from log_p by more than {WEIGHT_SCALE_THRESHOLD} - '
                                  'consider using q that matches target better')
            self._finalise_dtype()

    @staticmethod
    def _loader(sample, device, lib):
        """Host inputs on a multi-rank communicator are uploaded in shards and all-gathered over NVLink."""
        on_host = not (isinstance(sample, torch.Tensor) and sample.is_cuda)
        if on_host and lib.gfk_comm_size() > 1:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_backend() == 'nccl' \
                    and dist.get_world_size() == lib.gfk_comm_size():
                def sharded(a, dev, name):
                    if isinstance(a, torch.Tensor) and a.is_cuda:
                        return _to_device_f64(a, dev, name)
                    return _replicate_from_host(a, dev, lib).contiguous()
                return sharded
        return _to_device_f64

    # -- preconditioner -------------------------------------------------------------------------
    def _set_preconditioner(self, preconditioner):
        d, n = self.d, self.n
        self.preconditioner = preconditioner
        dense = None
        if isinstance(preconditioner, np.ndarray):      # explicit linv (vfk0_imq)
            linv = np.atleast_2d(np.asarray(preconditioner, dtype=float))
            if linv.shape != (d, d):
                raise ValueError('linv must be d x d')
            if not np.array_equal(linv, linv[0, 0] * np.identity(d)):
                dense = linv
        elif preconditioner == 'id':
            linv = np.identity(d)
        elif preconditioner in ('med', 'sclmed'):
            m2 = median_sq_device(self.x, self.device)
            if m2 == 0:
                raise Exception('Too few unique samples in sample.')
            if preconditioner == 'sclmed':
                m2 = m2 / np.log(np.minimum(MEDIAN_SUBSAMPLE, n))
            linv = scalar_linv(m2, d)
        elif preconditioner == 'smpcov':
            if n < 2:
                raise Exception('Too few unique samples in sample.')
            lib = _lib.load()
            ws = _bytes(lib.gfk_covariance_ws_bytes(n, d), self.device)
            cov_host = (ctypes.c_double * (d * d))()
            check(lib.gfk_covariance(_ptr(self.x), n, d, cov_host, _ptr(ws), _stream_ptr(self.device)))
            cov = np.array(cov_host[:]).reshape(d, d)
            if not np.all(np.linalg.eigvals(cov) > 0):
                raise Exception('Too few unique samples in sample.')
            linv = dense = np.linalg.inv(cov)
        else:
            try:
                value = float(preconditioner)
            except (TypeError, ValueError):
                raise ValueError('Incorrect preconditioner type.') from None
            linv = scalar_linv(value, d)
        self.linv = linv
        if dense is None:
            self.lam = float(linv[0, 0])
            self.lam_sq = float(np.dot(linv, linv)[0, 0])
        else:
            # dense L: the kernels use xl = x @ L and unit lambda (stein_math.cuh), which equals the reference's
            # x' L L x / (L ds)' x terms only for a SYMMETRIC L - anything else is refused, not symmetrised
            asym = float(np.max(np.abs(dense - dense.T)))
            if asym > 1e-12 * float(np.max(np.abs(dense))):
                raise ValueError('a dense linv must be symmetric (max |L - L^T| = %.3g): the reference evaluates '
                                 "x' L L x with L as given, which the device formulation reproduces only for "
                                 'symmetric L' % asym)
            sym = 0.5 * (dense + dense.T)          # removes rounding-level asymmetry only
            lib = _lib.load()
            sym_dev = torch.from_numpy(np.ascontiguousarray(sym)).to(self.device)
            self.xl = torch.empty_like(self.x)
            check(lib.gfk_rows_times_matrix(_ptr(self.x), n, d, _ptr(sym_dev), _ptr(self.xl), _stream_ptr(self.device)))
            self.lam = self.lam_sq = 1.0
            if self.mode == _lib.F64_STRICT:
                # numpy's BLAS summation order for np.dot(linv, .) is not reproducible: no strict formulation exists
                self.mode = _lib.F64_FAST
                warnings.warn("a dense preconditioner runs in the 'f64-fast' formulation (FMA sums); the strict "
                              "'f64' parity mode covers scalar preconditioners ('id', 'med', 'sclmed', numeric) only")
        self.trace = float(np.trace(linv))

    @property
    def precision(self):
        """The arithmetic mode the kernels actually run in ('f64' may have become 'f64-fast', see above)."""
        return {_lib.F64_STRICT: 'f64', _lib.F64_FAST: 'f64-fast', _lib.F32_FAST: 'f32'}[self.mode]

    def _finalise_dtype(self):
        """Build the arrays the kernels read: float32 copies for f32, exp(lw) for the fast modes."""
        lib = _lib.load()
        st = _stream_ptr(self.device)
        kx, ks, klw, kxl = self.x, self.s, self.lw, self.xl
        if self.mode == _lib.F32_FAST:
            def f32(t):
                if t is None:
                    return None
                out = torch.empty(t.shape, dtype=torch.float32, device=self.device)
                check(lib.gfk_to_f32(_ptr(t), _ptr(out), t.numel(), st))
                return out
            kx, ks, klw, kxl = f32(self.x), f32(self.s), f32(self.lw), f32(self.xl)
        if self.gf and self.mode != _lib.F64_STRICT:
            self.w = torch.empty_like(klw)
            check(lib.gfk_exp_weights(_ptr(klw), _ptr(self.w), self.n, int(self.mode == _lib.F32_FAST), st))
        elif self.gf:
            # strict mode: per-row double-double exp(lw) so that the pair weight exp(fl(lw_i + lw_j)) costs 12 IEEE
            # operations, is correctly rounded and machine independent (csrc/stein_math.cuh exp_sum_dd) - unless a
            # log-weight is so large that exp(lw_i) alone could overflow or go subnormal; then the device exp stays
            rng = (ctypes.c_double * 2)()
            check(lib.gfk_log_weights_last_range(rng))         # host statistics of the gfk_log_weights call above
            self.lw_range = (rng[0], rng[1])
            # GFK_EXP_DD=0: keep the device exp (A/B measurements only)
            if -EXP_DD_LIMIT <= rng[0] and rng[1] <= EXP_DD_LIMIT and os.environ.get('GFK_EXP_DD', '1') != '0':
                self.w = torch.empty((self.n, 2), dtype=torch.float64, device=self.device)
                check(lib.gfk_exp_weights_dd(_ptr(klw), _ptr(self.w), self.n, st))
        self._kx, self._ks, self._klw, self._kxl = kx, ks, klw, kxl
        self.struct = Problem(x=kx.data_ptr(), s=ks.data_ptr(),
                              lw=klw.data_ptr() if klw is not None else None,
                              w=self.w.data_ptr() if self.w is not None else None,
                              xl=kxl.data_ptr() if kxl is not None else None,
                              n=self.n, d=self.d, mode=self.mode, lambda_=self.lam, lambda_sq=self.lam_sq,
                              trace=self.trace, c=self.c)

    # -- operations -------------------------------------------------------------------------------
    def thin(self, n_points, row_range=None, return_objective=False, timing=None):
        """``_greedy_search`` over this problem; returns ``np.ndarray`` of ``uint32`` (length n_points).
        ``timing`` (dict, optional) receives ``loop_ms``: CUDA-event time of the gfk_thin launches alone."""
        lib = _lib.load()
        m = int(n_points)
        if m < 1:
            raise ValueError('n_points must be at least 1.')
        if row_range is None:
            from .distributed import shard_bounds
            row_range = shard_bounds(self.n, lib.gfk_comm_size(), lib.gfk_comm_rank())
        r0, r1 = row_range
        with torch.cuda.device(self.device):
            objective = torch.empty(r1 - r0, dtype=torch.float64, device=self.device)
            idx = torch.empty(m, dtype=torch.int32, device=self.device)
            ws = _bytes(lib.gfk_thin_ws_bytes(), self.device)
            if timing is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            check(lib.gfk_thin(ctypes.byref(self.struct), r0, r1, m, _ptr(objective), _ptr(idx), _ptr(ws),
                               _stream_ptr(self.device)))
            if timing is not None:
                e1.record()
            check(lib.gfk_thin_check(_ptr(ws), _stream_ptr(self.device)))
            out = idx.cpu().numpy().view(np.uint32)
            if timing is not None:
                timing['loop_ms'] = e0.elapsed_time(e1)
                timing['rows'] = r1 - r0
                timing['ws'] = ws
        if return_objective:
            return out, objective
        return out

    def _sel(self, sel):
        if sel is None:
            return None
        sel = np.ascontiguousarray(np.asarray(sel).astype(np.int64))
        if sel.ndim != 1 or (sel.size and (sel.min() < 0 or sel.max() >= self.n)):
            raise IndexError('selection indices out of range')
        return torch.from_numpy(sel.astype(np.uint32).view(np.int32)).to(self.device)

    def ksd(self, sel=None, m=None):
        """Cumulative KSD over rows ``sel`` (or the first ``m`` rows); float64 ndarray of length m."""
        lib = _lib.load()
        sel_dev = self._sel(sel)
        m = int(m if m is not None else (len(sel) if sel is not None else self.n))
        with torch.cuda.device(self.device):
            st = _stream_ptr(self.device)
            low = torch.empty(m, dtype=torch.float64, device=self.device)
            diag = torch.empty(m, dtype=torch.float64, device=self.device)
            ks = torch.empty(m, dtype=torch.float64, device=self.device)
            # opt-in (GFK_KSD_MMA=1): the row sums on the tensor cores - float32 mode, scalar preconditioner,
            # d <= 512 (csrc/ksd_mma.cuh; not yet measured on a GPU, hence not the default)
            use_mma = (os.environ.get('GFK_KSD_MMA') == '1' and lib.gfk_ksd_mma_supported(ctypes.byref(self.struct)))
            if use_mma:
                ws = _bytes(lib.gfk_ksd_mma_ws_bytes(m, self.d) + 256, self.device)
                base = (ws.data_ptr() + 255) // 256 * 256
                rc = lib.gfk_ksd_rowsums_mma(ctypes.byref(self.struct), _ptr(sel_dev), m, lib.gfk_comm_rank(),
                                             lib.gfk_comm_size(), _ptr(low), _ptr(diag), ctypes.c_void_p(base), st)
                if rc == _lib.GFK_EUNSUPPORTED:
                    use_mma = False                      # e.g. no cuTensorMapEncodeTiled in this driver
                else:
                    check(rc)
            if not use_mma:
                check(lib.gfk_ksd_rowsums(ctypes.byref(self.struct), _ptr(sel_dev), m, lib.gfk_comm_rank(),
                                          lib.gfk_comm_size(), _ptr(low), _ptr(diag), st))
            if lib.gfk_comm_size() > 1:
                check(lib.gfk_allreduce_sum(_ptr(low), m, st))
                check(lib.gfk_allreduce_sum(_ptr(diag), m, st))
            check(lib.gfk_ksd_finish(_ptr(low), _ptr(diag), m, _ptr(ks), st))
            if use_mma:
                check(lib.gfk_ksd_mma_check(ctypes.c_void_p(base), st))
            return ks.cpu().numpy()

    def kmat(self, sel=None, m=None):
        lib = _lib.load()
        sel_dev = self._sel(sel)
        m = int(m if m is not None else (len(sel) if sel is not None else self.n))
        with torch.cuda.device(self.device):
            out = torch.empty((m, m), dtype=torch.float64, device=self.device)
            check(lib.gfk_kmat(ctypes.byref(self.struct), _ptr(sel_dev), m, _ptr(out), _stream_ptr(self.device)))
            return out.cpu().numpy()

    def pairs(self, ia, ib):
        """Element-wise ``k(ia[t], ib[t])`` (the integrand protocol); float64 ndarray."""
        lib = _lib.load()
        ia = np.ascontiguousarray(ia, dtype=np.int64)
        ib = np.ascontiguousarray(ib, dtype=np.int64)
        if ia.shape != ib.shape:
            raise ValueError('index arrays must have the same shape')
        for a in (ia, ib):
            if a.size and (a.min() < 0 or a.max() >= self.n):
                raise IndexError('index out of range')
        with torch.cuda.device(self.device):
            a_dev = torch.from_numpy(ia.reshape(-1)).to(self.device)
            b_dev = torch.from_numpy(ib.reshape(-1)).to(self.device)
            out = torch.empty(ia.size, dtype=torch.float64, device=self.device)
            check(lib.gfk_pairs(ctypes.byref(self.struct), _ptr(a_dev), _ptr(b_dev), ia.size, _ptr(out),
                                _stream_ptr(self.device)))
            return out.cpu().numpy().reshape(ia.shape)


def thin_batched(problems, n_points):
    """``[p.thin(n_points) for p in problems]`` in ONE cooperative launch (``gfk_thin_batched``: one group of CTAs per
    chain), for independent problems on one device.  The launch is local to this rank whatever communicator is active
    (chains are dealt over ranks by the caller, no exchange).  Returns ``None`` - nothing launched - when the set is
    not eligible (d > 4, mixed kernel kinds or precisions, a dense preconditioner, a chain too long for its CTA
    group); the caller then thins the chains one after the other."""
    lib = _lib.load()
    m = int(n_points)
    if m < 1:
        raise ValueError('n_points must be at least 1.')
    n = len(problems)
    if n == 0:
        return []
    device = problems[0].device
    if any(p.device != device for p in problems):
        return None
    with torch.cuda.device(device):
        structs = (Problem * n)(*[p.struct for p in problems])
        objectives = [torch.empty(p.n, dtype=torch.float64, device=device) for p in problems]
        idx = [torch.empty(m, dtype=torch.int32, device=device) for _ in problems]
        obj_ptrs = (ctypes.c_void_p * n)(*[o.data_ptr() for o in objectives])
        idx_ptrs = (ctypes.c_void_p * n)(*[i.data_ptr() for i in idx])
        ws = _bytes(lib.gfk_thin_batched_ws_bytes(n) + 256, device)
        base = (ws.data_ptr() + 255) // 256 * 256
        rc = lib.gfk_thin_batched(structs, n, m, obj_ptrs, idx_ptrs, ctypes.c_void_p(base), _stream_ptr(device))
        if rc == _lib.GFK_EUNSUPPORTED:
            return None
        check(rc)
        check(lib.gfk_thin_batched_check(ctypes.c_void_p(base), n, _stream_ptr(device)))
        return [i.cpu().numpy().view(np.uint32) for i in idx]
