"""Driver of ``tests/host/libgfk_sim.so``: the real kernel sources of libgfk compiled for the SIMT host emulator
(tests/host/simt).  TEST INFRASTRUCTURE ONLY - the product package never loads this library; it lets the CPU suite
execute the CUDA kernels' actual code (index arithmetic, shared-memory layouts, barriers, reductions, the
flag-in-data grid exchange) and compare it with the oracle when no GPU is available.

"Device" pointers are host pointers; inputs are placed right in front of an inaccessible page, so a read or write
past the end of a buffer faults instead of going unnoticed.
"""
import ctypes
import mmap
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SIM_SO = os.path.join(HERE, 'host', 'libgfk_sim.so')
CSRC = os.path.join(ROOT, 'gradient_free_mcmc_postprocessing_b200', 'csrc')

F64_STRICT, F64_FAST, F32_FAST = 0, 1, 2
MODES = {'f64': F64_STRICT, 'f64-fast': F64_FAST, 'f32': F32_FAST}
MEDIAN_SUBSAMPLE = 1000

_lib = None
_libc = ctypes.CDLL(None, use_errno=True)
_PAGE = mmap.PAGESIZE


def _stale():
    if not os.path.exists(SIM_SO):
        return True
    t = os.path.getmtime(SIM_SO)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    srcs += [os.path.join(HERE, 'host', 'simt', f) for f in ('simt.h', 'simt.cpp', 'cu2cpp.py')]
    srcs.append(os.path.join(ROOT, 'include', 'gfk.h'))
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def load():
    """The emulator build of libgfk with the product's own ctypes prototypes attached."""
    global _lib
    if _lib is None:
        if _stale():
            subprocess.run(['bash', os.path.join(HERE, 'host', 'build.sh')], check=True, stdout=subprocess.DEVNULL)
        from gradient_free_mcmc_postprocessing_b200._lib import PROTOTYPES
        lib = ctypes.CDLL(SIM_SO)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        lib.simt_set_schedule.argtypes = [ctypes.c_int, ctypes.c_uint]
        lib.simt_set_schedule.restype = None
        _lib = lib
    return _lib


def set_schedule(random=False, seed=1):
    """Thread order between rendezvous points.  SIMT_FORCE_RANDOM=<seed> turns every test into a random-schedule
    run (a one-off robustness sweep: `SIMT_FORCE_RANDOM=7 python -m pytest tests -m "not gpu" -k simt`)."""
    forced = os.environ.get('SIMT_FORCE_RANDOM')
    if forced:
        random, seed = True, int(forced)
    load().simt_set_schedule(int(bool(random)), int(seed))


def set_sms(count):
    os.environ['SIMT_SMS'] = str(int(count))


class SimError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise SimError('libgfk (emulated) error %d: %s' % (rc, load().gfk_last_error().decode('utf-8', 'replace')))


def guarded(shape, dtype, fill=None):
    """ndarray whose last byte sits (up to alignment) right before a PROT_NONE page."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    span = (max(nbytes, 1) + _PAGE - 1) // _PAGE * _PAGE
    mm = mmap.mmap(-1, span + _PAGE)                 # stays mapped as long as the returned array (its .base) lives
    probe = ctypes.c_char.from_buffer(mm)
    base = ctypes.addressof(probe)
    del probe
    if _libc.mprotect(ctypes.c_void_p(base + span), ctypes.c_size_t(_PAGE), 0) != 0:
        raise OSError(ctypes.get_errno(), 'mprotect')
    off = (span - nbytes) // 16 * 16
    arr = np.frombuffer(mm, dtype=dtype, count=int(np.prod(shape, dtype=np.int64)), offset=off).reshape(shape)
    if fill is not None:
        arr[...] = fill
    return arr


def dev(a, dtype=np.float64):
    """Copy of ``a`` in a guarded "device" buffer."""
    a = np.ascontiguousarray(a, dtype=dtype)
    out = guarded(a.shape, dtype)
    out[...] = a
    return out


def ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


from gradient_free_mcmc_postprocessing_b200._lib import Problem   # mirror of struct gfk_problem (ctypes only)


class SimProblem:
    """The host orchestration of ``_engine.SteinProblem`` over numpy buffers and the emulated library."""

    def __init__(self, sample, gradient, log_p=None, log_q=None, standardize=True, preconditioner='id',
                 range_cap=None, precision='f64'):
        lib = load()
        self.mode = MODES[precision]
        sample = np.asarray(sample, dtype=float)
        n, d = sample.shape
        self.n, self.d = n, d
        x_raw, g_raw = dev(sample), dev(gradient)
        self.x, self.s = guarded((n, d), np.float64), guarded((n, d), np.float64)
        ws = guarded((lib.gfk_standardize_ws_bytes(n, d),), np.uint8)
        scl = (ctypes.c_double * d)()
        check(lib.gfk_standardize(ptr(x_raw), ptr(g_raw), n, d, int(bool(standardize)), ptr(self.x), ptr(self.s), scl,
                                  ptr(ws), None))
        self.scale = np.array(scl[:])
        self.xl = None
        dense = None
        if isinstance(preconditioner, np.ndarray):
            linv = np.atleast_2d(preconditioner)
            if not np.array_equal(linv, linv[0, 0] * np.identity(d)):
                dense = linv
        elif preconditioner == 'id':
            linv = np.identity(d)
        elif preconditioner in ('med', 'sclmed'):
            m2 = self.median_sq()
            if preconditioner == 'sclmed':
                m2 = m2 / np.log(np.minimum(MEDIAN_SUBSAMPLE, n))
            linv = np.linalg.inv(m2 * np.identity(d))
        else:
            linv = np.linalg.inv(float(preconditioner) * np.identity(d))
        self.linv = linv
        if dense is None:
            self.lam, self.lam_sq = float(linv[0, 0]), float(np.dot(linv, linv)[0, 0])
        else:
            self.xl = dev(self.x @ (0.5 * (dense + dense.T)))
            self.lam = self.lam_sq = 1.0
            if self.mode == F64_STRICT:
                self.mode = F64_FAST
        self.trace = float(np.trace(linv))
        self.lw = self.w = None
        self.spread = None
        if log_p is not None:
            lp, lq = dev(log_p), dev(log_q)
            self.lw = guarded((n,), np.float64)
            spread = ctypes.c_double(0.0)
            lws = guarded((lib.gfk_log_weights_ws_bytes(),), np.uint8)
            check(lib.gfk_log_weights(ptr(lp), ptr(lq), n, int(range_cap is not None),
                                      float(range_cap if range_cap is not None else 0.0), ptr(self.lw),
                                      ctypes.byref(spread), ptr(lws), None))
            self.spread = spread.value
        kx, ks, klw, kxl = self.x, self.s, self.lw, self.xl
        if self.mode == F32_FAST:
            def f32(t):
                if t is None:
                    return None
                out = guarded(t.shape, np.float32)
                check(lib.gfk_to_f32(ptr(t), ptr(out), t.size, None))
                return out
            kx, ks, klw, kxl = f32(kx), f32(ks), f32(klw), f32(kxl)
        if klw is not None and self.mode != F64_STRICT:
            self.w = guarded(klw.shape, klw.dtype)
            check(lib.gfk_exp_weights(ptr(klw), ptr(self.w), n, int(self.mode == F32_FAST), None))
        elif klw is not None:                  # strict mode: double-double exponentials when the range allows
            rng = (ctypes.c_double * 2)()
            check(lib.gfk_log_weights_last_range(rng))
            rng2 = (ctypes.c_double * 2)()                     # the kernel-computed range must agree
            rws = guarded((lib.gfk_log_weights_ws_bytes(),), np.uint8)
            check(lib.gfk_log_weights_range(ptr(klw), n, rng2, ptr(rws), None))
            assert (rng[0], rng[1]) == (rng2[0], rng2[1]), ((rng[0], rng[1]), (rng2[0], rng2[1]))
            self.lw_range = (rng[0], rng[1])
            if -350.0 <= rng[0] and rng[1] <= 350.0 and os.environ.get('GFK_TEST_NO_EXP_DD') != '1':
                self.w = guarded((n, 2), np.float64, fill=np.nan)
                check(lib.gfk_exp_weights_dd(ptr(klw), ptr(self.w), n, None))
        self._keep = (kx, ks, klw, kxl)
        self.struct = Problem(x=kx.ctypes.data, s=ks.ctypes.data, lw=klw.ctypes.data if klw is not None else None,
                              w=self.w.ctypes.data if self.w is not None else None,
                              xl=kxl.ctypes.data if kxl is not None else None, n=n, d=d, mode=self.mode,
                              lambda_=self.lam, lambda_sq=self.lam_sq, trace=self.trace, c=1.0)

    def median_sq(self):
        lib = load()
        ws = guarded((lib.gfk_median_ws_bytes(self.n, MEDIAN_SUBSAMPLE),), np.uint8)
        mid = (ctypes.c_double * 2)()
        check(lib.gfk_median_pdist(ptr(self.x), self.n, self.d, MEDIAN_SUBSAMPLE, mid, ptr(ws), None))
        return np.mean(np.array([mid[0], mid[1]])) ** 2

    def thin(self, m, row_range=None):
        lib = load()
        r0, r1 = row_range if row_range is not None else (0, self.n)
        objective = guarded((r1 - r0,), np.float64, fill=np.nan)
        idx = guarded((m,), np.uint32, fill=0xffffffff)
        ws = guarded((lib.gfk_thin_ws_bytes(),), np.uint8, fill=0xa5)
        check(lib.gfk_thin(ctypes.byref(self.struct), r0, r1, m, ptr(objective), ptr(idx), ptr(ws), None))
        check(lib.gfk_thin_check(ptr(ws), None))
        return idx.copy(), objective.copy()

    def _sel(self, sel):
        return None if sel is None else dev(np.asarray(sel), np.uint32)

    def ksd_rowsums(self, sel=None, m=None, part=0, n_parts=1):
        lib = load()
        sel_dev = self._sel(sel)
        m = int(m if m is not None else (len(sel) if sel is not None else self.n))
        low, diag = guarded((m,), np.float64, fill=np.nan), guarded((m,), np.float64, fill=np.nan)
        check(lib.gfk_ksd_rowsums(ctypes.byref(self.struct), ptr(sel_dev), m, part, n_parts, ptr(low), ptr(diag), None))
        return low.copy(), diag.copy()

    def ksd_rowsums_mma(self, sel=None, m=None, part=0, n_parts=1):
        """gfk_ksd_rowsums_mma (tensor-core row sums); None when the problem is not eligible."""
        lib = load()
        if not lib.gfk_ksd_mma_supported(ctypes.byref(self.struct)):
            return None
        sel_dev = self._sel(sel)
        m = int(m if m is not None else (len(sel) if sel is not None else self.n))
        low, diag = guarded((m,), np.float64, fill=np.nan), guarded((m,), np.float64, fill=np.nan)
        raw = guarded((lib.gfk_ksd_mma_ws_bytes(m, self.d) + 256,), np.uint8, fill=0xa5)
        base = (raw.ctypes.data + 255) // 256 * 256
        check(lib.gfk_ksd_rowsums_mma(ctypes.byref(self.struct), ptr(sel_dev), m, part, n_parts, ptr(low), ptr(diag),
                                      ctypes.c_void_p(base), None))
        check(lib.gfk_ksd_mma_check(ctypes.c_void_p(base), None))
        return low.copy(), diag.copy()

    def ksd_from_rowsums(self, low, diag):
        lib = load()
        m = len(low)
        low_d, diag_d = dev(low), dev(diag)
        ks = guarded((m,), np.float64, fill=np.nan)
        check(lib.gfk_ksd_finish(ptr(low_d), ptr(diag_d), m, ptr(ks), None))
        return ks.copy()

    def ksd(self, sel=None, m=None, n_parts=1):
        lib = load()
        m = int(m if m is not None else (len(sel) if sel is not None else self.n))
        low, diag = np.zeros(m), np.zeros(m)
        for part in range(n_parts):          # what the all-reduce over ranks adds up
            l, dg = self.ksd_rowsums(sel, m, part, n_parts)
            low += l
            diag += dg
        low_d, diag_d = dev(low), dev(diag)
        ks = guarded((m,), np.float64, fill=np.nan)
        check(lib.gfk_ksd_finish(ptr(low_d), ptr(diag_d), m, ptr(ks), None))
        return ks.copy()

    def kmat(self, sel=None, m=None):
        lib = load()
        sel_dev = self._sel(sel)
        m = int(m if m is not None else (len(sel) if sel is not None else self.n))
        out = guarded((m, m), np.float64, fill=np.nan)
        check(lib.gfk_kmat(ctypes.byref(self.struct), ptr(sel_dev), m, ptr(out), None))
        return out.copy()

    def pairs(self, ia, ib):
        lib = load()
        ia_d, ib_d = dev(ia, np.int64), dev(ib, np.int64)
        out = guarded((ia_d.size,), np.float64, fill=np.nan)
        check(lib.gfk_pairs(ctypes.byref(self.struct), ptr(ia_d), ptr(ib_d), ia_d.size, ptr(out), None))
        return out.copy()


def thin_batched(problems, m):
    """gfk_thin_batched over SimProblems: index vectors and objectives per chain, or None when the set is not eligible."""
    lib = load()
    n = len(problems)
    structs = (Problem * n)(*[p.struct for p in problems])
    objectives = [guarded((p.n,), np.float64, fill=np.nan) for p in problems]
    idx = [guarded((m,), np.uint32, fill=0xffffffff) for _ in problems]
    obj_ptrs = (ctypes.c_void_p * n)(*[o.ctypes.data for o in objectives])
    idx_ptrs = (ctypes.c_void_p * n)(*[i.ctypes.data for i in idx])
    nbytes = lib.gfk_thin_batched_ws_bytes(n)
    raw = guarded((nbytes + 256,), np.uint8, fill=0xa5)
    base = (raw.ctypes.data + 255) // 256 * 256
    rc = lib.gfk_thin_batched(structs, n, m, obj_ptrs, idx_ptrs, ctypes.c_void_p(base), None)
    if rc == 4:                                          # GFK_EUNSUPPORTED
        return None
    check(rc)
    check(lib.gfk_thin_batched_check(ctypes.c_void_p(base), n, None))
    return [i.copy() for i in idx], [o.copy() for o in objectives]


# ---- proxies and energy distance through the emulated kernels -----------------------------------
def student_t_proxy(x, mu, sigma, df):
    from math import lgamma
    lib = load()
    x = np.ascontiguousarray(x, dtype=float)
    n, d = x.shape
    cst = lgamma(0.5 * (df + d)) - lgamma(0.5 * df) - d / 2. * np.log(df * np.pi) - 0.5 * np.linalg.slogdet(sigma)[1]
    xd, md, pd_ = dev(x), dev(mu), dev(np.linalg.inv(sigma))
    log_q, grad_q = guarded((n,), np.float64, fill=np.nan), guarded((n, d), np.float64, fill=np.nan)
    check(lib.gfk_student_t_proxy(ptr(xd), n, d, ptr(md), ptr(pd_), ctypes.c_double(df), ctypes.c_double(cst),
                                  ptr(log_q), ptr(grad_q), None))
    return log_q.copy(), grad_q.copy()


def kde_proxy(sample, weights=None, points=None, bw_method='silverman'):
    from gradient_free_mcmc_postprocessing_b200.proxies import kde_bandwidth
    lib = load()
    sample = np.ascontiguousarray(sample, dtype=float)
    h, w = kde_bandwidth(sample, weights, bw_method)
    d = h.shape[0]
    chol = np.linalg.cholesky(h)
    whiten = np.ascontiguousarray(np.linalg.inv(chol))
    log_norm = -0.5 * (d * np.log(2 * np.pi) + 2.0 * np.sum(np.log(np.diag(chol))))
    data = dev(sample)
    pts = data if points is None else dev(points)
    n, m = data.shape[0], pts.shape[0]
    wd = None if weights is None else dev(w)
    log_q, grad_q = guarded((m,), np.float64, fill=np.nan), guarded((m, d), np.float64, fill=np.nan)
    ws = guarded((lib.gfk_kde_ws_bytes(n, m, d),), np.uint8, fill=0xff)
    whiten_d = dev(whiten)                            # a named reference: the mapping lives as long as the array
    check(lib.gfk_kde_proxy(ptr(data), ptr(wd), n, ptr(pts), m, d, ptr(whiten_d), ctypes.c_double(log_norm),
                            ptr(log_q), ptr(grad_q), ptr(ws), None))
    return log_q.copy(), grad_q.copy()


def energy_cumulative(x, y):
    lib = load()
    xd, yd = dev(x), dev(y)
    n, m, d = xd.shape[0], yd.shape[0], xd.shape[1]
    e2 = guarded((m,), np.float64, fill=np.nan)
    ws = guarded((lib.gfk_energy_ws_bytes(n, m),), np.uint8, fill=0xff)
    check(lib.gfk_energy_distance_cumulative(ptr(xd), n, ptr(yd), m, d, ptr(e2), ptr(ws), None))
    return e2.copy()
