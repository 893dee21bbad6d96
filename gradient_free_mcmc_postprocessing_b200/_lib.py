"""ctypes binding of ``libgfk.so`` (the C ABI declared in ``include/gfk.h``).

The library is built in-tree by ``csrc/build.sh`` / ``__graft_entry__.build()``.
There is no CPU fallback: importing this module without the built library, or
calling into it without a CUDA device, raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GFK_LIB') or os.path.join(_HERE, 'libgfk.so')   # GFK_LIB: A/B builds only

GFK_OK = 0
GFK_EINVAL, GFK_ECUDA, GFK_ENCCL, GFK_EUNSUPPORTED, GFK_ENUMERIC = 1, 2, 3, 4, 5
F64_STRICT, F64_FAST, F32_FAST = 0, 1, 2
MODES = {'f64': F64_STRICT, 'f64-strict': F64_STRICT, 'f64-fast': F64_FAST, 'f32': F32_FAST}


class GfkError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'libgfk error {code}: {message}')
        self.code = code
        self.message = message


class Problem(ctypes.Structure):
    """Mirror of ``struct gfk_problem`` (include/gfk.h)."""
    _fields_ = [
        ('x', ctypes.c_void_p), ('s', ctypes.c_void_p), ('lw', ctypes.c_void_p),
        ('w', ctypes.c_void_p), ('xl', ctypes.c_void_p),
        ('n', ctypes.c_int64), ('d', ctypes.c_int32), ('mode', ctypes.c_int32),
        ('lambda_', ctypes.c_double), ('lambda_sq', ctypes.c_double),
        ('trace', ctypes.c_double), ('c', ctypes.c_double),
    ]


# name -> (restype, argtypes); every symbol include/gfk.h declares
_vp, _i32, _i64, _dbl, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t
_pp = ctypes.POINTER(Problem)
_pd = ctypes.POINTER(ctypes.c_double)
PROTOTYPES = {
    'gfk_version': (ctypes.c_int, []),
    'gfk_launch_count': (ctypes.c_longlong, [ctypes.c_int]),
    'gfk_last_error': (ctypes.c_char_p, []),
    'gfk_device_info': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                       ctypes.c_char_p, ctypes.c_int]),
    'gfk_standardize_ws_bytes': (_sz, [_i64, _i32]),
    'gfk_standardize': (ctypes.c_int, [_vp, _vp, _i64, _i32, ctypes.c_int, _vp, _vp, _pd, _vp, _vp]),
    'gfk_median_ws_bytes': (_sz, [_i64, _i32]),
    'gfk_median_pdist': (ctypes.c_int, [_vp, _i64, _i32, _i32, _pd, _vp, _vp]),
    'gfk_covariance_ws_bytes': (_sz, [_i64, _i32]),
    'gfk_covariance': (ctypes.c_int, [_vp, _i64, _i32, _pd, _vp, _vp]),
    'gfk_rows_times_matrix': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    'gfk_to_f32': (ctypes.c_int, [_vp, _vp, _i64, _vp]),
    'gfk_log_weights_ws_bytes': (_sz, []),
    'gfk_log_weights': (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, _dbl, _vp, _pd, _vp, _vp]),
    'gfk_exp_weights': (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, _vp]),
    'gfk_exp_weights_dd': (ctypes.c_int, [_vp, _vp, _i64, _vp]),
    'gfk_log_weights_last_range': (ctypes.c_int, [_pd]),
    'gfk_log_weights_range': (ctypes.c_int, [_vp, _i64, _pd, _vp, _vp]),
    'gfk_gaussian_proxy': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _dbl, _vp, _vp, _vp]),
    'gfk_student_t_proxy': (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _dbl, _dbl, _vp, _vp, _vp]),
    'gfk_kde_ws_bytes': (_sz, [_i64, _i64, _i32]),
    'gfk_kde_proxy': (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _dbl, _vp, _vp, _vp, _vp]),
    'gfk_energy_ws_bytes': (_sz, [_i64, _i64]),
    'gfk_energy_distance_cumulative': (ctypes.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]),
    'gfk_thin_ws_bytes': (_sz, []),
    'gfk_thin': (ctypes.c_int, [_pp, _i64, _i64, _i32, _vp, _vp, _vp, _vp]),
    'gfk_thin_check': (ctypes.c_int, [_vp, _vp]),
    'gfk_thin_batched_ws_bytes': (_sz, [_i32]),
    'gfk_thin_batched': (ctypes.c_int, [_pp, _i32, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp, _vp]),
    'gfk_thin_batched_check': (ctypes.c_int, [_vp, _i32, _vp]),
    'gfk_ksd_rowsums': (ctypes.c_int, [_pp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    'gfk_ksd_mma_supported': (ctypes.c_int, [_pp]),
    'gfk_ksd_mma_ws_bytes': (_sz, [_i64, _i32]),
    'gfk_ksd_rowsums_mma': (ctypes.c_int, [_pp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    'gfk_ksd_mma_check': (ctypes.c_int, [_vp, _vp]),
    'gfk_ksd_finish': (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    'gfk_kmat': (ctypes.c_int, [_pp, _vp, _i64, _vp, _vp]),
    'gfk_pairs': (ctypes.c_int, [_pp, _vp, _vp, _i64, _vp, _vp]),
    'gfk_comm_unique_id': (ctypes.c_int, [_vp, ctypes.c_char_p]),
    'gfk_comm_init': (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _vp, ctypes.c_char_p]),
    'gfk_p2p_local_handle': (ctypes.c_int, [_vp]),
    'gfk_p2p_open': (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    'gfk_p2p_enabled': (ctypes.c_int, []),
    'gfk_p2p_selftest': (ctypes.c_int, [ctypes.c_int, _pd, _vp]),
    'gfk_comm_size': (ctypes.c_int, []),
    'gfk_comm_rank': (ctypes.c_int, []),
    'gfk_allreduce_sum': (ctypes.c_int, [_vp, _i64, _vp]),
    'gfk_comm_finalize': (ctypes.c_int, []),
    'gfk_fma_peak': (ctypes.c_int, [ctypes.c_int, _pd, _vp]),
    'gfk_flush_l2': (ctypes.c_int, [_vp, _sz, _vp]),
}

_lib = None


def load():
    """Load libgfk.so (once) and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'{LIB_PATH} is missing: build the CUDA library first '
            '(python -c "import __graft_entry__ as g; g.build()" or csrc/build.sh). '
            'There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        if not hasattr(lib, name) and os.environ.get('GFK_LIB'):
            continue                  # A/B runs against an older build
        fn = getattr(lib, name)       # AttributeError if the build is stale
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != GFK_OK:
        msg = load().gfk_last_error()
        raise GfkError(rc, msg.decode('utf-8', 'replace') if msg else '')


def mode_code(precision):
    if precision is None:
        precision = os.environ.get('GFK_PRECISION', 'f64')
    try:
        return MODES[precision]
    except KeyError:
        raise ValueError(f'precision must be one of {sorted(MODES)}, got {precision!r}') from None
