"""ctypes binding of ``libqinfer_b200.so`` (the C ABI declared in include/qinfer_b200.h).

The product has no CPU fallback: if the shared library is missing or a call
fails this module raises, loudly.
"""
import ctypes
import os

QB_ABI_VERSION = 4
QB_SMALL_MAX = 4096
QB_MAX_D = 64
QB_MAX_RANKS = 16
QB_IPC_HANDLE_BYTES = 64
(QB_STAT_NORM, QB_STAT_SUMSQ, QB_STAT_MIN, QB_STAT_NBAD, QB_STAT_INV_NORM, QB_STAT_NESS, QB_STAT_TAG,
 QB_STAT_SKIPPED, QB_STAT_ATTN) = range(9)
QB_MAX_FUSE = 8
QB_STAT_COUNT = 16
QB_MODEL_PRECESSION, QB_MODEL_RB, QB_MODEL_TOMOGRAPHY, QB_MODEL_COIN = 1, 2, 3, 4
QB_SCAN_FAST, QB_SCAN_EXACT, QB_SCAN_FAST_GUIDE, QB_SCAN_FAST_GUIDE_SCALED = 0, 1, 2, 3
QB_WALK_ADD, QB_WALK_FIXED, QB_WALK_LEARNED = 0, 1, 2
QB_COUNT_AUTO, QB_COUNT_HISTOGRAM, QB_COUNT_TREE = 0, 1, 2

_LIB_PATH = os.environ.get("QB_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                         "libqinfer_b200.so")


class QbModel(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("d", ctypes.c_int32), ("binomial", ctypes.c_int32),
                ("interleaved", ctypes.c_int32), ("min_freq", ctypes.c_double),
                ("likelihood_power", ctypes.c_double), ("d_extra", ctypes.c_int32), ("extra_rule", ctypes.c_int32),
                ("fast_math", ctypes.c_int32), ("reserved0", ctypes.c_int32)]


class QbExpparams(ctypes.Structure):
    _fields_ = [("t", ctypes.c_double), ("w_", ctypes.c_double), ("m", ctypes.c_int64),
                ("reference", ctypes.c_int32), ("reserved", ctypes.c_int32), ("n_meas", ctypes.c_int64),
                ("meas", ctypes.c_double * QB_MAX_D)]


class QbUpdateCtl(ctypes.Structure):
    _fields_ = [("h_mirror", ctypes.c_void_p), ("tag", ctypes.c_double), ("zero_weight_thresh", ctypes.c_double),
                ("resample_below", ctypes.c_double), ("guard", ctypes.c_int32), ("check_resample", ctypes.c_int32),
                ("chain_prev_tag", ctypes.c_double), ("n_ranks", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("d_peer_mailbox", ctypes.c_void_p * QB_MAX_RANKS), ("d_error_flag", ctypes.c_void_p),
                ("h_shard_norms", ctypes.c_void_p)]


class QbError(RuntimeError):
    pass


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_F64 = ctypes.c_double
_SZ = ctypes.c_size_t
_U64 = ctypes.c_uint64

# name -> (restype, argtypes); every symbol include/qinfer_b200.h declares
SIGNATURES = {
    "qb_abi_version": (ctypes.c_int, []),
    "qb_last_error": (ctypes.c_char_p, []),
    "qb_device_sm_count": (ctypes.c_int, []),
    "qb_struct_sizes": (None, [ctypes.POINTER(ctypes.c_int32)]),
    "qb_weights_set_uniform": (ctypes.c_int, [_P, _I64, _P, _P]),
    "qb_weights_set_uniform_global": (ctypes.c_int, [_P, _I64, _I64, _P, _P]),
    "qb_weights_normalized": (ctypes.c_int, [_P, _I64, _P, _P, _P]),
    "qb_weights_restat": (ctypes.c_int, [_P, _I64, _P, _P, _SZ, _P]),
    "qb_weights_clip": (ctypes.c_int, [_P, _I64, _P, _P, _SZ, _P]),
    "qb_weights_min": (ctypes.c_int, [_P, _I64, _P, _P]),
    "qb_update_workspace_bytes": (_SZ, [_I64, _I32]),
    "qb_fused_update": (ctypes.c_int, [ctypes.POINTER(QbModel), ctypes.POINTER(QbExpparams), _I64, _P, _I64,
                                       _P, _P, _P, _P, ctypes.POINTER(QbUpdateCtl), _P, _SZ, _P]),
    "qb_fused_update_multi": (ctypes.c_int, [ctypes.POINTER(QbModel), ctypes.POINTER(QbExpparams),
                                             ctypes.POINTER(_I64), _I32, ctypes.c_uint32, _P, _I64, _P, _P, _P, _P,
                                             _P, ctypes.POINTER(QbUpdateCtl), _P, _SZ, _P]),
    "qb_likelihood": (ctypes.c_int, [ctypes.POINTER(QbModel), ctypes.POINTER(QbExpparams), _I32,
                                     ctypes.POINTER(_I64), _I32, _P, _I64, _P, _P]),
    "qb_hypothetical_update": (ctypes.c_int, [ctypes.POINTER(QbModel), ctypes.POINTER(QbExpparams), _I32,
                                              ctypes.POINTER(_I64), _I32, _P, _P, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "qb_design_workspace_bytes": (_SZ, [_I64, _I32, _I32]),
    "qb_design_sums": (ctypes.c_int, [ctypes.POINTER(QbModel), ctypes.POINTER(QbExpparams), ctypes.POINTER(_I64), _I32,
                                      _P, _P, _P, _I64, ctypes.POINTER(_F64), _P, _P, _P, _SZ, _P]),
    "qb_are_models_valid": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _P, _P]),
    "qb_moments_workspace_bytes": (_SZ, [_I64, _I32]),
    "qb_moments": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _SZ, _P]),
    "qb_cdf_workspace_bytes": (_SZ, [_I64]),
    "qb_cdf": (ctypes.c_int, [_P, _P, _I64, _P, _I32, _P, _SZ, _P]),
    "qb_cdf_chained": (ctypes.c_int, [_P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "qb_cdf_exact_fallback_flag": (ctypes.c_int, [_P, _I64, ctypes.POINTER(_I32), _P]),
    "qb_draw_workspace_bytes": (_SZ, [_I64]),
    "qb_draw": (ctypes.c_int, [_P, _I64, _P, _I64, _P, _P, _P, _SZ, _P]),
    "qb_lw_move_workspace_bytes": (_SZ, [_I32]),
    "qb_lw_move": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, _P, ctypes.POINTER(_F64),
                                  ctypes.POINTER(_F64), _F64, _P, _I64, _P, _I32, _P, _P, _P, _SZ, _P]),
    "qb_compact_workspace_bytes": (_SZ, [_I64]),
    "qb_compact_invalid": (ctypes.c_int, [_P, _I64, _P, _P, _P, _SZ, _P]),
    "qb_lw_retry": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, _P, _P, _I64, ctypes.POINTER(_F64),
                                   ctypes.POINTER(_F64), _F64, _P, _P, _P, _P, _I32, _P, _SZ, _P]),
    "qb_lw_draw_move": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, _P, _P, _SZ, _I32,
                                       ctypes.POINTER(_F64), ctypes.POINTER(_F64), _F64, _U64, _U64, _U64, _U64, _I32,
                                       _I64, _P, _I32, _P, _P, _P]),
    "qb_lw_draw_retry": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, _P, _P, _SZ, _I32,
                                        ctypes.POINTER(_F64), ctypes.POINTER(_F64), _F64, _U64, _U64, _U64, _U64, _I32,
                                        _P, _I64, _I32, _P, _P, _P, _P]),
    "qb_lw_merge_move": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, _P, _P, _SZ, _I32,
                                        ctypes.POINTER(_F64), ctypes.POINTER(_F64), _F64, _U64, _U64, _U64, _U64, _I32,
                                        _I64, _P, _I32, _P, _P, _P, _P, _P, _P]),
    "qb_lw_merge_retry": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, ctypes.POINTER(_F64),
                                         ctypes.POINTER(_F64), _F64, _U64, _U64, _P, _I64, _P, _P, _P, _P, _P]),
    "qb_lw_binned_workspace_bytes": (_SZ, [_I64, _I64]),
    "qb_lw_binned_sums": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _F64, _P, _SZ, _P]),
    "qb_lw_small_resample": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _P, _P, _I64, _I32, _F64, _F64, _F64, _P, _P, _I64,
                                            _P, _P, _P, _P, _P, _P, _I32, _P, _P, _F64, _P]),
    "qb_lw_binned_shard_consts": (ctypes.c_int, [_P, _I32, _I32, _F64, _F64, _F64, _P, _F64, _P, _SZ, _P]),
    "qb_lw_binned_count": (ctypes.c_int, [_I64, _I64, _U64, _U64, _I32, _P, _SZ, _P]),
    "qb_binomial_sample": (ctypes.c_int, [_I64, _F64, _I64, _U64, _U64, _P, _P]),
    "qb_lw_binned_prepare": (ctypes.c_int, [_P, _P, _P, _I64, _I32, _I64, _U64, _U64, _I32, _P, _P, _F64, _P, _SZ,
                                            _P]),
    "qb_lw_binned_move": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _P, _P, _I64, _I32, ctypes.POINTER(_F64),
                                         ctypes.POINTER(_F64), _F64, _U64, _U64, _U64, _U64, _I64, _P, _I64, _P, _P,
                                         _I64, _P, _I32, _I32, _I32, _P, _P, _P, _P, _F64, _P, _SZ, _P]),
    "qb_lw_binned_retry": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _I64, _I32, ctypes.POINTER(_F64),
                                          ctypes.POINTER(_F64), _F64, _U64, _U64, _U64, _I32, _I32, _U64, _I64, _P,
                                          _I64, _P, _P, _P, _P, _F64, _P, _SZ, _P]),
    "qb_lw_binned_resample": (ctypes.c_int, [ctypes.POINTER(QbModel), _P, _P, _P, _I64, _I32, _I64, _F64, _F64, _F64,
                                             _U64, _U64, _I32, _U64, _U64, _U64, _P, _P, _I64, _P, _I32, _I32, _I32,
                                             _P, _P, _P, _P, _F64, _P, _SZ, _P]),
    "qb_mailbox_create": (ctypes.c_int, [_I32, ctypes.POINTER(_P)]),
    "qb_mailbox_destroy": (ctypes.c_int, [_P]),
    "qb_ipc_get_handle": (ctypes.c_int, [_P, ctypes.c_char_p]),
    "qb_ipc_open_handle": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_P)]),
    "qb_ipc_close_handle": (ctypes.c_int, [_P]),
    "qb_shard_classify": (ctypes.c_int, [_P, _I64, ctypes.POINTER(_F64), _I32, _P, _P, _P]),
    "qb_shard_bucket": (ctypes.c_int, [_P, _P, _I64, ctypes.POINTER(_F64), ctypes.POINTER(_I64), _I32, _P, _P, _P,
                                       _P]),
    "qb_gather_rows": (ctypes.c_int, [_P, _I32, _P, _I64, _P, _P]),
    "qb_tomo_canonicalize": (ctypes.c_int, [_P, _I64, _I32, _P, _I32, _P]),
    "qb_tomo_canonicalize_screened": (ctypes.c_int, [_P, _I64, _I32, _P, _I32, _P, _P, _P, _P, _SZ, _P]),
    "qb_tomo_canonicalize_ld": (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _I32, _P]),
    "qb_tomo_canonicalize_screened_ld": (ctypes.c_int, [_P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P, _SZ, _P]),
    "qb_readside_workspace_bytes": (_SZ, []),
    "qb_weights_entropy": (ctypes.c_int, [_P, _P, _I64, _P, _P, _SZ, _P]),
    "qb_weight_mass_hist": (ctypes.c_int, [_P, _P, _I64, _I32, _I32, _U64, _P, _P, _P]),
    "qb_weights_select": (ctypes.c_int, [_P, _P, _I64, _U64, _I32, _P, _P]),
    "qb_walk_step": (ctypes.c_int, [_P, _I64, _I32, _I32, ctypes.POINTER(_I32), ctypes.POINTER(_I32), _I32,
                                    ctypes.POINTER(_F64), ctypes.POINTER(_I32), _F64, _F64, _P, _I32, _P]),
    "qb_poison_likelihood": (ctypes.c_int, [_P, _I64, _P, _I32, _F64, _F64, _P]),
    "qb_rng_uniform": (ctypes.c_int, [_P, _I64, _U64, _U64, _P]),
    "qb_rng_normal": (ctypes.c_int, [_P, _I64, _U64, _U64, _P]),
    "qb_mt19937_workspace_bytes": (_SZ, [_I64, _I64]),
    "qb_mt19937_uniform": (ctypes.c_int, [_P, _I32, _I64, _P, _P, ctypes.POINTER(_I32), _P, _SZ, _P]),
    "qb_mt19937_normal": (ctypes.c_int, [_P, _I32, _I32, _F64, _I64, _P, _P, ctypes.POINTER(_I32),
                                         ctypes.POINTER(_I32), ctypes.POINTER(_F64), _P, _SZ, _P]),
}

_lib = None


def library_path():
    return _LIB_PATH


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise QbError(
            "libqinfer_b200.so is not built (%s). Run `python python-qinfer_b200/build.py`; "
            "this engine has no CPU fallback." % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if lib.qb_abi_version() != QB_ABI_VERSION:
        raise QbError("libqinfer_b200.so ABI version %d, expected %d" % (lib.qb_abi_version(), QB_ABI_VERSION))
    sizes = (ctypes.c_int32 * 3)()
    lib.qb_struct_sizes(sizes)
    if list(sizes) != [ctypes.sizeof(QbModel), ctypes.sizeof(QbExpparams), ctypes.sizeof(QbUpdateCtl)]:
        raise QbError("struct layout mismatch between the binding and libqinfer_b200.so: %r" % list(sizes))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().qb_last_error()
        raise QbError("qinfer_b200 call failed (%d): %s" % (rc, msg.decode() if msg else "?"))


def f64_array(values):
    return (ctypes.c_double * len(values))(*[float(v) for v in values])


# ---- NVTX ranges (SURVEY §5 tracing): QB_NVTX=1 wraps the host entry points in named ranges (visible in nsys / ncu
# timelines); unset, the functions are left untouched — no cost on the hot path ------------------------------------
NVTX = os.environ.get("QB_NVTX", "0") == "1"


def nvtx_range(name):
    def deco(fn):
        if not NVTX:
            return fn
        import functools
        import torch

        @functools.wraps(fn)
        def inner(*args, **kwargs):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return inner
    return deco
