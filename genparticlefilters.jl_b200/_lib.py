"""ctypes binding of libgenpf_cuda.so (the C ABI declared in include/genpf.h).

This is the same ABI Julia's ``ccall`` binds.  There is no CPU
fallback anywhere in this package: if the shared library is missing or a CUDA device
is unavailable every compute call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GENPF_LIBRARY overrides the in-tree build (kernel-variant experiments); there is still no fallback
LIB_PATH = os.environ.get("GENPF_LIBRARY") or os.path.join(_HERE, "libgenpf_cuda.so")

# status codes / enums (include/genpf.h)
OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_INVALID_WEIGHTS, ERR_UNKNOWN_METHOD = -1, -2, -3, -4
ERR_NOMEM, ERR_UNSUPPORTED, ERR_STATE, ERR_ASSERT = -5, -6, -7, -8
MULTINOMIAL, RESIDUAL, STRATIFIED = 0, 1, 2
SORT_PARTICLES, SUBSTATE, INDEX_BASE1, DEVICE_PTRS, CHECK, UNIFORMS_STRATA = 1, 2, 4, 8, 16, 32
VALID, INV_NAN_INPUT, INV_ALL_NEGINF, INV_ZERO_TOTAL, INV_NAN_TOTAL = 0, 1, 2, 3, 4
LAYOUT_CONTIGUOUS, LAYOUT_INTERLEAVED = 0, 1
KEEPFIRST, SAMPLE = 0, 1
PRIO_NONE, PRIO_SCALE, PRIO_COLUMN = 0, 1, 2
NOISE_LEAN, NOISE_PHILOX53, KEEP_HISTORY = 0, 1, 2
RUN_GRAPH = 1

METHODS = {"multinomial": MULTINOMIAL, "residual": RESIDUAL, "stratified": STRATIFIED}
# the exact @warn texts of safe_softmax, utils.jl:120,124,132,135
WARNINGS = {
    INV_NAN_INPUT: "NaN found in input values. Returning NaN weights.",
    INV_ALL_NEGINF: "All input values are -Inf. Returning uniform weights.",
    INV_ZERO_TOTAL: "All weights are zero. Returning uniform weights.",
    INV_NAN_TOTAL: "Total weight is NaN. Returning NaN weights.",
}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p
i32, i64, u32, u64, f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double

# every symbol include/genpf.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "genpf_version": (i32, []),
    "genpf_last_error": (C.c_char_p, []),
    "genpf_device_count": (i32, [_i32p]),
    "genpf_set_device": (i32, [i32]),
    "genpf_synchronize": (i32, []),
    "genpf_logsumexp": (i32, [_vp, i64, u32, _dp]),
    "genpf_ess": (i32, [_vp, i64, u32, _dp]),
    "genpf_normalize": (i32, [_vp, i64, u32, _vp, _vp, _dp, _dp, _i32p]),
    "genpf_resample": (i32, [i32, _vp, _vp, i64, i64, _vp, u64, u32, _vp, _vp, _dp, _i32p]),
    "genpf_resample_segmented": (i32, [i32, _vp, _vp, i64, _ip, i64, _vp, u64, u32, _vp, _vp, _i32p]),
    "genpf_weighted_mean_var": (i32, [_vp, _vp, i64, u32, _dp, _dp]),
    "genpf_replicate_host": (i32, [_vp, i64, i64, i32, u32, _vp, _vp]),
    "genpf_dereplicate_host": (i32, [_vp, i64, i64, i32, i32, _vp, u64, u32, _vp, _vp]),
    "genpf_coalesce_host": (i32, [_vp, _vp, i64, u32, _vp, _vp, _ip]),
    "genpf_proportionmap_host": (i32, [_vp, _vp, i64, u32, _vp, _vp, _ip]),
    "genpf_optimal_resize": (i32, [_vp, i64, i64, _dp, u64, u32, _vp, _vp, _ip, _dp, _i32p]),
    "genpf_uniforms": (i32, [u64, u64, i64, u32, _vp]),
    "genpf_debug_cumweights": (i32, [_vp, i64, u32, _vp]),
    "genpf_debug_sortperm": (i32, [_vp, i64, u32, _vp]),
    "genpf_model_builtin": (i32, [C.c_char_p, _i32p]),
    "genpf_model_info": (i32, [i32, _i32p, _i32p, _i32p, _i32p]),
    "genpf_model_caps": (i32, [i32, _i32p]),
    "genpf_model_compile": (i32, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, _i32p]),
    "genpf_model_export": (i32, [i32, _vp, i64, _ip]),
    "genpf_model_load_image": (i32, [_vp, i64, _i32p]),
    "genpf_model_plugin_sources": (i32, [i32, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]),
    "genpf_initialize_proposal": (i32, [_vp, _vp, _vp, _vp, _vp]),
    "genpf_update_proposal": (i32, [_vp, i64, _vp, _vp, _vp, _vp]),
    "genpf_update_translate": (i32, [_vp, i64, _vp, _vp, _vp, _vp]),
    "genpf_rejuvenate_reweight_proposal": (i32, [_vp, i64, _vp, _vp, i32, _vp, _vp]),
    "genpf_filter_create": (i32, [i32, _vp, i32, i64, i64, u64, u32, C.POINTER(_vp)]),
    "genpf_filter_destroy": (i32, [_vp]),
    "genpf_filter_size": (i32, [_vp, _ip, _ip]),
    "genpf_initialize": (i32, [_vp, _vp, _vp]),
    "genpf_initialize_with_noise": (i32, [_vp, _vp, _vp, _vp, _vp]),
    "genpf_initialize_stratified": (i32, [_vp, _vp, _vp, i32, _vp, i32, i32, _vp, _vp]),
    "genpf_update_stratified": (i32, [_vp, i64, _vp, _vp, i32, _vp, i32, i32, _vp, _vp]),
    "genpf_update": (i32, [_vp, i64, _vp, _vp]),
    "genpf_update_with_noise": (i32, [_vp, i64, _vp, _vp, _vp, _vp]),
    "genpf_ess_dev": (i32, [_vp, _vp]),
    "genpf_lml_dev": (i32, [_vp, _vp]),
    "genpf_resample_dev": (i32, [_vp, i32, i32, f64, _vp, i64, u32, _vp, _vp]),
    "genpf_rejuvenate_mh": (i32, [_vp, i64, _vp, _vp, i32, _vp]),
    "genpf_rejuvenate_mh_with_noise": (i32, [_vp, i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "genpf_rejuvenate_reweight": (i32, [_vp, i64, _vp, _vp, i32]),
    "genpf_rejuvenate_reweight_with_noise": (i32, [_vp, i64, _vp, _vp, _vp, _vp]),
    "genpf_step": (i32, [_vp, i64, _vp, _vp, _vp, _vp, i32, f64, i32, _vp]),
    "genpf_run_steps": (i32, [_vp, i64, i64, _vp, _vp, i32, f64, i32, u32]),
    "genpf_introduce": (i32, [_vp, i64, _vp, _vp, i32, _vp, _vp]),
    "genpf_filter_set_first_filter": (i32, [_vp, i64]),
    "genpf_step_with_noise": (i32, [_vp, i64, _vp, _vp, _vp, _vp, i32, f64, i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "genpf_mean_var": (i32, [_vp, i32, i64, _vp, _vp]),
    "genpf_replicate": (i32, [_vp, i64, i32]),
    "genpf_dereplicate": (i32, [_vp, i64, i32, i32, _vp]),
    "genpf_coalesce": (i32, [_vp, _ip]),
    "genpf_filter_get_progress": (i32, [_vp, _ip, _ip, _vp]),
    "genpf_filter_set_progress": (i32, [_vp, i64, i64, _vp]),
    "genpf_proportionmap": (i32, [_vp, i32, i64, _dp, _dp, i64, _ip]),
    "genpf_optimal_resize_dev": (i32, [_vp, i64, _dp, u32, _ip, _dp, _i32p]),
    "genpf_get_log_weights": (i32, [_vp, _vp]),
    "genpf_set_log_weights": (i32, [_vp, _vp]),
    "genpf_get_parents": (i32, [_vp, _vp, u32]),
    "genpf_get_field": (i32, [_vp, i32, i64, _vp]),
    "genpf_set_field": (i32, [_vp, i32, i64, _vp]),
    "genpf_get_accepts": (i32, [_vp, _vp]),
    "genpf_filter_sync": (i32, [_vp]),
    "genpf_launch_count": (i64, []),
    "genpf_filter_stream": (i32, [_vp, C.POINTER(_vp)]),
    "genpf_shard_ipc_export": (i32, [_vp, _vp, _ip]),
    "genpf_shard_attach": (i32, [_vp, i32, i32, _vp, _vp, _vp, _vp, _vp]),
    "genpf_shard_detach": (i32, [_vp]),
    "genpf_shard_initialize": (i32, [_vp, _vp, _vp]),
    "genpf_shard_begin_step": (i32, [_vp]),
    "genpf_shard_scan": (i32, [_vp]),
    "genpf_shard_push": (i32, [_vp, i64, _vp, _vp, _vp, _vp, i32]),
    "genpf_shard_finish": (i32, [_vp]),
    "genpf_shard_step_p2p": (i32, [_vp, i64, _vp, _vp, _vp, _vp, i32]),
    "genpf_shard_step_p2p_with_noise": (i32, [_vp, i64, _vp, _vp, _vp, _vp, i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "genpf_shard_oend": (i32, [_vp, _vp, _i32p]),
    "genpf_shard_stats": (i32, [_vp, _dp, _dp, _i32p]),
    "genpf_profile_begin": (i32, []),
    "genpf_profile_end": (i32, [C.c_char_p, i64]),
}

_lib = None


class GenPFError(RuntimeError):
    """Raised for every non-zero status; mirrors Julia's ErrorException on this path."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def load():
    """Load libgenpf_cuda.so, binding every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != OK:
        msg = load().genpf_last_error().decode("utf-8", "replace")
        raise GenPFError(status, msg or f"genpf status {status}")


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(_vp)
