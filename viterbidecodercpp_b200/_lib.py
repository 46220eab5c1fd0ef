"""Loader for the in-tree CUDA library (csrc/libviterbi_b200.so).  There is no fallback: if the library is missing the import
of anything that computes fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libviterbi_b200.so")
VITB_MAX_R = 16

VITB_END_STATE_BEST = 2 ** 64 - 1     # (size_t)-1
VITB_OK, VITB_ERR_ARG, VITB_ERR_UNSUPPORTED, VITB_ERR_CUDA, VITB_ERR_STATE, VITB_ERR_NOMEM = 0, -1, -2, -3, -4, -5
VITB_TIE_SCALAR, VITB_TIE_SIMD, VITB_TIE_SIMD_SAT = 0, 1, 2


class vitb_params(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("R", C.c_int32), ("G", C.c_uint32 * VITB_MAX_R), ("soft_bytes", C.c_int32),
        ("soft_decision_high", C.c_int32), ("soft_decision_low", C.c_int32),
        ("soft_decision_max_error", C.c_uint32), ("initial_start_error", C.c_uint32),
        ("initial_non_start_error", C.c_uint32), ("renormalisation_threshold", C.c_uint32),
        ("tie_break", C.c_int32), ("device", C.c_int32),
    ]


class vitb_batch_opts(C.Structure):
    _fields_ = [("row_stride", C.c_size_t), ("starting_state", C.c_size_t), ("end_state", C.c_size_t)]


# every symbol include/viterbi_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
API = [
    ("vitb_create", C.c_int, [C.POINTER(vitb_params), C.POINTER(_P)]),
    ("vitb_destroy", C.c_int, [_P]),
    ("vitb_is_supported", C.c_int, [C.POINTER(vitb_params)]),
    ("vitb_set_traceback_length", C.c_int, [_P, C.c_size_t]),
    ("vitb_get_traceback_length", C.c_int, [_P, C.POINTER(C.c_size_t)]),
    ("vitb_reset", C.c_int, [_P, C.c_size_t]),
    ("vitb_update", C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_uint64)]),
    ("vitb_get_error", C.c_int, [_P, C.c_size_t, C.POINTER(C.c_uint32)]),
    ("vitb_chainback", C.c_int, [_P, _P, C.c_size_t, C.c_size_t]),
    ("vitb_get_current_decoded_bit", C.c_int, [_P, C.POINTER(C.c_size_t)]),
    ("vitb_get_metrics", C.c_int, [_P, _P]),
    ("vitb_set_metrics", C.c_int, [_P, _P]),
    ("vitb_set_current_decoded_bit", C.c_int, [_P, C.c_size_t]),
    ("vitb_get_decisions", C.c_int, [_P, C.c_size_t, C.c_size_t, _P]),
    ("vitb_decode_batch", C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.POINTER(vitb_batch_opts), _P, _P, _P]),
    ("vitb_decode_batch_dev", C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.POINTER(vitb_batch_opts), _P, _P, _P, _P]),
    ("vitb_set_traceback_window", C.c_int, [_P, C.c_size_t]),
    ("vitb_get_window_mismatches", C.c_int, [_P, C.POINTER(C.c_uint64)]),
    ("vitb_set_pipelining", C.c_int, [_P, C.c_int]),
    ("vitb_batch_flush", C.c_int, [_P, _P]),
    ("vitb_decode_batch_async", C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.POINTER(vitb_batch_opts), _P, _P, _P, _P]),
    ("vitb_set_puncture_schedule", C.c_int, [_P, _P, C.c_size_t, C.c_int32]),
    ("vitb_synth_frames_dev", C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_float, C.c_uint64, _P, _P, C.c_size_t, _P]),
    ("vitb_synth_frames", C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_float, C.c_uint64, _P, _P, C.c_size_t]),
    ("vitb_quantise", C.c_int, [_P, _P, C.c_size_t, C.c_float, C.c_float, _P]),
    ("vitb_ber_trial", C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_float, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("vitb_decode_batch_multi", C.c_int, [C.POINTER(_P), C.c_int, _P, C.c_size_t, C.c_size_t, C.POINTER(vitb_batch_opts), _P, _P, _P]),
    ("vitb_workspace_bytes", C.c_int, [_P, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("vitb_set_workspace_limit", C.c_int, [_P, C.c_size_t]),
    ("vitb_kernel_launch_count", C.c_int, [_P, C.POINTER(C.c_uint64)]),
    ("vitb_set_profiling", C.c_int, [_P, C.c_int]),
    ("vitb_get_stage_ms", C.c_int, [_P, C.POINTER(C.c_float * 4)]),
    ("vitb_set_variant", C.c_int, [_P, C.c_int]),
    ("vitb_set_history_kernel", C.c_int, [_P, C.c_int]),
    ("vitb_set_traceback_segments", C.c_int, [_P, C.c_int, C.c_int]),
    ("vitb_get_variants", C.c_int, [_P, C.POINTER(C.c_int), C.c_int]),
    ("vitb_kernel_name", C.c_char_p, [_P]),
    ("vitb_last_cuda_error", C.c_int, [_P]),
    ("vitb_status_string", C.c_char_p, [C.c_int]),
    ("vitb_device_count", C.c_int, []),
    ("vitb_version", C.c_char_p, []),
]

_lib = None


def load():
    """dlopen the CUDA library and declare every entry point.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C viterbidecodercpp_b200/csrc` (or __graft_entry__.build()); "
                "there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, restype, argtypes in API:
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib
