"""ctypes binding of libmdb200.so (include/mdb200.h).

The shared library is the product: CUDA kernels for sm_100a behind a C ABI.  Importing this module
fails loudly when the library has not been built; every compute call fails loudly (MDB_ERR_CUDA)
when no CUDA device is usable.  There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

# The pipeline overlaps chunks on about ten CUDA streams; CUDA maps streams onto
# CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8), read when the process creates its CUDA
# context -- ask for the maximum unless the user chose a value (see csrc/api.cu, mdb_library_loaded).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmdb200.so")

MDB_OK = 0
MDB_ERR_NO_PEAKS_DETECTED = 1
MDB_ERR_EMPTY_SIGNAL_REGION = 2
MDB_ERR_EMPTY_SIGNAL_FREE_REGION = 3
MDB_ERR_INVALID_SMOOTHING_SETTINGS = 4
MDB_ERR_INVALID_SELECTION_SETTINGS = 5
MDB_ERR_INVALID_FITTING_SETTINGS = 6
MDB_ERR_INVALID_IGNORE_REGION = 7
MDB_ERR_EMPTY_DATA = 10
MDB_ERR_DATA_LENGTH_MISMATCH = 11
MDB_ERR_NON_UNIFORM_SPACING = 12
MDB_ERR_INVALID_INTENSITIES = 13
MDB_ERR_INVALID_SIGNAL_BOUNDARIES = 14
MDB_ERR_REFERENCE_PANIC = 100
MDB_ERR_CUDA = 200
MDB_ERR_INVALID_ARGUMENT = 201
MDB_ERR_UNSUPPORTED = 202

MDB_SMOOTHING_IDENTITY, MDB_SMOOTHING_MOVING_AVERAGE = 0, 1
MDB_SELECTION_DETECTOR_ONLY, MDB_SELECTION_NOISE_SCORE_FILTER = 0, 1
MDB_SCORING_MINIMUM_SUM = 0
MDB_FITTING_ANALYTICAL = 0
MDB_MEM_HOST, MDB_MEM_DEVICE = 0, 1
MDB_SUPERPOSITION_EXACT, MDB_SUPERPOSITION_FAST = 0, 1
MDB_FIT_EXACT, MDB_FIT_CORRECTED, MDB_FIT_ULP = 0, 1, 2
ABI_VERSION = 2
KERNEL_NAMES = ["smooth", "detect", "select", "fit_init", "fit_iter", "retain", "mse_superposition",
                "mse_reduce", "superposition_vec", "small_fused"]


class Lorentzian3(C.Structure):
    _fields_ = [("sfhw", C.c_double), ("hw2", C.c_double), ("maxp", C.c_double)]


class SmoothingSettings(C.Structure):
    _fields_ = [("kind", C.c_int32), ("iterations", C.c_uint64), ("window_size", C.c_uint64)]


class SelectionSettings(C.Structure):
    _fields_ = [("kind", C.c_int32), ("scoring_method", C.c_int32), ("threshold", C.c_double)]


class FittingSettings(C.Structure):
    _fields_ = [("kind", C.c_int32), ("iterations", C.c_uint64)]


class SpectrumView(C.Structure):
    _fields_ = [("chemical_shifts", C.c_void_p), ("intensities", C.c_void_p), ("len", C.c_size_t),
                ("signal_boundaries", C.c_double * 2)]


# Every symbol include/mdb200.h declares: (name, restype, argtypes).  tests/test_abi.py checks
# this table against the header and against the built library.
_P = C.c_void_p
_DP = C.POINTER(C.c_double)
SIGNATURES = [
    ("mdb_abi_version", C.c_uint32, []),
    ("mdb_last_error_message", C.c_char_p, []),
    ("mdb_device_count", C.c_int, []),
    ("mdb_set_device_count", C.c_int, [C.c_int]),
    ("mdb_set_superposition_mode", C.c_int, [C.c_int]),
    ("mdb_superposition_mode", C.c_int, []),
    ("mdb_measure_fp64_rate", C.c_int, [_DP, _DP]),
    ("mdb_host_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("mdb_host_free", C.c_int, [_P]),
    ("mdb_release_workspaces", C.c_int, []),
    ("mdb_kernel_launch_count", C.c_uint64, []),
    ("mdb_transfer_bytes", None, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("mdb_reset_kernel_launch_count", None, []),
    ("mdb_profile_enable", None, [C.c_int]),
    ("mdb_profile_reset", None, []),
    ("mdb_profile_read", C.c_int, [C.c_int, _DP, C.POINTER(C.c_uint64), _DP]),
    ("mdb_spectrum_validate", C.c_int, [_P, C.c_size_t, _P, C.c_size_t, _DP, _DP]),
    ("mdb_deconvoluter_default", C.c_int, [C.POINTER(_P)]),
    ("mdb_deconvoluter_new", C.c_int, [C.POINTER(SmoothingSettings), C.POINTER(SelectionSettings),
                                      C.POINTER(FittingSettings), C.POINTER(_P)]),
    ("mdb_deconvoluter_clone", C.c_int, [_P, C.POINTER(_P)]),
    ("mdb_deconvoluter_free", None, [_P]),
    ("mdb_deconvoluter_smoothing_settings", C.c_int, [_P, C.POINTER(SmoothingSettings)]),
    ("mdb_deconvoluter_selection_settings", C.c_int, [_P, C.POINTER(SelectionSettings)]),
    ("mdb_deconvoluter_fitting_settings", C.c_int, [_P, C.POINTER(FittingSettings)]),
    ("mdb_deconvoluter_ignore_regions", C.c_int64, [_P, _DP, C.c_size_t]),
    ("mdb_deconvoluter_set_smoothing_settings", C.c_int, [_P, C.POINTER(SmoothingSettings)]),
    ("mdb_deconvoluter_set_selection_settings", C.c_int, [_P, C.POINTER(SelectionSettings)]),
    ("mdb_deconvoluter_set_fitting_settings", C.c_int, [_P, C.POINTER(FittingSettings)]),
    ("mdb_deconvoluter_add_ignore_region", C.c_int, [_P, C.c_double, C.c_double]),
    ("mdb_deconvoluter_clear_ignore_regions", None, [_P]),
    ("mdb_deconvoluter_set_superposition_mode", C.c_int, [_P, C.c_int]),
    ("mdb_deconvoluter_superposition_mode", C.c_int, [_P]),
    ("mdb_deconvoluter_set_fit_arithmetic", C.c_int, [_P, C.c_int]),
    ("mdb_deconvoluter_fit_arithmetic", C.c_int, [_P]),
    ("mdb_batch_len", C.c_size_t, [_P]),
    ("mdb_batch_status", C.c_int, [_P, C.c_size_t]),
    ("mdb_batch_n_lorentzians", C.c_size_t, [_P, C.c_size_t]),
    ("mdb_batch_lorentzians", C.POINTER(Lorentzian3), [_P, C.c_size_t]),
    ("mdb_batch_mse", C.c_double, [_P, C.c_size_t]),
    ("mdb_batch_n_peaks", C.c_size_t, [_P, C.c_size_t]),
    ("mdb_batch_peaks", C.POINTER(C.c_int32), [_P, C.c_size_t]),
    ("mdb_batch_free", None, [_P]),
    ("mdb_batch_totals", None, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    ("mdb_batch_export", C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    ("mdb_deconvolute_spectra", C.c_int, [_P, C.POINTER(SpectrumView), C.c_size_t, C.c_int, C.POINTER(_P)]),
    ("mdb_deconvoluter_optimize_settings", C.c_int, [_P, C.POINTER(SpectrumView), C.c_int, _DP]),
    ("mdb_superposition_vec", C.c_int, [_P, C.c_size_t, _P, C.c_size_t, _P, C.c_int]),
    ("mdb_superposition_vec_mode", C.c_int, [_P, C.c_size_t, _P, C.c_size_t, _P, C.c_int, C.c_int]),
    ("mdb_stage_smooth", C.c_int, [_P, C.c_size_t, C.c_uint64, C.c_uint64, _P]),
    ("mdb_stage_smooth_batch", C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint64, _P, _DP]),
    ("mdb_stage_detect", C.c_int, [_P, C.c_size_t, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("mdb_stage_select", C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _P, C.c_size_t,
                                  _P, C.c_size_t, C.POINTER(C.c_size_t), _DP]),
    ("mdb_stage_fit", C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, C.c_uint64, _P,
                               C.POINTER(C.c_size_t), _P]),
]

_lib = None


def load():
    """Load libmdb200.so; raises ImportError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C metabodecon_rust_b200/csrc`.  This package has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SIGNATURES:
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.mdb_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: unexpected ABI version {lib.mdb_abi_version()}")
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().mdb_last_error_message()
    return msg.decode("utf-8", "replace") if msg else ""
