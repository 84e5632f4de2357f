"""ctypes binding of libsponet_b200.so (include/sponet_b200.h).

The library is the product: there is no Python or CPU fallback behind it.  Loading fails loudly
when the shared object is missing, and every run call fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPONET_B200_LIB selects another build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("SPONET_B200_LIB") or os.path.join(_HERE, "libsponet_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_STREAM_EXHAUSTED, ERR_UNSUPPORTED, ERR_ZERO_RATE = range(6)
NUM_COUNTERS = 4


class StreamExhausted(RuntimeError):
    """Replay mode: the injected uint64 stream was too short (the caller retries with a longer one)."""


class sponet_cnvm_params(C.Structure):
    _fields_ = [
        ("num_opinions", C.c_int32),
        ("r_imit", C.c_double),
        ("r_noise", C.c_double),
        ("prob_imit", C.c_void_p),
        ("prob_noise", C.c_void_p),
        ("degree_alpha", C.c_void_p),
    ]


class sponet_cntm_params(C.Structure):
    _fields_ = [
        ("next_event_rate", C.c_double),
        ("noise_prob", C.c_double),
        ("threshold_01", C.c_double),
        ("threshold_10", C.c_double),
    ]


class sponet_replay_work(C.Structure):
    _fields_ = [
        ("state_begin", C.c_int64),
        ("state_end", C.c_int64),
        ("run_begin", C.c_int64),
        ("run_end", C.c_int64),
        ("stream_offset", C.c_int64),
        ("stream_len", C.c_int64),
    ]


class sponet_shares_cv(C.Structure):
    _fields_ = [
        ("dimension", C.c_int32),
        ("idx", C.c_void_p),
        ("divisor", C.c_double),
        ("num_classes", C.c_int32),
        ("agent_class", C.c_void_p),
        ("agent_weight", C.c_void_p),
        ("divisors", C.c_void_p),
        ("kind", C.c_int32),
        ("graph", C.c_void_p),
        ("degree_alpha", C.c_void_p),
        ("rates", C.c_double * 4),
    ]


class sponet_native_opts(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("num_runs_total", C.c_int64),
        ("run_begin", C.c_int64),
        ("run_end", C.c_int64),
        ("event_counts", C.c_void_p),
        ("force_global_state", C.c_int32),
        ("warps_per_chain", C.c_int32),
        ("final_state_only", C.c_int32),
        ("hist_bins", C.c_int32),
        ("hist_lo", C.c_int64),
        ("hist_width", C.c_int64),
        ("out_hist", C.c_void_p),
        ("chain_offset", C.c_int64),
    ]


_lock = threading.Lock()
_lib = None

_vp, _i64, _i32, _f64, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_uint64

_SIGNATURES = {
    "sponet_last_error": (C.c_char_p, []),
    "sponet_abi_version": (C.c_int, []),
    "sponet_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sponet_build_alias_table": (C.c_int, [_vp, _i64, _vp, _vp]),
    "sponet_cnvm_event_rates": (C.c_int, [_vp, _i64, _f64, _f64, _f64, C.POINTER(_f64), C.POINTER(_f64)]),
    "sponet_graph_create": (C.c_int, [_i64, _vp, _vp, C.POINTER(_vp)]),
    "sponet_graph_destroy": (C.c_int, [_vp]),
    "sponet_graph_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(C.c_int)]),
    "sponet_graph_neighbor_table": (C.c_int, [_vp, C.POINTER(_i32), _vp]),
    "sponet_graph_batch_create_er": (C.c_int, [_i64, _f64, C.c_int, _u64, _i64, _i64, _i64, _i64, _i32, C.c_int, C.POINTER(_vp)]),
    "sponet_graph_batch_create_ba": (C.c_int, [_i64, _i32, _u64, _i64, _i64, _i64, _i64, C.c_int, C.POINTER(_vp)]),
    "sponet_graph_batch_create_ws": (C.c_int, [_i64, _i32, _f64, _u64, _i64, _i64, _i64, _i64, _i32, C.c_int, C.POINTER(_vp)]),
    "sponet_graph_batch_create_regular": (C.c_int, [_i64, _i32, _u64, _i64, _i64, _i64, _i64, _i32, C.c_int, C.POINTER(_vp)]),
    "sponet_graph_batch_set_weights": (C.c_int, [_vp, _vp, _i64]),
    "sponet_graph_batch_get": (C.c_int, [_vp, _i64, _vp, _vp, _i64]),
    "sponet_graph_batch_trim": (C.c_int, []),
    "sponet_cnvm_replay": (C.c_int, [_vp, _i64, C.POINTER(sponet_cnvm_params), _vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "sponet_cntm_replay": (C.c_int, [_vp, C.POINTER(sponet_cntm_params), _vp, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "sponet_cnvm_replay_all": (C.c_int, [_vp, _i64, C.POINTER(sponet_cnvm_params), _vp, _f64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sponet_cntm_replay_all": (C.c_int, [_vp, C.POINTER(sponet_cntm_params), _vp, _f64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sponet_cnvm_run": (C.c_int, [_vp, _i64, C.POINTER(sponet_cnvm_params), _vp, _i64, _vp, _i64, C.POINTER(sponet_native_opts), C.POINTER(sponet_shares_cv), _vp, _vp, _vp, _vp, _vp, _vp]),
    "sponet_cnvm_run_sweep": (C.c_int, [_vp, _i64, C.POINTER(sponet_cnvm_params), _i64, _vp, _i32, _vp, _i64, C.POINTER(sponet_native_opts), C.POINTER(sponet_shares_cv), _vp, _vp, _vp, _vp, _vp, _vp]),
    "sponet_cnvm_run_complete_counts": (C.c_int, [_i64, C.POINTER(sponet_cnvm_params), _vp, _i64, _vp, _i64, C.POINTER(sponet_native_opts), C.POINTER(sponet_shares_cv), _vp, _vp, _vp, _vp, _vp]),
    "sponet_cntm_run": (C.c_int, [_vp, C.POINTER(sponet_cntm_params), _vp, _i64, _vp, _i64, C.POINTER(sponet_native_opts), C.POINTER(sponet_shares_cv), _vp, _vp, _vp, _vp, _vp, _vp]),
    "sponet_native_event_counts": (C.c_int, [_u64, _i64, _i64, _vp, _i64, _f64, _vp, _vp]),
    "sponet_sample_states": (C.c_int, [_i64, _i32, _vp, _i32, _i64, _u64, _vp, _vp]),
    "sponet_opinion_shares": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "sponet_moments": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "sponet_interfaces": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "sponet_propensities": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _f64, _vp, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises when it has not been built: no fallback."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing. Build it with `python -m sponet_b200.build` "
                    "(needs nvcc); sponet_b200 has no CPU fallback."
                )
            L = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            if L.sponet_abi_version() != 4:
                raise ImportError("libsponet_b200.so: ABI version mismatch, rebuild the library")
            _lib = L
    return _lib


def last_error() -> str:
    return lib().sponet_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map C-ABI return codes onto the exception types the reference raises."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_STREAM_EXHAUSTED:
        raise StreamExhausted(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_ZERO_RATE:
        raise ZeroDivisionError(msg)
    raise RuntimeError(msg)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().sponet_device_count(C.byref(n))
    return n.value if rc == OK else 0


def require_device() -> None:
    n = C.c_int(0)
    check(lib().sponet_device_count(C.byref(n)))


def ptr(a) -> int | None:
    """Raw address of a numpy array or a torch tensor (host or device); None stays NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch.Tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)!r}")


def current_stream_ptr() -> int | None:
    """torch's current CUDA stream as a cudaStream_t (None = default stream)."""
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_stream().cuda_stream or None
    except Exception:
        pass
    return None
