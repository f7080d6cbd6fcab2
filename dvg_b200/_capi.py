"""ctypes binding of libdvg_b200.so (include/dvg_b200.h).  No torch types cross this boundary:
device pointers are passed as integers (``tensor.data_ptr()``) and the stream as ``cudaStream_t``.

The library is REQUIRED: if it is missing it is built with nvcc; if that fails the import raises.
There is no CPU or PyTorch fallback for the hot path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint8, c_void_p

from . import build as _build

DVG_FP32, DVG_BF16X3, DVG_BF16 = 0, 1, 2
DVG_LSTM, DVG_GAUSSIAN_LSTM = 0, 1
VARIANTS = {"fp32": DVG_FP32, "bf16x3": DVG_BF16X3, "bf16": DVG_BF16}

EXPORTS = [
    "dvg_last_error", "dvg_version", "dvg_device_info",
    "dvg_lstm_prepare", "dvg_lstm_refresh", "dvg_lstm_destroy", "dvg_lstm_reserve",
    "dvg_lstm_chain_begin", "dvg_lstm_chain_end",
    "dvg_lstm_state_bytes", "dvg_lstm_state_packed_offset", "dvg_lstm_state_repack",
    "dvg_lstm_step", "dvg_lstm_profile", "dvg_gauss_lstm_step",
    "dvg_gp_prepare", "dvg_gp_refresh", "dvg_gp_prepare_factors", "dvg_gp_refresh_factors", "dvg_gp_factorize", "dvg_gp_factorize_workspace", "dvg_gp_destroy", "dvg_gp_predict", "dvg_gp_trigger",
    "dvg_gp_rsample", "dvg_gp_export", "dvg_rollout_step", "dvg_eval_seq_finn", "dvg_eval_seq", "dvg_rollout_score",
    "dvg_moving_mnist_draws", "dvg_moving_mnist",
]


class DvgError(RuntimeError):
    pass


class LstmDims(ctypes.Structure):
    _fields_ = [("kind", c_int), ("input_size", c_int), ("hidden_size", c_int), ("n_layers", c_int),
                ("output_size", c_int)]


class GpDims(ctypes.Structure):
    _fields_ = [("num_dims", c_int), ("num_inducing", c_int), ("jitter", c_float), ("noise_lower_bound", c_float)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) libdvg_b200.so.  Raises DvgError when unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not _build.is_current() and not os.environ.get("DVG_LIB_NOREBUILD"):     # (developer A/B of pre-built tagged libraries)
        try:
            _build.build(verbose=False)
        except Exception as e:  # stale-but-present library is still usable (e.g. no nvcc on the box)
            if not os.path.exists(_build.LIB):
                raise DvgError(f"libdvg_b200.so is missing and could not be built: {e}") from e
    try:
        lib = ctypes.CDLL(_build.LIB)
    except OSError as e:
        raise DvgError(f"cannot load {_build.LIB}: {e}") from e
    P = c_void_p
    PP = POINTER(c_void_p)
    lib.dvg_last_error.restype = c_char_p
    lib.dvg_last_error.argtypes = []
    lib.dvg_version.restype = c_int
    lib.dvg_device_info.argtypes = [POINTER(c_int)] * 3
    lib.dvg_lstm_prepare.argtypes = [POINTER(c_void_p), POINTER(LstmDims), P, P, PP, PP, PP, PP, P, P, P, P, P]
    lib.dvg_lstm_refresh.argtypes = [P, P, P, PP, PP, PP, PP, P, P, P, P, P]
    lib.dvg_lstm_destroy.argtypes = [P]
    lib.dvg_lstm_reserve.argtypes = [P, c_int]
    lib.dvg_lstm_chain_begin.argtypes = [P, P]
    lib.dvg_lstm_chain_end.argtypes = [P, P]
    lib.dvg_lstm_state_bytes.restype = c_size_t
    lib.dvg_lstm_state_bytes.argtypes = [P, c_int]
    lib.dvg_lstm_state_packed_offset.restype = c_size_t
    lib.dvg_lstm_state_packed_offset.argtypes = [P, c_int]
    lib.dvg_lstm_state_repack.argtypes = [P, c_int, P, P]
    lib.dvg_lstm_step.argtypes = [P, c_int, c_int, P, c_int, P, P, P, c_int, P, c_int, P]
    lib.dvg_lstm_profile.argtypes = [P, c_int, c_int, P, c_int, P, P, P, c_int, POINTER(c_float), c_int, P]
    lib.dvg_gauss_lstm_step.argtypes = [P, c_int, c_int, P, c_int, P, P, P, P, P, P, P]
    lib.dvg_gp_prepare.argtypes = [POINTER(c_void_p), POINTER(GpDims), P, P, P, P, P, P, P, P]
    lib.dvg_gp_refresh.argtypes = [P, P, P, P, P, P, P, P, P]
    lib.dvg_gp_destroy.argtypes = [P]
    lib.dvg_gp_predict.argtypes = [P, c_int, P, c_int, P, P, c_int, P, c_int, P]
    lib.dvg_gp_prepare_factors.argtypes = [POINTER(c_void_p), POINTER(GpDims), P, P, P, P, P, P]
    lib.dvg_gp_factorize.argtypes = [POINTER(GpDims), P, P, P, P, P, P, P, P, c_size_t, P]
    lib.dvg_gp_factorize_workspace.restype = c_size_t
    lib.dvg_gp_factorize_workspace.argtypes = [POINTER(GpDims), c_int]
    lib.dvg_gp_refresh_factors.argtypes = [P, P, P, P, P, P, P]
    lib.dvg_gp_trigger.argtypes = [P, c_int, P, c_int, P, P, c_int, P, c_int, c_float, P, P, P, P]
    lib.dvg_gp_rsample.argtypes = [P, c_int, c_int, P, c_int, P, P, P, c_int, P]
    lib.dvg_gp_export.argtypes = [P, P, P, P, P, P]
    lib.dvg_rollout_step.argtypes = [P, P, c_int, c_int, P, c_int, P, P, P, c_int, c_int, P, P, c_int, P, c_int, c_float,
                                     P, P, P, P, P]
    lib.dvg_eval_seq_finn.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P]
    lib.dvg_eval_seq.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P]
    lib.dvg_rollout_score.argtypes = [c_int, c_int, c_int, c_int, P, P, P, P]
    lib.dvg_moving_mnist_draws.argtypes = [c_int, c_int]
    lib.dvg_moving_mnist.argtypes = [c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_int, P, P, P]
    for name in EXPORTS:
        fn = getattr(lib, name)  # raises AttributeError if a declared symbol is not exported
        if name not in ("dvg_last_error", "dvg_lstm_state_bytes", "dvg_lstm_state_packed_offset", "dvg_gp_factorize_workspace"):
            fn.restype = c_int
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().dvg_last_error().decode(errors="replace")
        raise DvgError(f"{what or 'dvg call'} failed (status {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return ctypes.cast(arr, POINTER(c_void_p)), arr


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
