"""ctypes binding of ``libbigcodec_b200.so`` (the C ABI declared in include/bigcodec_b200.h).

PyTorch is used by the callers only for device memory and streams; what crosses this
boundary is raw device pointers, ints and a ``cudaStream_t``.  There is no fallback:
if the library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_size_t, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BC_LIB_PATH") or os.path.join(_HERE, "libbigcodec_b200.so")   # BC_LIB_PATH: A/B builds of the same ABI

BC_CONV_SNAKE_IN = 1
BC_CONV_TANH_OUT = 2

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3}

# every symbol include/bigcodec_b200.h declares (checked by tests/test_cabi_symbols.py)
_SIGNATURES = {
    "bc_abi_version": (c_int, []),
    "bc_last_error": (c_char_p, []),
    "bc_policy": (c_int, [c_char_p, c_size_t]),
    "bc_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "bc_transpose_bct_to_btc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "bc_transpose_btc_to_bct": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "bc_snake_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "bc_conv1d_fwd": (c_int, [c_void_p] * 7 + [c_int] * 14 + [c_void_p]),
    "bc_tc_plan": (c_int, [c_int] * 6 + [POINTER(c_int)] * 3),
    "bc_resunit_plan": (c_int, [c_int] * 4 + [POINTER(c_int)] * 4),
    "bc_resunit_fwd": (c_int, [c_void_p] * 10 + [c_int] * 7 + [c_void_p]),
    "bc_stream_plan": (c_int, [c_int] * 7 + [POINTER(c_int)]),
    "bc_conv1d_stream_fwd": (c_int, [c_void_p] * 7 + [c_int] * 11 + [c_void_p]),
    "bc_resunit_stream_fwd": (c_int, [c_void_p] * 10 + [c_int] * 7 + [c_void_p]),
    "bc_stream_pair_ok": (c_int, [c_int] * 7),
    "bc_conv1d_stream_pair_fwd": (c_int, [c_void_p] * 7 + [c_int] * 11 + [c_void_p]),
    "bc_resunit_stream_pair_fwd": (c_int, [c_void_p] * 10 + [c_int] * 7 + [c_void_p]),
    "bc_convtr1d_stream_pair_fwd": (c_int, [c_void_p] * 6 + [c_int] * 8 + [c_void_p]),
    "bc_convtr1d_stream_fwd": (c_int, [c_void_p] * 6 + [c_int] * 8 + [c_void_p]),
    "bc_convtr1d_fwd": (c_int, [c_void_p] * 6 + [c_int] * 8 + [c_void_p]),
    "bc_lstm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "bc_lstm_packed_whh_floats": (c_size_t, [c_int]),
    "bc_lstm_pack_whh": (c_int, [c_void_p, c_void_p, c_int]),
    "bc_lstm_recurrent_fwd": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "bc_lstm_tc_slice_cols": (c_int, [c_int]),
    "bc_lstm_tc_max_batch": (c_int, [c_int, c_int]),
    "bc_lstm_tc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bc_lstm_tc_recurrent_fwd": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "bc_lstm_tc_recurrent_chunk_fwd": (c_int, [c_void_p] * 6 + [c_int] * 7 + [c_void_p]),
    "bc_lstm_tc_ctas": (c_int, [c_int, c_int, c_int]),
    "bc_vq_encode": (c_int, [c_void_p] * 7 + [c_int] * 4 + [c_void_p]),
    "bc_vq_dequant": (c_int, [c_void_p] * 7 + [c_int] * 5 + [c_void_p]),
    "bc_fsq_encode": (c_int, [c_void_p] * 7 + [c_int] * 3 + [c_void_p]),
    "bc_ipc_alloc": (c_int, [POINTER(c_void_p), c_size_t]),
    "bc_ipc_free": (c_int, [c_void_p]),
    "bc_ipc_export": (c_int, [c_void_p, c_char_p]),
    "bc_ipc_open": (c_int, [c_char_p, POINTER(c_void_p)]),
    "bc_ipc_close": (c_int, [c_void_p]),
    "bc_peer_copy": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "bc_code_histogram": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    "bc_code_entropy": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "bc_debug_set_ru_trace": (c_int, [c_void_p]),
    "bc_debug_set_stream_trace": (c_int, [c_void_p]),
    "bc_debug_set_lstm_trace": (c_int, [c_void_p]),
    "bc_indices_to_int16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

_lib = None


def load_library() -> ctypes.CDLL:
    """Load the CUDA library; raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m audiotokenization_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.bc_abi_version() != 4:
        raise RuntimeError(f"stale {LIB_PATH}: ABI {lib.bc_abi_version()} != 4; rebuild")
    _lib = lib
    return lib


def policy() -> str:
    """The kernel-selection knobs the library read from the environment (once, at first use)."""
    buf = ctypes.create_string_buffer(256)
    check(load_library().bc_policy(buf, 256), "bc_policy")
    return buf.value.decode()


def last_error() -> str:
    return load_library().bc_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the BigCodec hot path runs only on a CUDA (sm_100a) device; "
            "there is no CPU fallback")


def device_info(dev: int = 0):
    lib = load_library()
    sm, maj, mnr, mem = c_int(), c_int(), c_int(), c_size_t()
    rc = lib.bc_device_info(dev, ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr), ctypes.byref(mem))
    check(rc, "bc_device_info")
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "total_mem": mem.value}
