"""ctypes binding of the C-ABI library (include/stableavatar_b200.h).

The product path has no CPU fallback: `lib()` raises if `libsa_b200.so` is missing (build it with
`python -m stableavatar_b200.build`), and every wrapper raises `RuntimeError` carrying `sa_last_error()` when an
entry point returns a negative code. Tensors are torch CUDA tensors used only as device memory + the current stream.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libsa_b200.so"
_lib = None

SA_BF16, SA_F32 = 0, 1
ACT_NONE, ACT_GELU_TANH, ACT_SILU, ACT_GELU_ERF = 0, 1, 2, 3
RES_NONE, RES_ADD, RES_GATED = 0, 1, 2


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p), ("bias", C.c_void_p),
        ("res", C.c_void_p), ("gate", C.c_void_p),
        ("lda", C.c_int64), ("ldw", C.c_int64), ("ldc", C.c_int64), ("ldr", C.c_int64), ("gate_ld", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias_dtype", C.c_int32), ("out_dtype", C.c_int32), ("res_dtype", C.c_int32),
        ("act", C.c_int32), ("res_mode", C.c_int32), ("round_y", C.c_int32),
        ("rows_per_batch", C.c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_bs", C.c_int64), ("q_ls", C.c_int64), ("k_bs", C.c_int64), ("k_ls", C.c_int64),
        ("v_bs", C.c_int64), ("v_ls", C.c_int64), ("o_bs", C.c_int64), ("o_ls", C.c_int64),
        ("batch", C.c_int32), ("heads", C.c_int32), ("q_len", C.c_int32), ("kv_len", C.c_int32),
        ("scale", C.c_float), ("accumulate", C.c_int32),
    ]


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load libsa_b200.so (once). Fails loudly if it was not built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} not found: build the CUDA library first (python -m stableavatar_b200.build). "
                "stableavatar_b200 has no CPU / eager fallback.")
        l = C.CDLL(str(_LIB_PATH))
        l.sa_last_error.restype = C.c_char_p
        l.sa_version.restype = C.c_int
        _lib = l
    return _lib


launch_count = 0  # kernels of this library launched through the wrappers (bench.py reports it as gpu_launches)


def check(rc: int, what: str) -> None:
    global launch_count
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().sa_last_error().decode()}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return SA_BF16
    if t.dtype == torch.float32:
        return SA_F32
    raise TypeError(f"unsupported dtype {t.dtype}")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()
