"""ctypes binding of libbutd_b200.so (the C-ABI declared in include/butd_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, this
raises.  PyTorch is used only for device memory and streams — the tensors' `data_ptr()`s and
the current CUDA stream are what crosses the boundary.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbutd_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_LL = ctypes.c_longlong

# name -> argtypes (stream is always last and always void*)
_SIGNATURES = {
    "bd_fps": [_P, _I, _I, _I, _I, _P, _P, _P],
    "bd_fps_set_cluster": [_I],
    "bd_gather_points": [_P, _P, _I, _I, _I, _I, _P, _P],
    "bd_gather_points_grad": [_P, _P, _I, _I, _I, _I, _P, _P],
    "bd_ball_query": [_P, _P, _I, _I, _I, _I, _F, _I, _P, _P],
    "bd_group_points": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "bd_group_points_grad": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "bd_three_nn": [_P, _P, _I, _I, _I, _P, _P, _P],
    "bd_three_interpolate": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "bd_three_interpolate_grad": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "bd_gather_rows": [_P, _I, _P, _I, _I, _I, _I, _P, _I, _P],
    "bd_group_rows": [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P],
    "bd_maxpool_rows": [_P, _I, _I, _I, _P, _P],
    "bd_fp_interp_concat": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P],
    "bd_linear_f32": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "bd_add_layernorm_f32": [_P, _P, _P, _P, _P, _I, _I, _F, _P],
    "bd_attention_f32": [_P, _I, _LL, _P, _I, _LL, _P, _I, _LL, _P, _P, _I, _LL, _I, _I, _I, _I, _I, _F, _P],
    "bd_topk_sigmoid": [_P, _I, _I, _I, _P, _P],
    "bd_l2_normalize_rows": [_P, _P, _I, _I, _P],
    "bd_embedding_rows": [_P, _I, _P, _I, _P, _I, _P],
    "bd_add_rows": [_P, _I, _P, _I, _P, _I, _I, _I, _P],
    "bd_transpose_rows": [_P, _I, _I, _I, _P, _P],
    "bd_concat_rows": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _P],
}

EXPORTED = sorted(list(_SIGNATURES) + ["bd_version", "bd_last_error", "bd_arch", "bd_fps_resident_capacity"])

_lib = None
launch_count = 0  # kernels enqueued through this binding (bench.py reports it as gpu_launches)


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m butd_detr_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _I
    lib.bd_version.restype = _I
    lib.bd_last_error.restype = ctypes.c_char_p
    lib.bd_arch.restype = ctypes.c_char_p
    lib.bd_fps_resident_capacity.restype = _I
    _lib = lib
    return lib


def last_error():
    return load().bd_last_error().decode()


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke `name(*args, current_stream)`; raise RuntimeError with the library's message on failure."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args, stream_ptr())
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {lib.bd_last_error().decode()}")


def check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("CPU not supported: butd_detr_b200 runs on CUDA tensors only")
