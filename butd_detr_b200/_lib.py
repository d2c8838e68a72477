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
    "bd_set_pdl": [_I],
    "bd_linear_tc_set_occupancy": [_I],
    "bd_fps_grid": [_P, _I, _I, _I, _I, _P, _P, _P, _P],
    "bd_fps_grid_set_warps": [_I],
    "bd_fps_grid_stats": [_I, _P],
    "bd_attention_tc_set_small_nk": [_I],
    "bd_attention_tc_set_direct": [_I],
    "bd_attention_tc_set_short": [_I],
    "bd_linear_stream_set": [_I],
    "bd_matcher_cost": [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _F, _F, _F, _P, _P],
    "bd_hungarian": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "bd_linear_smallk": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bd_roberta_embed": [_P, _P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "bd_attention_tc_pack_kv": [_P, _I, _LL, _P, _I, _LL, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "bd_attention_tc_packed": [_P, _I, _LL, _P, _P, _I, _LL, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P],
    "bd_attention_tc_h": [_P, _I, _LL, _P, _I, _LL, _P, _I, _LL, _P, _P, _I, _LL, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P],
    "bd_grid_build": [_P, _I, _I, _I, _F, _P, _P],
    "bd_ball_query_grid_query": [_P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P],
    "bd_gather_points": [_P, _P, _I, _I, _I, _I, _P, _P],
    "bd_gather_points_grad": [_P, _P, _I, _I, _I, _I, _P, _P],
    "bd_ball_query": [_P, _P, _I, _I, _I, _I, _F, _I, _P, _P],
    "bd_ball_query_grid": [_P, _P, _I, _I, _I, _I, _F, _I, _P, _P, _P],
    "bd_group_points": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "bd_group_points_grad": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "bd_three_nn": [_P, _P, _I, _I, _I, _P, _P, _P],
    "bd_three_interpolate": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "bd_three_interpolate_grad": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "bd_gather_rows": [_P, _I, _P, _I, _I, _I, _I, _P, _I, _P],
    "bd_group_rows": [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P],
    "bd_maxpool_rows": [_P, _I, _I, _I, _P, _P],
    "bd_fp_interp_concat": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P],
    "bd_fp_interp_concat_h": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P],
    "bd_linear_f32": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "bd_linear_tc": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bd_linear_ln_tc": [_P, _I, _P, _I, _P, _P, _P, _I, _P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bd_linear_tc_h": [_P, _I, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bd_linear_ln_tc_h": [_P, _I, _I, _P, _P, _P, _I, _P, _P, _F, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bd_linear_tc_set_debug": [_P],
    "bd_sa_group_linear_tc": [_P, _P, _I, _I, _P, _I, _P, _I, _I, _I, _I, _F, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "bd_sa_mlp_tc_h": [_P, _P, _I, _I, _I, _P, _I, _P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _P, _I, _P, _P, _I, _P, _I, _P, _I, _I, _P],
    "bd_sa_mlp_tc": [_P, _P, _I, _I, _P, _I, _P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _P, _I, _P, _P, _I, _P, _I, _I, _P],
    "bd_linear_pool_tc": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "bd_add_layernorm_f32": [_P, _P, _P, _P, _P, _I, _I, _F, _P],
    "bd_attention_f32": [_P, _I, _LL, _P, _I, _LL, _P, _I, _LL, _P, _P, _I, _LL, _I, _I, _I, _I, _I, _F, _P],
    "bd_attention_tc": [_P, _I, _LL, _P, _I, _LL, _P, _I, _LL, _P, _P, _I, _LL, _I, _I, _I, _I, _I, _F, _I, _P, _P],
    "bd_topk_sigmoid": [_P, _I, _I, _I, _P, _P],
    "bd_l2_normalize_rows": [_P, _P, _I, _I, _P],
    "bd_embedding_rows": [_P, _I, _P, _I, _P, _I, _P],
    "bd_add_rows": [_P, _I, _P, _I, _P, _I, _I, _I, _P],
    "bd_transpose_rows": [_P, _I, _I, _I, _P, _P],
    "bd_concat_rows": [_P, _I, _I, _P, _I, _I, _P, _I, _I, _P],
}

EXPORTED = sorted(list(_SIGNATURES) + ["bd_version", "bd_last_error", "bd_arch", "bd_fps_resident_capacity",
                                            "bd_fps_grid_capacity", "bd_fps_grid_scratch_bytes",
                                            "bd_attention_tc_workspace_bytes", "bd_ball_query_grid_workspace_bytes",
                                            "bd_attention_tc_select", "bd_grid_order"])

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
    lib.bd_fps_grid_capacity.restype = _I
    lib.bd_fps_grid_scratch_bytes.restype = _LL
    lib.bd_fps_grid_scratch_bytes.argtypes = [_I, _I]
    lib.bd_ball_query_grid_workspace_bytes.restype = _LL
    lib.bd_ball_query_grid_workspace_bytes.argtypes = [_I, _I]
    lib.bd_attention_tc_workspace_bytes.restype = _LL
    lib.bd_attention_tc_workspace_bytes.argtypes = [_I, _I, _I, _I, _I]
    lib.bd_grid_order.restype = ctypes.c_void_p
    lib.bd_grid_order.argtypes = [_P, _I, _I]
    lib.bd_attention_tc_select.restype = _I
    lib.bd_attention_tc_select.argtypes = [_I]
    _lib = lib
    return lib


def last_error():
    return load().bd_last_error().decode()


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_profiler = None


class Profiler:
    """Per-launch CUDA-event timing of every C-ABI call made while active (eager launches only).
    Events are recorded on the stream the kernel is launched on.  Keys are the entry point plus
    its integer size arguments, e.g. `bd_fps(6, 8, 50000, 2048)`."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _profiler
        _profiler = self
        return self

    def __exit__(self, *exc):
        global _profiler
        _profiler = None

    def table(self):
        """[{name, launches, total_ms, mean_ms, share}] sorted by total time (call after a sync)."""
        agg = {}
        for key, args, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            a = agg.setdefault(key, [0, 0.0, args])
            a[0] += 1
            a[1] += ms
        total = sum(a[1] for a in agg.values()) or 1.0
        rows = [{"name": k, "launches": n, "total_ms": t, "mean_ms": t / n, "share": t / total, "args": list(args)}
                for k, (n, t, args) in agg.items()]
        return sorted(rows, key=lambda r: -r["total_ms"])


def call(name, *args):
    """Invoke `name(*args, current_stream)`; raise RuntimeError with the library's message on failure."""
    global launch_count
    lib = load()
    if _profiler is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream_ptr())
        e1.record()
        sizes = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and 0 <= a < (1 << 24))
        _profiler.records.append((f"{name}{sizes}", args, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream_ptr())
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {lib.bd_last_error().decode()}")


def check_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("CPU not supported: butd_detr_b200 runs on CUDA tensors only")
