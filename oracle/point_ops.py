"""TEST INFRASTRUCTURE — ctypes front-end of `oracle/point_ops_ref.c` (the CPU restatement of
the reference's CUDA-only point ops).  Exposes the same nine functions, argument order and
dtypes as the reference's `pointnet2._ext` pybind module
(`/root/reference/pointnet2/_ext_src/src/bindings.cpp:11-24`) on CPU torch tensors, so the
reference's unmodified Python (`pointnet2_utils.py`, `pointnet2_modules.py`) can run on it
when generating golden vectors, and so tests can compare our CUDA kernels against it.

Never imported by the product package `butd_detr_b200`.
"""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "point_ops_ref.c")
LIB = os.path.join(HERE, "liboracle_pointops.so")

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", "-shared", "-fPIC",
               SRC, "-o", LIB, "-lm"]
        subprocess.check_call(cmd)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int))


def num_threads():
    return lib().orc_num_threads()


def opt_n_threads(n):
    return lib().orc_opt_n_threads(int(n))


def furthest_point_sampling(points, nsamples):
    B, N, _ = points.shape
    out = torch.zeros(B, nsamples, dtype=torch.int32)
    lib().orc_fps(_f(points), B, N, int(nsamples), _i(out))
    return out


def gather_points(points, idx):
    B, C, N = points.shape
    m = idx.shape[1]
    out = torch.zeros(B, C, m, dtype=torch.float32)
    lib().orc_gather(_f(points), _i(idx), B, C, N, m, _f(out))
    return out


def gather_points_grad(grad_out, idx, n):
    B, C, m = grad_out.shape
    out = torch.zeros(B, C, n, dtype=torch.float32)
    lib().orc_gather_grad(_f(grad_out), _i(idx), B, C, int(n), m, _f(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    B, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.zeros(B, m, nsample, dtype=torch.int32)
    lib().orc_ball_query(_f(new_xyz), _f(xyz), B, n, m, ctypes.c_float(radius), int(nsample), _i(out))
    return out


def group_points(points, idx):
    B, C, n = points.shape
    _, m, ns = idx.shape
    out = torch.zeros(B, C, m, ns, dtype=torch.float32)
    lib().orc_group(_f(points), _i(idx), B, C, n, m, ns, _f(out))
    return out


def group_points_grad(grad_out, idx, n):
    B, C, m, ns = grad_out.shape
    out = torch.zeros(B, C, n, dtype=torch.float32)
    lib().orc_group_grad(_f(grad_out), _i(idx), B, C, int(n), m, ns, _f(out))
    return out


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros(B, n, 3, dtype=torch.float32)
    idx = torch.zeros(B, n, 3, dtype=torch.int32)
    lib().orc_three_nn(_f(unknown), _f(known), B, n, m, _f(dist2), _i(idx))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(B, C, n, dtype=torch.float32)
    lib().orc_three_interpolate(_f(points), _i(idx), _f(weight), B, C, m, n, _f(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, n = grad_out.shape
    out = torch.zeros(B, C, m, dtype=torch.float32)
    lib().orc_three_interpolate_grad(_f(grad_out), _i(idx), _f(weight), B, C, n, int(m), _f(out))
    return out
