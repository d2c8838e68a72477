"""Drop-in replacement for the reference's native plugin module `pointnet2._ext`.

Same nine functions, argument order, dtypes, output allocation (callee allocates, zero
filled) and error behaviour (RuntimeError on non-contiguous / wrong dtype / CPU tensors) as
the pybind11 module built from `/root/reference/pointnet2/_ext_src`
(`src/bindings.cpp:11-24`; checks in `include/utils.h:10-30`), but every op runs on the
sm_100a kernels of libbutd_b200.so through the C-ABI.  Installing it for the unmodified
reference Python is one line:

    import sys, butd_detr_b200.pointnet2_ext as e; sys.modules["pointnet2._ext"] = e
"""
import torch

from . import _lib


GRID_MIN_POINTS = 8192  # clouds at least this large use the cell list (ball query, bucketed FPS)
FPS_GRID_MIN_BATCH = 38  # as engine.FPS_GRID_MIN_BATCH (measured cross-over with the cluster kernel)


def use_grid_ball_query(n, radius, nsample, m):
    """One rule for engine.py and this module (DESIGN.md §4, configs[4] sweep): the cell list wins
    when a ball holds few points compared with the scan the ordered brute force needs to find
    `nsample` of them, i.e. for large clouds and small balls."""
    from .engine import grid_ball_query_rule
    return grid_ball_query_rule(n, radius, nsample, m)


def _check(t, name, dtype):
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {'an int' if dtype == torch.int32 else 'a float'} tensor")
    if not t.is_cuda:
        raise RuntimeError("CPU not supported")


def furthest_point_sampling(points, nsamples):
    """(B,N,3) f32 -> (B,nsamples) i32   [sampling.cpp:70-91]"""
    _check(points, "points", torch.float32)
    B, N, _ = points.shape
    out = torch.zeros(B, nsamples, dtype=torch.int32, device=points.device)
    lib = _lib.load()
    with torch.cuda.device(points.device):
        if GRID_MIN_POINTS <= N <= lib.bd_fps_grid_capacity() and B >= FPS_GRID_MIN_BATCH:
            ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device=points.device)
            scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8, device=points.device)
            _lib.call("bd_grid_build", points.data_ptr(), 3, B, N, -1.0, ws.data_ptr())  # cell size picked from the extent
            _lib.call("bd_fps_grid", points.data_ptr(), 3, B, N, int(nsamples), ws.data_ptr(), scratch.data_ptr(),
                      out.data_ptr())
            return out
        tmp = None
        if N > lib.bd_fps_resident_capacity():
            tmp = torch.empty(B, N, dtype=torch.float32, device=points.device)
        _lib.call("bd_fps", points.data_ptr(), 3, B, N, int(nsamples), _lib.ptr(tmp), out.data_ptr())
    return out


def gather_points(points, idx):
    """(B,C,N) f32, (B,m) i32 -> (B,C,m)   [sampling.cpp:20-43]"""
    _check(points, "points", torch.float32)
    _check(idx, "idx", torch.int32)
    B, C, N = points.shape
    m = idx.shape[1]
    out = torch.zeros(B, C, m, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.call("bd_gather_points", points.data_ptr(), idx.data_ptr(), B, C, N, m, out.data_ptr())
    return out


def gather_points_grad(grad_out, idx, n):
    """(B,C,m) f32, (B,m) i32, N -> (B,C,N)   [sampling.cpp:45-69]"""
    _check(grad_out, "grad_out", torch.float32)
    _check(idx, "idx", torch.int32)
    B, C, m = grad_out.shape
    out = torch.empty(B, C, n, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.call("bd_gather_points_grad", grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), m, out.data_ptr())
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,m,3), (B,n,3), r, ns -> (B,m,ns) i32   [ball_query.cpp:13-37]"""
    _check(new_xyz, "new_xyz", torch.float32)
    _check(xyz, "xyz", torch.float32)
    B, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.zeros(B, m, nsample, dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        # cell-list search (identical output): pays off when the ball holds about nsample points;
        # for small nsample the ordered brute-force scan exits early and wins (measured)
        if use_grid_ball_query(n, radius, nsample, m):
            ws = torch.empty(_lib.load().bd_ball_query_grid_workspace_bytes(B, n), dtype=torch.uint8, device=xyz.device)
            _lib.call("bd_ball_query_grid", new_xyz.data_ptr(), xyz.data_ptr(), 3, B, n, m, float(radius),
                      int(nsample), out.data_ptr(), ws.data_ptr())
        else:
            _lib.call("bd_ball_query", new_xyz.data_ptr(), xyz.data_ptr(), 3, B, n, m, float(radius), int(nsample),
                      out.data_ptr())
    return out


def group_points(points, idx):
    """(B,C,n) f32, (B,m,ns) i32 -> (B,C,m,ns)   [group_points.cpp:17-40]"""
    _check(points, "points", torch.float32)
    _check(idx, "idx", torch.int32)
    B, C, n = points.shape
    _, m, ns = idx.shape
    out = torch.zeros(B, C, m, ns, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.call("bd_group_points", points.data_ptr(), idx.data_ptr(), B, C, n, m, ns, out.data_ptr())
    return out


def group_points_grad(grad_out, idx, n):
    """(B,C,m,ns) f32, (B,m,ns) i32, n -> (B,C,n)   [group_points.cpp:42-65]"""
    _check(grad_out, "grad_out", torch.float32)
    _check(idx, "idx", torch.int32)
    B, C, m, ns = grad_out.shape
    out = torch.empty(B, C, n, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.call("bd_group_points_grad", grad_out.data_ptr(), idx.data_ptr(), B, C, int(n), m, ns, out.data_ptr())
    return out


def three_nn(unknown, known):
    """(B,n,3), (B,m,3) -> [dist2 (B,n,3) f32 (squared), idx (B,n,3) i32]   [interpolate.cpp:19-45]"""
    _check(unknown, "unknowns", torch.float32)
    _check(known, "knows", torch.float32)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros(B, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.zeros(B, n, 3, dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        _lib.call("bd_three_nn", unknown.data_ptr(), known.data_ptr(), B, n, m, dist2.data_ptr(), idx.data_ptr())
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B,C,m) f32, (B,n,3) i32, (B,n,3) f32 -> (B,C,n)   [interpolate.cpp:47-75]"""
    _check(points, "points", torch.float32)
    _check(idx, "idx", torch.int32)
    _check(weight, "weight", torch.float32)
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(B, C, n, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.call("bd_three_interpolate", points.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, C, m, n,
                  out.data_ptr())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B,C,n) f32, (B,n,3) i32, (B,n,3) f32, m -> (B,C,m)   [interpolate.cpp:76-104]"""
    _check(grad_out, "grad_out", torch.float32)
    _check(idx, "idx", torch.int32)
    _check(weight, "weight", torch.float32)
    B, C, n = grad_out.shape
    out = torch.empty(B, C, m, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.call("bd_three_interpolate_grad", grad_out.data_ptr(), idx.data_ptr(), weight.data_ptr(), B, C, n,
                  int(m), out.data_ptr())
    return out
