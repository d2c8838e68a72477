"""GPU parity of the nine pointnet2._ext ops (through the C-ABI via the drop-in module) against
the CPU oracle and, when oracle/_ref holds it, the reference's own CUDA extension.
Indices must be bit-exact; distances / interpolated values exact (same FMA contraction)."""
import os
import sys

import numpy as np
import pytest
import torch

from pointops_cases import BALL_CASES, FPS_CASES, ball_inputs, case_seed, cloud

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext(cuda_lib):
    from butd_detr_b200 import pointnet2_ext
    return pointnet2_ext


@pytest.fixture(scope="module")
def ref_ext():
    """The reference's unmodified CUDA extension (built in the container into oracle/_ref)."""
    from oracle import build_ref_ext
    try:
        return build_ref_ext.load_ref_ext()
    except Exception as e:  # pragma: no cover
        print("reference ext unavailable:", e)
        return None


@pytest.mark.parametrize("case", FPS_CASES, ids=[c[0] for c in FPS_CASES])
def test_fps_bit_exact(case, ext, oracle_lib, ref_ext):
    name, kind, B, N, m = case
    xyz = cloud(case_seed(name), N, kind, B)
    want = oracle_lib.furthest_point_sampling(xyz, m)
    got = ext.furthest_point_sampling(xyz.cuda(), m).cpu()
    assert torch.equal(got, want), f"{name}: first mismatch at {(got != want).nonzero()[:3].tolist()}"
    if ref_ext is not None:
        ref = ref_ext.furthest_point_sampling(xyz.cuda(), m).cpu()
        assert torch.equal(got, ref), f"{name}: differs from the reference CUDA kernel"


@pytest.mark.parametrize("cluster", [8, 16])
def test_fps_cluster_sizes_agree(cluster, ext, oracle_lib, cuda_lib):
    xyz = cloud(3, 50000, "room", 2)
    want = oracle_lib.furthest_point_sampling(xyz, 512)
    cuda_lib.load().bd_fps_set_cluster(cluster)
    try:
        got = ext.furthest_point_sampling(xyz.cuda(), 512).cpu()
    finally:
        cuda_lib.load().bd_fps_set_cluster(-1)
    assert torch.equal(got, want)


@pytest.mark.parametrize("kind,B,N,m", [("room", 2, 50000, 2048), ("room", 10, 50000, 1024), ("lattice", 9, 40000, 700),
                                        ("dup", 3, 50000, 600), ("uniform", 12, 30000, 512)])
def test_fps_ordered_equals_fps(cuda_lib, oracle_lib, kind, B, N, m):
    """bd_fps_ordered (cell-list order + per-thread pruning) returns bd_fps's indices bit for bit
    (8-CTA clusters for B <= 8, 4-CTA clusters beyond; ties, duplicates, partially empty CTAs)."""
    lib = cuda_lib.load()
    xyz = cloud(11 + B, N, kind, B).cuda()
    want = torch.zeros(B, m, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_fps", xyz.data_ptr(), 3, B, N, m, None, want.data_ptr())
    ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
    cuda_lib.call("bd_grid_build", xyz.data_ptr(), 3, B, N, 0.2, ws.data_ptr())
    got = torch.full((B, m), -1, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_fps_ordered", xyz.data_ptr(), 3, B, N, m, lib.bd_grid_order(ws.data_ptr(), B, N), None,
                  got.data_ptr())
    assert torch.equal(got, want), f"first mismatch at {(got != want).nonzero()[:3].tolist()}"
    if B <= 2:
        assert torch.equal(got.cpu(), oracle_lib.furthest_point_sampling(xyz.cpu(), m))


def test_fps_strided_input_matches_contiguous(cuda_lib, oracle_lib):
    """FPS straight from the (B,N,6) point cloud (ld = 6) == FPS on the xyz copy."""
    from butd_detr_b200 import synth
    pc = torch.from_numpy(synth.synth_scene(5, 6000, 8)["point_clouds"])[None].cuda()
    out = torch.zeros(1, 777, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_fps", pc.data_ptr(), 6, 1, 6000, 777, None, out.data_ptr())
    want = oracle_lib.furthest_point_sampling(pc[..., :3].contiguous().cpu(), 777)
    assert torch.equal(out.cpu(), want)


def test_fps_on_fps_ordered_points_is_identity(ext):
    """models/backbone_module.py:122 relies on this."""
    xyz = cloud(9, 20000, "uniform", 1).cuda()
    i1 = ext.furthest_point_sampling(xyz, 2048)
    lvl = torch.gather(xyz, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    i2 = ext.furthest_point_sampling(lvl, 1024)
    assert torch.equal(i2.cpu(), torch.arange(1024, dtype=torch.int32)[None])


@pytest.mark.parametrize("case", BALL_CASES, ids=[c[0] for c in BALL_CASES])
def test_ball_query_bit_exact(case, ext, oracle_lib, ref_ext):
    xyz, new_xyz, r, ns = ball_inputs(case)
    want = oracle_lib.ball_query(new_xyz, xyz, r, ns)
    got = ext.ball_query(new_xyz.cuda(), xyz.cuda(), r, ns).cpu()
    assert torch.equal(got, want)
    if ref_ext is not None:
        assert torch.equal(got, ref_ext.ball_query(new_xyz.cuda(), xyz.cuda(), r, ns).cpu())


def test_ball_query_full_scene_50k(ext, oracle_lib):
    xyz = cloud(21, 50000, "room", 1)
    inds = oracle_lib.furthest_point_sampling(xyz, 2048)
    new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for r, ns in ((0.2, 64), (0.4, 32), (0.8, 16)):
        want = oracle_lib.ball_query(new_xyz, xyz, r, ns)
        got = ext.ball_query(new_xyz.cuda(), xyz.cuda(), r, ns).cpu()
        assert torch.equal(got, want)


def test_gather_group_roundtrip(ext, oracle_lib):
    g = torch.Generator().manual_seed(1)
    pts = torch.randn(2, 19, 700, generator=g)
    idx = torch.randint(0, 700, (2, 300), generator=g, dtype=torch.int32)
    assert torch.equal(ext.gather_points(pts.cuda(), idx.cuda()).cpu(), oracle_lib.gather_points(pts, idx))
    gidx = torch.randint(0, 700, (2, 50, 16), generator=g, dtype=torch.int32)
    assert torch.equal(ext.group_points(pts.cuda(), gidx.cuda()).cpu(), oracle_lib.group_points(pts, gidx))
    # backward ops: scatter-add (atomic order differs -> tolerance)
    go = torch.randn(2, 19, 300, generator=g)
    torch.testing.assert_close(ext.gather_points_grad(go.cuda(), idx.cuda(), 700).cpu(),
                               oracle_lib.gather_points_grad(go, idx, 700), rtol=1e-5, atol=1e-5)
    gg = torch.randn(2, 19, 50, 16, generator=g)
    torch.testing.assert_close(ext.group_points_grad(gg.cuda(), gidx.cuda(), 700).cpu(),
                               oracle_lib.group_points_grad(gg, gidx, 700), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,m,kind", [(512, 256, "room"), (1024, 512, "room"), (300, 2, "uniform"),
                                        (257, 1500, "lattice")])
def test_three_nn_and_interpolate(n, m, kind, ext, oracle_lib, ref_ext):
    unknown, known = cloud(31, n, kind, 2), cloud(32, m, kind, 2)
    d_want, i_want = oracle_lib.three_nn(unknown, known)
    d_got, i_got = ext.three_nn(unknown.cuda(), known.cuda())
    assert torch.equal(i_got.cpu(), i_want)
    assert torch.equal(d_got.cpu(), d_want)  # same FMA contraction -> bitwise equal (inf when m < 3)
    if ref_ext is not None:
        d_ref, i_ref = ref_ext.three_nn(unknown.cuda(), known.cuda())
        assert torch.equal(i_got, i_ref) and torch.equal(d_got, d_ref)
    if m < 3:
        return
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(2, 33, m, generator=g)
    w = torch.rand(2, n, 3, generator=g)
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    want = oracle_lib.three_interpolate(feats, i_want, w)
    got = ext.three_interpolate(feats.cuda(), i_want.cuda(), w.cuda()).cpu()
    assert torch.equal(got, want)
    go = torch.randn(2, 33, n, generator=g)
    torch.testing.assert_close(ext.three_interpolate_grad(go.cuda(), i_want.cuda(), w.cuda(), m).cpu(),
                               oracle_lib.three_interpolate_grad(go, i_want, w, m), rtol=1e-5, atol=1e-5)


def test_error_behaviour_matches_reference(ext):
    """TORCH_CHECK-style RuntimeErrors (utils.h:10-30), 'CPU not supported' (ball_query.cpp:32-34)."""
    x = torch.randn(1, 64, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        ext.furthest_point_sampling(x, 8)
    with pytest.raises(RuntimeError, match="contiguous"):
        ext.furthest_point_sampling(torch.randn(1, 3, 64).cuda().transpose(1, 2), 8)
    with pytest.raises(RuntimeError, match="int tensor"):
        ext.gather_points(torch.randn(1, 3, 64).cuda(), torch.zeros(1, 8, dtype=torch.int64).cuda())


def test_reference_python_runs_unmodified_on_the_drop_in(ext, oracle_lib):
    """The reference's QueryAndGroup semantics through our _ext (restated here because
    /root/reference is not on the GPU box): grouped = cat((xyz[idx]-centre)/r, feats[idx])."""
    xyz = cloud(41, 3000, "room", 2)
    feats = torch.randn(2, 5, 3000, generator=torch.Generator().manual_seed(3))
    inds = oracle_lib.furthest_point_sampling(xyz, 128)
    new_xyz = oracle_lib.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    idx = ext.ball_query(new_xyz.cuda(), xyz.cuda(), 0.3, 16)
    g_xyz = ext.group_points(xyz.transpose(1, 2).contiguous().cuda(), idx)
    g_xyz = (g_xyz - new_xyz.cuda().transpose(1, 2).unsqueeze(-1)) / 0.3
    g_f = ext.group_points(feats.cuda(), idx)
    got = torch.cat([g_xyz, g_f], 1).cpu()
    from oracle import model_ref
    want = model_ref.query_and_group(xyz, new_xyz, feats, 0.3, 16)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)


GRID_CASES = BALL_CASES + [("dense_overflow", "dup", 1, 6000, 64, 2.0, 32), ("wide", "uniform", 2, 9000, 300, 1.7, 16)]


@pytest.mark.parametrize("case", GRID_CASES, ids=[c[0] for c in GRID_CASES])
def test_ball_query_grid_equals_bruteforce(case, cuda_lib, oracle_lib):
    """Cell-list ball query == ordered brute force, bit for bit (incl. > 768-hit balls, empty balls,
    centres outside the cloud's bounding box, d2 == r2 lattice ties)."""
    xyz, new_xyz, r, ns = ball_inputs(case) if case in BALL_CASES else (
        cloud(case_seed(case[0]), case[3], case[1], case[2]), cloud(77, case[4], "uniform", case[2]) * 1.2, case[5], case[6])
    B, n, m = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
    want = oracle_lib.ball_query(new_xyz.contiguous(), xyz, r, ns)
    xd, cd = xyz.cuda(), new_xyz.contiguous().cuda()
    out = torch.full((B, m, ns), -1, dtype=torch.int32, device="cuda")
    ws = torch.empty(cuda_lib.load().bd_ball_query_grid_workspace_bytes(B, n), dtype=torch.uint8, device="cuda")
    cuda_lib.call("bd_ball_query_grid", cd.data_ptr(), xd.data_ptr(), 3, B, n, m, float(r), ns, out.data_ptr(),
                  ws.data_ptr())
    assert torch.equal(out.cpu(), want)


def test_ball_query_grid_full_scene_sweep(cuda_lib, oracle_lib):
    """BASELINE.json configs[4]: 50k points, nsample in {16,32,64} x radius in {0.2,0.4,0.8}."""
    xyz = cloud(21, 50000, "room", 1)
    inds = oracle_lib.furthest_point_sampling(xyz, 2048)
    new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    xd, cd = xyz.cuda(), new_xyz.cuda()
    ws = torch.empty(cuda_lib.load().bd_ball_query_grid_workspace_bytes(1, 50000), dtype=torch.uint8, device="cuda")
    for r in (0.2, 0.4, 0.8):
        for ns in (16, 32, 64):
            out = torch.full((1, 2048, ns), -1, dtype=torch.int32, device="cuda")
            cuda_lib.call("bd_ball_query_grid", cd.data_ptr(), xd.data_ptr(), 3, 1, 50000, 2048, r, ns, out.data_ptr(),
                          ws.data_ptr())
            assert torch.equal(out.cpu(), oracle_lib.ball_query(new_xyz, xyz, r, ns)), (r, ns)
