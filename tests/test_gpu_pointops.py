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


def _fps_grid(cuda_lib, xyz, m, radius, warps=16):
    """bd_grid_build + bd_fps_grid on a (B,N,3) CUDA tensor."""
    lib = cuda_lib.load()
    B, N, _ = xyz.shape
    ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8, device="cuda")
    out = torch.full((B, m), -1, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_grid_build", xyz.data_ptr(), 3, B, N, float(radius), ws.data_ptr())
    lib.bd_fps_grid_set_warps(warps)
    try:
        cuda_lib.call("bd_fps_grid", xyz.data_ptr(), 3, B, N, m, ws.data_ptr(), scratch.data_ptr(), out.data_ptr())
    finally:
        lib.bd_fps_grid_set_warps(0)
    return out


@pytest.mark.parametrize("kind,B,N,m,radius", [("room", 2, 50000, 2048, 0.2), ("room", 3, 50000, 700, -1.0),
                                               ("lattice", 9, 40000, 700, 0.2), ("dup", 3, 50000, 600, 0.2),
                                               ("uniform", 12, 30000, 512, -1.0), ("uniform", 2, 8192, 512, 0.05),
                                               ("room", 1, 65536, 300, 0.2), ("lattice", 2, 9001, 9001, 0.3),
                                               ("uniform", 2, 777, 300, -1.0)])
@pytest.mark.parametrize("warps", [16, 32, 8])
def test_fps_grid_bit_exact(cuda_lib, oracle_lib, kind, B, N, m, radius, warps):
    """bd_fps_grid (cell-list order, buckets of 32, bounding-box pruning) returns the oracle's indices bit
    for bit: ties on lattices, heavy duplication, more samples than distinct points, a partial last
    bucket, cells of any size (radius <= 0: picked from the extent)."""
    xyz = cloud(11 + B, N, kind, B)
    want = oracle_lib.furthest_point_sampling(xyz, m)
    got = _fps_grid(cuda_lib, xyz.cuda(), m, radius, warps).cpu()
    assert torch.equal(got, want), f"first mismatch at {(got != want).nonzero()[:3].tolist()}"


@pytest.mark.parametrize("B,m", [(9, 2048), (16, 2048), (37, 1024), (128, 2048)])
def test_fps_benched_batch_sizes_bit_exact(cuda_lib, oracle_lib, ref_ext, ext, B, m):
    """The FPS variants the benchmark's batch sizes select — bd_fps with 4-CTA clusters (B > 8) and
    bd_fps_grid — on B full 50k-point scenes, against the C oracle (every scene) and the reference's
    own CUDA kernel (sampling_gpu.cu:74-178)."""
    xyz = cloud(500 + B, 50000, "room", B)
    want = oracle_lib.furthest_point_sampling(xyz, m)
    xd = xyz.cuda()
    got = torch.full((B, m), -1, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_fps", xd.data_ptr(), 3, B, 50000, m, None, got.data_ptr())
    assert torch.equal(got.cpu(), want), "bd_fps (cluster kernel) differs from the oracle"
    got_grid = _fps_grid(cuda_lib, xd, m, 0.2).cpu()
    assert torch.equal(got_grid, want), f"bd_fps_grid differs from the oracle at {(got_grid != want).nonzero()[:3].tolist()}"
    assert torch.equal(ext.furthest_point_sampling(xd, m).cpu(), want), "drop-in furthest_point_sampling differs"
    if ref_ext is not None:
        assert torch.equal(ref_ext.furthest_point_sampling(xd, m).cpu(), want), "oracle differs from the reference kernel"


def test_fps_grid_two_scenes_per_sm_at_the_bench_batch(cuda_lib, oracle_lib):
    """296 scenes (bench.py's default batch): bd_fps_grid picks the 8-warp kernel, two CTAs per SM.  Same indices as
    the 16-warp kernel on every scene, and as the C oracle on the first / last scenes of both halves."""
    B, m = 296, 2048
    xyz = cloud(900, 50000, "room", B)
    xd = xyz.cuda()
    auto = _fps_grid(cuda_lib, xd, m, 0.2, warps=0)
    assert torch.equal(auto, _fps_grid(cuda_lib, xd, m, 0.2, warps=16))
    assert torch.equal(auto, _fps_grid(cuda_lib, xd, m, 0.2, warps=8))
    pick = [0, 147, 148, 295]
    assert torch.equal(auto[pick].cpu(), oracle_lib.furthest_point_sampling(xyz[pick].contiguous(), m))


def test_fps_grid_skipped_points_and_strided_rows(cuda_lib, oracle_lib):
    """Points with |p|^2 <= 1e-3 never win (sampling_gpu.cu:105-106); rows of 6 floats (ld = 6)."""
    from butd_detr_b200 import synth
    pc = torch.from_numpy(synth.synth_scene(5, 20000, 8)["point_clouds"])[None].clone()
    pc[0, 100:4000, :3] *= 0.004            # a dense blob inside the skip radius
    pc[0, 0, :3] = 0.0                      # the start point itself is skipped-class
    lib = cuda_lib.load()
    pcd = pc.cuda()
    ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(1, 20000), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(1, 20000), dtype=torch.uint8, device="cuda")
    out = torch.zeros(1, 1500, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_grid_build", pcd.data_ptr(), 6, 1, 20000, 0.2, ws.data_ptr())
    cuda_lib.call("bd_fps_grid", pcd.data_ptr(), 6, 1, 20000, 1500, ws.data_ptr(), scratch.data_ptr(), out.data_ptr())
    want = oracle_lib.furthest_point_sampling(pc[..., :3].contiguous(), 1500)
    assert torch.equal(out.cpu(), want)


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


@pytest.mark.parametrize("ns", [1, 7, 16, 33, 64, 65, 100])
@pytest.mark.parametrize("kind,n,m,r", [("room", 20000, 500, 0.25), ("dup", 9000, 200, 0.9), ("lattice", 12000, 300, 0.5)])
def test_ball_query_grid_selection_sizes(cuda_lib, oracle_lib, kind, n, m, r, ns):
    """Register selection of the cell-list query at every group size (1 .. 64: sorted file of 128
    candidates + index threshold, hundreds to thousands of hits per ball on the dense clouds;
    > 64 falls through to the ordered scan)."""
    xyz = cloud(61, n, kind, 2)
    new_xyz = (xyz[:, :m] + 0.01).contiguous()
    want = oracle_lib.ball_query(new_xyz, xyz, r, ns)
    out = torch.full((2, m, ns), -1, dtype=torch.int32, device="cuda")
    ws = torch.empty(cuda_lib.load().bd_ball_query_grid_workspace_bytes(2, n), dtype=torch.uint8, device="cuda")
    xd, cd = xyz.cuda(), new_xyz.cuda()
    cuda_lib.call("bd_ball_query_grid", cd.data_ptr(), xd.data_ptr(), 3, 2, n, m, float(r), ns, out.data_ptr(), ws.data_ptr())
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("build_radius", [-1.0, 0.07, 0.5])
def test_ball_query_on_a_grid_of_another_cell_size(cuda_lib, oracle_lib, build_radius):
    """bd_ball_query_grid_query on a cell list built for a different radius (cells smaller than the
    ball: more cells are visited; larger: fewer) — identical indices."""
    xyz = cloud(62, 15000, "room", 2)
    new_xyz = xyz[:, :400].contiguous()
    xd, cd = xyz.cuda(), new_xyz.cuda()
    ws = torch.empty(cuda_lib.load().bd_ball_query_grid_workspace_bytes(2, 15000), dtype=torch.uint8, device="cuda")
    cuda_lib.call("bd_grid_build", xd.data_ptr(), 3, 2, 15000, build_radius, ws.data_ptr())
    for r, ns in ((0.2, 64), (0.35, 16)):
        out = torch.full((2, 400, ns), -1, dtype=torch.int32, device="cuda")
        cuda_lib.call("bd_ball_query_grid_query", cd.data_ptr(), xd.data_ptr(), 3, 2, 15000, 400, r, ns, out.data_ptr(),
                      ws.data_ptr())
        assert torch.equal(out.cpu(), oracle_lib.ball_query(new_xyz, xyz, r, ns)), (build_radius, r, ns)


def test_grid_build_survives_non_finite_points(cuda_lib, oracle_lib):
    """A NaN / inf coordinate must not hang the cell-list build (it used to loop forever doubling the
    cell size); the finite points are still binned and queried."""
    xyz = cloud(63, 9000, "uniform", 1)
    xyz[0, 5] = float("inf")
    xyz[0, 6, 1] = float("nan")
    xd = xyz.cuda()
    cd = xyz[:, 100:164].contiguous().cuda()
    ws = torch.empty(cuda_lib.load().bd_ball_query_grid_workspace_bytes(1, 9000), dtype=torch.uint8, device="cuda")
    out = torch.full((1, 64, 16), -1, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_ball_query_grid", cd.data_ptr(), xd.data_ptr(), 3, 1, 9000, 64, 0.3, 16, out.data_ptr(), ws.data_ptr())
    torch.cuda.synchronize()
    # non-finite points are never inside a ball (NaN / inf distances), exactly as in the ordered scan
    assert torch.equal(out.cpu(), oracle_lib.ball_query(xyz[:, 100:164].contiguous(), xyz, 0.3, 16))
