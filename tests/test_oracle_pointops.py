"""CPU: pins the C restatement of the point ops (oracle/point_ops_ref.c) against
  (1) tests/golden/pointops_refcuda.npz — outputs of the reference's OWN CUDA kernels, built
      unmodified and run on a B200 (tests/golden/make_pointops_golden.py), bit-for-bit;
  (2) brute-force numpy restatements of the semantics table in SURVEY.md Appendix A;
  (3) structural invariants the reference relies on."""
import os

import numpy as np
import pytest
import torch

from pointops_cases import BALL_CASES, FPS_CASES, ball_inputs, case_seed, cloud

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pointops_refcuda.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case", FPS_CASES, ids=[c[0] for c in FPS_CASES])
def test_fps_matches_reference_cuda(case, oracle_lib, gold):
    name, kind, B, N, m = case
    got = oracle_lib.furthest_point_sampling(cloud(case_seed(name), N, kind, B), m).numpy()
    assert np.array_equal(got, gold["fps/" + name])


@pytest.mark.parametrize("case", BALL_CASES, ids=[c[0] for c in BALL_CASES])
def test_ball_query_matches_reference_cuda(case, oracle_lib, gold):
    xyz, new_xyz, r, ns = ball_inputs(case)
    assert np.array_equal(oracle_lib.ball_query(new_xyz, xyz, r, ns).numpy(), gold["ball/" + case[0]])


@pytest.mark.parametrize("n,m,kind", [(512, 256, "room"), (1024, 512, "room"), (300, 2, "uniform"),
                                        (257, 1500, "lattice")])
def test_three_nn_interpolate_match_reference_cuda(n, m, kind, oracle_lib, gold):
    unknown, known = cloud(31, n, kind, 2), cloud(32, m, kind, 2)
    d, i = oracle_lib.three_nn(unknown, known)
    assert np.array_equal(i.numpy(), gold[f"nn_idx/{n}_{m}"])
    assert np.array_equal(d.numpy(), gold[f"nn_dist2/{n}_{m}"])  # inf == inf when m < 3
    if m >= 3:
        g = torch.Generator().manual_seed(2)
        feats = torch.randn(2, 33, m, generator=g)
        w = torch.rand(2, n, 3, generator=g)
        w = (w / w.sum(-1, keepdim=True)).contiguous()
        assert np.array_equal(oracle_lib.three_interpolate(feats, i, w).numpy(), gold[f"interp/{n}_{m}"])


def _d2(a, b):
    """fp32 squared distance with the reference's contraction, via float64 emulation of FMA."""
    dx, dy, dz = [(a[..., i] - b[..., i]).astype(np.float32) for i in range(3)]
    t = (dy * dy).astype(np.float32)
    t = (dx.astype(np.float64) * dx.astype(np.float64) + t.astype(np.float64)).astype(np.float32)
    return (dz.astype(np.float64) * dz.astype(np.float64) + t.astype(np.float64)).astype(np.float32)


def test_ball_query_semantics_bruteforce(oracle_lib):
    xyz, new_xyz, r, ns = ball_inputs(("lattice", "lattice", 2, 3000, 333, 0.5, 16))
    got = oracle_lib.ball_query(new_xyz, xyz, r, ns).numpy()
    r2 = np.float32(r) * np.float32(r)
    for b in range(2):
        for j in range(0, 333, 7):
            d2 = _d2(new_xyz[b, j].numpy()[None], xyz[b].numpy())
            hits = np.nonzero(d2 < r2)[0][:ns]
            want = np.zeros(ns, np.int32)
            if len(hits):
                want[:] = hits[0]
                want[:len(hits)] = hits
            assert np.array_equal(got[b, j], want)


def test_fps_semantics_start_skip_and_ties(oracle_lib):
    # starts at 0, never selects near-origin points (|p|^2 <= 1e-3), handles all-skipped input
    pts = torch.tensor([[[1.0, 0, 0], [0.01, 0.0, 0.0], [0, 2.0, 0], [0.0, 0.0, 0.02], [-3.0, 0, 0]]])
    out = oracle_lib.furthest_point_sampling(pts, 4)[0].tolist()
    assert out[0] == 0 and 1 not in out[1:] and 3 not in out[1:]
    assert out[:3] == [0, 4, 2]
    zeros = torch.zeros(1, 40, 3)
    assert oracle_lib.furthest_point_sampling(zeros, 5)[0].tolist() == [0] * 5
    # tie rule: equal maxima -> smallest bit-reversed (k mod 512), e.g. k=256 beats k=128
    n = 600
    pts = torch.zeros(1, n, 3)
    pts[0, :, 0] = 1.0                       # everything at the same place as point 0 ...
    pts[0, 128] = torch.tensor([5.0, 0, 0])  # ... except two points at the same distance
    pts[0, 256] = torch.tensor([-3.0, 0, 0])
    assert oracle_lib.furthest_point_sampling(pts, 2)[0].tolist() == [0, 256]


def test_fps_of_fps_ordered_prefix_is_identity(oracle_lib):
    xyz = cloud(9, 20000, "uniform", 1)
    i1 = oracle_lib.furthest_point_sampling(xyz, 2048)
    lvl = torch.gather(xyz, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    assert torch.equal(oracle_lib.furthest_point_sampling(lvl, 1024), torch.arange(1024, dtype=torch.int32)[None])


def test_gather_group_grads_are_adjoint(oracle_lib):
    g = torch.Generator().manual_seed(0)
    pts = torch.randn(2, 5, 64, generator=g)
    idx = torch.randint(0, 64, (2, 20, 4), generator=g, dtype=torch.int32)
    out = oracle_lib.group_points(pts, idx)
    assert torch.equal(out, torch.gather(pts.unsqueeze(2).expand(-1, -1, 20, -1), 3,
                                         idx.long().unsqueeze(1).expand(-1, 5, -1, -1)))
    go = torch.randn(2, 5, 20, 4, generator=g)
    lhs = (out * go).sum()
    rhs = (pts * oracle_lib.group_points_grad(go, idx, 64)).sum()
    assert abs(float(lhs - rhs)) < 1e-3
    i1 = torch.randint(0, 64, (2, 9), generator=g, dtype=torch.int32)
    go1 = torch.randn(2, 5, 9, generator=g)
    assert abs(float((oracle_lib.gather_points(pts, i1) * go1).sum()
                     - (pts * oracle_lib.gather_points_grad(go1, i1, 64)).sum())) < 1e-3


def test_opt_n_threads_matches_cuda_utils(oracle_lib):
    for n, want in ((1, 1), (2, 2), (3, 2), (37, 32), (511, 256), (512, 512), (513, 512), (50000, 512), (4096, 512)):
        assert oracle_lib.opt_n_threads(n) == want
