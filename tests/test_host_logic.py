"""CPU: host-side logic of the tensor-core path — weight packing into the 128-byte-swizzle operand
layout (csrc/tc_common.cuh), tiling policy, the bench's per-kernel work figures."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _unpack(P, N, K, BN, n_sub):
    """Invert pack_weight_tc: Wp[ng][kc][part][sub][row][chunk ^ (row % 8)][8] -> (parts, N, K)."""
    ng, nch, parts = P.shape[0], P.shape[1], P.shape[2]
    out = torch.zeros(parts, ng * n_sub * BN, nch * 64, dtype=P.dtype)
    for g in range(ng):
        for c in range(nch):
            for s in range(n_sub):
                blk = P[g, c, :, s]                                # (parts, BN, 8 slots, 8)
                for r in range(BN):
                    for slot in range(8):
                        chunk = slot ^ (r % 8)
                        out[:, (g * n_sub + s) * BN + r, c * 64 + chunk * 8:c * 64 + chunk * 8 + 8] = blk[:, r, slot]
    return out[:, :N, :K]


@pytest.mark.parametrize("N,K,kw", [(288, 288, {}), (288, 136, {"wide": True}), (256, 64, {"full_rows": True}), (3, 288, {})])
def test_pack_weight_tc_layout_and_formats(N, K, kw):
    from butd_detr_b200.engine import pack_weight_tc
    g = torch.Generator().manual_seed(N + K)
    W = torch.randn(N, K, generator=g)
    # split 3: bf16 hi + bf16 lo reconstruct W to ~2^-16; the 16-byte chunk j of row r holds k-chunk j ^ (r % 8)
    P, (BN, KC, nch, n_sub) = pack_weight_tc(W, 3, **kw)
    assert KC == 64 and nch == -(-K // 64) and BN % 16 == 0 and P.dtype == torch.bfloat16
    parts = _unpack(P, N, K, BN, n_sub).float()
    assert torch.equal(parts[0], W.bfloat16().float())
    assert (parts[0] + parts[1] - W).abs().max() <= 2.0 ** -15 * W.abs().max()
    # split 1: the same layout holding fp16 bit patterns (the kernels' idesc says F16)
    P1, _ = pack_weight_tc(W, 1, **kw)
    assert P1.shape[2] == 1
    assert torch.equal(_unpack(P1, N, K, BN, n_sub)[0].view(torch.float16).float(), W.half().float())


def test_lin_tiling_policy():
    from butd_detr_b200.engine import lin_tiling, tc_tiling
    assert lin_tiling(8192, 288) == (False, None)      # 64 row tiles x 2 column tiles fit one wave: narrow
    assert lin_tiling(8192, 576)[0] is True            # 256 CTAs would not: two accumulators per CTA
    assert lin_tiling(65536, 64) == (False, None)      # a single 64-wide accumulator cannot be widened
    BN, KC, nch, n_sub = tc_tiling(288, 288, 3, wide=True)
    assert (BN, n_sub, nch) == (144, 2, 5)
    assert tc_tiling(288, 288, 3, full_rows=True)[3] == 2 and tc_tiling(256, 128, 1, full_rows=True)[0] == 128


def test_bench_work_figures():
    import bench
    # fused SA level: 2 * rows * (K1 N1 + N1 N2 + N2 N3) with K1 = C + 3 (SA1 = 3.32 GF / scene)
    a = [0, 0, 6, 3, 0, 6, 0, 1, 50000, 2048, 64, 0.2, 0, 0, 64, 0, 0, 64, 0, 0, 128]
    bound, flops = bench.algorithmic_work("bd_sa_mlp_tc", a)
    assert bound == "tensor" and abs(flops - 2 * 2048 * 64 * (6 * 64 + 64 * 64 + 64 * 128)) < 1
    assert abs(flops / 1e9 - 3.32) < 0.02
    assert bench.algorithmic_work("bd_fps", [0, 6, 4, 50000, 2048]) == ("hbm", 4 * (12 * 50000 + 4 * 2048))
    assert bench.algorithmic_work("bd_fps_grid", [0, 6, 4, 50000, 2048]) == ("hbm", 4 * (12 * 50000 + 4 * 2048))
    assert bench.PARITY_GATE == {"fp32": 1e-3, "bf16x3": 1e-3, "fp16": 1e-2}  # north_star's tolerances
    assert not hasattr(bench, "PARITY_MEASURED")  # the error is measured in every run (bench.measure_parity)
    att = bench.attention_summary(4000.0, [])
    assert abs(att["attention_gemm_flop_roofline_frac"] - 4000 * 17.9e9 / (bench.tensor_peak() * 1e12)) < 1e-12


def test_round_2b_modules_refuse_cpu_tensors():
    """The text-side engine and the device matcher have no CPU path: they fail loudly on CPU inputs / devices."""
    import pytest
    import torch
    from butd_detr_b200 import text_encoder
    from butd_detr_b200.matcher import HungarianMatcher
    with pytest.raises(RuntimeError, match="CPU not supported"):
        text_encoder.RobertaEngine({}, None, "cpu")
    out = {"pred_logits": torch.zeros(1, 4, 8), "pred_boxes": torch.zeros(1, 4, 6)}
    tg = [{"boxes": torch.zeros(2, 6), "positive_map": torch.zeros(2, 8), "labels": torch.zeros(2, dtype=torch.int64)}]
    with pytest.raises(RuntimeError, match="CPU not supported"):
        HungarianMatcher(1, 5, 2, True)(out, tg)


def test_grid_ball_query_rule_values():
    """One rule for the engine and the pointnet2._ext drop-in: the cell list for SA1 (50k points, r = 0.2) and SA2
    (2048 points, r = 0.4, nsample = 32 — exactly on the nsample >= 200 r^2 boundary), the ordered scan elsewhere."""
    from butd_detr_b200.engine import grid_ball_query_rule
    assert grid_ball_query_rule(50000, 0.2, 64, 2048) and grid_ball_query_rule(50000, 0.2, 16, 2048)
    assert grid_ball_query_rule(2048, 0.4, 32, 1024)
    assert not grid_ball_query_rule(1024, 0.8, 16, 512) and not grid_ball_query_rule(512, 1.2, 16, 256)
    assert not grid_ball_query_rule(50000, 0.8, 16, 2048) and not grid_ball_query_rule(50000, 0.2, 100, 2048)
