"""GPU numerics of the tcgen05 kernels (split 1: fp16 operands, split 3: bf16 hi/lo operands; fp32
accumulate) vs PyTorch on the same rounded operands.  Tolerance: accumulation-order noise only (1e-4 relative)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("M,N,K,relu,add", [(128, 64, 16, False, False), (128, 144, 288, False, False),
                                             (256, 288, 288, True, False), (1024, 576, 288, False, True),
                                             (80, 288, 288, False, False), (300, 256, 288, True, False),
                                             (1000, 128, 136, True, False), (4099, 64, 8, True, False),
                                             (513, 160, 768, False, False), (512, 256, 512, True, False),
                                             (200, 864, 288, False, False), (77, 64, 288, False, False)])
@pytest.mark.parametrize("split", [1, 3])
def test_linear_tc(cuda_lib, M, N, K, relu, add, split):
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    A2 = torch.randn(M, K, device="cuda", generator=g) if add else None
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split)
    Y = torch.full((M, N), float("nan"), device="cuda")
    cuda_lib.call("bd_linear_tc", A.data_ptr(), K, cuda_lib.ptr(A2), K, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N,
                  M, N, K, KC, nch, BN, nsub, int(relu), split)
    torch.cuda.synchronize()
    a = A + A2 if add else A
    if split == 1:  # fp16 operands: compare with the same rounding applied
        want = F.linear(a.half().double(), W.half().double(), b.double())
        tol = 1e-4
    else:           # bf16x3: fp32-grade
        want = F.linear(a.double(), W.double(), b.double())
        tol = 1e-4
    want = (want.relu() if relu else want).float()
    torch.testing.assert_close(Y, want, rtol=tol, atol=tol)


def test_linear_tc_strided(cuda_lib):
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(9)
    big = torch.randn(300, 864, device="cuda", generator=g)
    W = torch.randn(160, 288, device="cuda", generator=g) / 17
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W)
    out = torch.zeros(300, 288, device="cuda")
    x = big[:, 288:576]
    cuda_lib.call("bd_linear_tc", x.data_ptr(), 864, None, 0, Wp.data_ptr(), None, out[:, 128:].data_ptr(), 288,
                  300, 160, 288, KC, nch, BN, nsub, 0, 1)
    want = (x.half().double() @ W.half().double().t()).float()
    torch.testing.assert_close(out[:, 128:], want, rtol=1e-4, atol=1e-4)
    assert float(out[:, :128].abs().max()) == 0.0


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("M,N,K,add", [(256, 288, 288, False), (1024, 288, 256, False), (80, 288, 288, True),
                                        (130, 160, 64, False), (2048, 288, 288, False)])
def test_linear_ln_tc(cuda_lib, M, N, K, add, split):
    """Fused out-projection / FFN-2 + residual + LayerNorm epilogue."""
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(M * 3 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    A2 = torch.randn(M, K, device="cuda", generator=g) if add else None
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g)
    gam, bet = torch.rand(N, device="cuda", generator=g) + 0.5, torch.randn(N, device="cuda", generator=g)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split, full_rows=True)
    Y = torch.full((M, N), float("nan"), device="cuda")
    cuda_lib.call("bd_linear_ln_tc", A.data_ptr(), K, cuda_lib.ptr(A2), K, Wp.data_ptr(), b.data_ptr(), R.data_ptr(), N,
                  gam.data_ptr(), bet.data_ptr(), 1e-5, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, split)
    a = A + A2 if add else A
    if split == 1:
        lin = F.linear(a.half().double(), W.half().double(), b.double())
    else:
        lin = F.linear(a.double(), W.double(), b.double())
    want = F.layer_norm(R.double() + lin, (N,), gam.double(), bet.double(), 1e-5).float()
    torch.testing.assert_close(Y, want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 1024, 1024, False), (2, 80, 1024, False), (2, 1024, 80, True),
                                             (3, 256, 132, True), (1, 32, 16, True), (2, 70, 65, True),
                                             (1, 256, 256, False), (2, 300, 1000, True), (1, 129, 257, False)])
def test_attention_tc(cuda_lib, B, Lq, Lk, masked, split, impl):
    H, hd = 8, 36
    cuda_lib.load().bd_attention_tc_select(impl)
    E = H * hd
    g = _g(Lq * 7 + Lk)
    qkv_q = torch.randn(B, Lq, 2 * E, device="cuda", generator=g)  # q lives in a wider fused buffer
    kv = torch.randn(B, Lk, 2 * E, device="cuda", generator=g)
    q, k, v = qkv_q[..., :E], kv[..., :E], kv[..., E:]
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=torch.Generator().manual_seed(Lk))
        mask = (torch.arange(Lk)[None] >= lens[:, None]).cuda()
    out = torch.full((B, Lq, E), float("nan"), device="cuda")
    m8 = mask.to(torch.uint8).contiguous() if masked else None
    ws = torch.empty(cuda_lib.load().bd_attention_tc_workspace_bytes(B, H, Lq, Lk, split), dtype=torch.uint8, device="cuda")
    cuda_lib.call("bd_attention_tc", q.data_ptr(), 2 * E, Lq * 2 * E, k.data_ptr(), 2 * E, Lk * 2 * E,
                  v.data_ptr(), 2 * E, Lk * 2 * E, cuda_lib.ptr(m8), out.data_ptr(), E, Lq * E, B, H, Lq, Lk, hd,
                  1.0 / math.sqrt(hd), split, ws.data_ptr())
    qh = q.reshape(B, Lq, H, hd).transpose(1, 2).double()
    kh = k.reshape(B, Lk, H, hd).transpose(1, 2).double()
    vh = v.reshape(B, Lk, H, hd).transpose(1, 2).double()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    want = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    cuda_lib.load().bd_attention_tc_select(1)
    tol = 5e-3 if split == 1 else 1e-4
    torch.testing.assert_close(out, want, rtol=tol, atol=tol)


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("B,n,m,ns,C,widths", [(2, 1000, 64, 64, 3, (64, 64, 128)), (2, 777, 33, 32, 3, (64, 64, 64)), (40, 3000, 512, 64, 3, (64, 64, 128)),
                                               (2, 512, 96, 32, 128, (128, 128, 256)),
                                               (1, 300, 40, 16, 256, (128, 128, 256)), (3, 200, 24, 16, 8, (64, 128, 64))])
def test_sa_mlp_tc(cuda_lib, B, n, m, ns, C, widths, split):
    """Fused QueryAndGroup + 3-layer SharedMLP + max-pool vs the same maths in torch (fp64)."""
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(n + m * 3 + C)
    pc = torch.randn(B, n, 3 + C, device="cuda", generator=g)  # [xyz | feats] rows, like the (B,N,6) cloud
    xyz, feats = pc[..., :3], pc[..., 3:]
    idx = torch.randint(0, n, (B, m, ns), device="cuda", generator=g, dtype=torch.int32)
    new_xyz = torch.randn(B, m, 3, device="cuda", generator=g)
    radius = 0.7
    K = [C + 3, widths[0], widths[1]]
    Ws = [torch.randn(widths[i], K[i], device="cuda", generator=g) / math.sqrt(K[i]) for i in range(3)]
    bs = [torch.randn(widths[i], device="cuda", generator=g) * 0.1 for i in range(3)]
    kp = (C + 3 + 7) // 8 * 8
    W0 = F.pad(torch.cat([Ws[0][:, 3:], Ws[0][:, :3]], 1), (0, kp - (C + 3)))  # [feats | xyz | 0]
    Wp = [pack_weight_tc(w, split, full_rows=True)[0] for w in (W0, Ws[1], Ws[2])]
    out = torch.full((B * m, widths[2]), float("nan"), device="cuda")
    cuda_lib.call("bd_sa_mlp_tc", idx.data_ptr(), feats.data_ptr(), 3 + C, C, xyz.data_ptr(), 3 + C, new_xyz.data_ptr(),
                  B, n, m, ns, radius, Wp[0].data_ptr(), bs[0].data_ptr(), widths[0], Wp[1].data_ptr(), bs[1].data_ptr(),
                  widths[1], Wp[2].data_ptr(), bs[2].data_ptr(), widths[2], out.data_ptr(), widths[2], split)
    bi = torch.arange(B, device="cuda")[:, None, None]
    gx = (xyz[bi, idx.long()] - new_xyz[:, :, None, :]) / radius
    x = torch.cat([gx, feats[bi, idx.long()]], -1).double()  # reference order [xyz | feats]
    for i in range(3):
        if split == 1:
            x = x.float().half().double()
            w = Ws[i].half().double()
        else:
            w = Ws[i].double()
        x = torch.relu(x @ w.T + bs[i].double())
    want = x.max(2).values.reshape(B * m, -1).float()
    tol = 1e-2 if split == 1 else 2e-4
    torch.testing.assert_close(out, want, rtol=tol, atol=tol)
    if C % 8 == 0:
        # 16-bit feature rows between the levels (bd_sa_mlp_tc_h): fp16 gather source + fp16 copy of the pooled rows;
        # in the fp16 mode the result must be the SAME bits (the fp32 entry rounds the features to fp16 itself)
        f16 = feats.half().contiguous()
        xyz_c = xyz.contiguous()
        src32 = f16.float() if split == 1 else feats.contiguous()
        ref = torch.full_like(out, float("nan"))
        cuda_lib.call("bd_sa_mlp_tc", idx.data_ptr(), src32.contiguous().data_ptr(), C, C, xyz_c.data_ptr(), 3, new_xyz.data_ptr(),
                      B, n, m, ns, radius, Wp[0].data_ptr(), bs[0].data_ptr(), widths[0], Wp[1].data_ptr(), bs[1].data_ptr(),
                      widths[1], Wp[2].data_ptr(), bs[2].data_ptr(), widths[2], ref.data_ptr(), widths[2], split)
        out2 = torch.full_like(out, float("nan"))
        out16 = torch.full((B * m, widths[2]), float("nan"), device="cuda", dtype=torch.float16)
        cuda_lib.call("bd_sa_mlp_tc_h", idx.data_ptr(), f16.data_ptr(), C, C, 1, xyz_c.data_ptr(), 3, new_xyz.data_ptr(),
                      B, n, m, ns, radius, Wp[0].data_ptr(), bs[0].data_ptr(), widths[0], Wp[1].data_ptr(), bs[1].data_ptr(),
                      widths[1], Wp[2].data_ptr(), bs[2].data_ptr(), widths[2], out2.data_ptr(), widths[2], out16.data_ptr(),
                      widths[2], split)
        if split == 1:
            assert torch.equal(out2, ref)
        else:
            torch.testing.assert_close(out2, ref, rtol=5e-3, atol=5e-3)  # fp16-rounded features vs fp32 features
        assert torch.equal(out16, out2.half())


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("M,N,K,add,wide", [(20000, 576, 288, False, True), (20000, 576, 288, True, False),
                                            (33000, 288, 288, False, False), (19999, 256, 264, True, True)])
def test_linear_tc_many_row_tiles(cuda_lib, M, N, K, add, wide, split):
    """More row tiles than SMs (several waves, the two-CTAs-per-SM policy), a ragged last tile and a
    K tail — the tensor copies of the activations zero-fill both; narrow and wide tilings."""
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(M + N + K + 1)
    A = torch.randn(M, K, device="cuda", generator=g)
    A2 = torch.randn(M, K, device="cuda", generator=g) if add else None
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split, wide=wide and not (add and split == 3))
    Y = torch.full((M, N), float("nan"), device="cuda")
    cuda_lib.call("bd_linear_tc", A.data_ptr(), K, cuda_lib.ptr(A2), K, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N,
                  M, N, K, KC, nch, BN, nsub, 1, split)
    a = A + A2 if add else A
    if split == 1:
        want = F.linear(a.half().double(), W.half().double(), b.double())
    else:
        want = F.linear(a.double(), W.double(), b.double())
    torch.testing.assert_close(Y, want.relu().float(), rtol=1e-4, atol=1e-4)


def test_launch_policies_do_not_change_results(cuda_lib):
    """Programmatic dependent launch on / off and one / two CTAs per SM: bit-identical outputs."""
    from butd_detr_b200.engine import pack_weight_tc
    lib = cuda_lib.load()
    M, N, K = 40000, 288, 288
    g = _g(7)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    outs = []
    try:
        for pdl, occ in ((1, 1), (0, 1), (1, 0), (0, 0)):
            lib.bd_set_pdl(pdl)
            lib.bd_linear_tc_set_occupancy(occ)
            Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 3)
            Y = torch.full((M, N), float("nan"), device="cuda")
            H = torch.empty(M, N, device="cuda")
            # a dependent chain: the second kernel reads what the first one wrote
            cuda_lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), H.data_ptr(), N, M, N, K,
                          KC, nch, BN, nsub, 1, 3)
            cuda_lib.call("bd_linear_tc", H.data_ptr(), N, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K,
                          KC, nch, BN, nsub, 0, 3)
            torch.cuda.synchronize()
            outs.append(Y)
    finally:
        lib.bd_set_pdl(1)
        lib.bd_linear_tc_set_occupancy(1)
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert torch.isfinite(outs[0]).all()


def test_fp16_operands_saturate_instead_of_overflowing(cuda_lib):
    """fp16 operand mode: activations / weights beyond 65504 saturate (cvt.rn.satfinite) — the result
    stays finite and equals the product of the clamped operands; NaN inputs still propagate."""
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(77)
    M, N, K = 256, 64, 64
    A = torch.randn(M, K, device="cuda", generator=g)
    A[3, 5], A[100, 0], A[200, 63] = 3.0e5, -7.0e4, 65504.0
    W = torch.randn(N, K, device="cuda", generator=g) / 8
    W[7, 5] = 1.0e6
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1)
    Y = torch.full((M, N), float("nan"), device="cuda")
    cuda_lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), None, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, 1)
    assert bool(torch.isfinite(Y).all())
    want = (A.clamp(-65504, 65504).half().double() @ W.clamp(-65504, 65504).half().double().t()).float()
    torch.testing.assert_close(Y, want, rtol=1e-4, atol=1e-2)
    A[9, 9] = float("nan")
    cuda_lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), None, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, 1)
    assert bool(torch.isnan(Y[9]).all()) and bool(torch.isfinite(Y[8]).all())


@pytest.mark.parametrize("precision", ["bf16x3", "fp16"])
def test_unfused_sa_level_path_matches_fused(cuda_lib, precision):
    """engine.FUSED_SA = False routes the set-abstraction levels through bd_group_rows / bd_sa_group_linear_tc
    -> bd_linear_tc -> bd_linear_pool_tc instead of the one-kernel bd_sa_mlp_tc: same features."""
    from butd_detr_b200 import BeaUTyDETR, engine, synth
    model = BeaUTyDETR(num_queries=32, num_decoder_layers=1, num_encoder_layers=1, text_encoder=None, precision=precision)
    synth.fill_state_dict_(model.state_dict(), 0)
    model = model.cuda().eval()
    inputs = {k: v.cuda() for k, v in synth.synth_batch(23, 2, 4096, 16, 32).items()}
    fused = {k: v.clone() for k, v in model(inputs).items() if torch.is_tensor(v)}
    engine.FUSED_SA = False
    try:
        model.invalidate_engine()
        unfused = model(inputs)
    finally:
        engine.FUSED_SA = True
        model.invalidate_engine()
    tol = 2e-3 if precision == "bf16x3" else 3e-2  # fp2_features are O(10); hidden activations are re-rounded between kernels
    for k in ("sa1_features", "sa2_features", "sa3_features", "sa4_features", "fp2_features"):
        torch.testing.assert_close(unfused[k], fused[k], rtol=tol, atol=tol)
    for k in ("sa1_inds", "sa2_inds"):
        assert torch.equal(unfused[k], fused[k])


@pytest.mark.parametrize("small_nk", [0, 8])
@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 1024, 80, True), (3, 256, 132, True), (2, 300, 1000, True),
                                             (2, 80, 1024, False), (1, 129, 257, False), (5, 80, 80, True)])
def test_attention_tc_tile_variants(cuda_lib, B, Lq, Lk, masked, split, small_nk):
    """Both CTA shapes of the warp-specialised kernel on every sequence length: two ping-ponged query
    tiles per CTA (small_nk = 0) and one query tile per CTA with two CTAs per SM (small_nk = 8)."""
    H, hd = 8, 36
    E = H * hd
    lib = cuda_lib.load()
    g = _g(Lq * 11 + Lk + small_nk)
    q = torch.randn(B, Lq, E, device="cuda", generator=g)
    kv = torch.randn(B, Lk, 2 * E, device="cuda", generator=g)
    k, v = kv[..., :E], kv[..., E:]
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=torch.Generator().manual_seed(Lk))
        mask = (torch.arange(Lk)[None] >= lens[:, None]).cuda()
    out = torch.full((B, Lq, E), float("nan"), device="cuda")
    m8 = mask.to(torch.uint8).contiguous() if masked else None
    ws = torch.empty(lib.bd_attention_tc_workspace_bytes(B, H, Lq, Lk, split), dtype=torch.uint8, device="cuda")
    lib.bd_attention_tc_set_small_nk(small_nk)
    try:
        cuda_lib.call("bd_attention_tc", q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E,
                      Lk * 2 * E, cuda_lib.ptr(m8), out.data_ptr(), E, Lq * E, B, H, Lq, Lk, hd, 1.0 / math.sqrt(hd), split,
                      ws.data_ptr())
    finally:
        lib.bd_attention_tc_set_small_nk(1 << 30)
    qh = q.reshape(B, Lq, H, hd).transpose(1, 2).double()
    kh = k.reshape(B, Lk, H, hd).transpose(1, 2).double()
    vh = v.reshape(B, Lk, H, hd).transpose(1, 2).double()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    want = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    tol = 5e-3 if split == 1 else 1e-4
    torch.testing.assert_close(out, want, rtol=tol, atol=tol)


@pytest.mark.parametrize("a_half,y_half", [(1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K,relu", [(1024, 576, 288, False), (300, 256, 288, True), (131, 288, 256, False),
                                         (4099, 64, 8, True), (2048, 864, 288, False), (77, 3, 288, False)])
def test_linear_tc_half_activations(cuda_lib, M, N, K, relu, a_half, y_half):
    """fp16 activations in HBM: A read by tensor copy straight into the operand layout, Y written as fp16
    rows — same values as the fp32 entry point up to the output rounding."""
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(M + N + K + a_half * 2 + y_half)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1)
    Ain = A.half().contiguous() if a_half else A
    ldy = (N + 7) // 8 * 8
    Y = torch.full((M, ldy), float("nan"), device="cuda", dtype=torch.float16 if y_half else torch.float32)
    cuda_lib.call("bd_linear_tc_h", Ain.data_ptr(), K, a_half, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), ldy, y_half,
                  M, N, K, KC, nch, BN, nsub, int(relu))
    want = F.linear(A.half().double(), W.half().double(), b.double())
    want = (want.relu() if relu else want).float()
    tol = 2e-3 if y_half else 1e-4
    torch.testing.assert_close(Y[:, :N].float(), want, rtol=tol, atol=tol)


@pytest.mark.parametrize("a_half,shadow", [(1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(1024, 288, 256), (300, 288, 288), (2048, 160, 64)])
def test_linear_ln_tc_half_activations(cuda_lib, M, N, K, a_half, shadow):
    from butd_detr_b200.engine import pack_weight_tc
    g = _g(M * 5 + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g)
    gam, bet = torch.rand(N, device="cuda", generator=g) + 0.5, torch.randn(N, device="cuda", generator=g)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, full_rows=True)
    Ain = A.half().contiguous() if a_half else A
    Y = torch.full((M, N), float("nan"), device="cuda")
    Y16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16) if shadow else None
    cuda_lib.call("bd_linear_ln_tc_h", Ain.data_ptr(), K, a_half, Wp.data_ptr(), b.data_ptr(), R.data_ptr(), N,
                  gam.data_ptr(), bet.data_ptr(), 1e-5, Y.data_ptr(), N, cuda_lib.ptr(Y16), N, M, N, K, KC, nch, BN, nsub)
    lin = F.linear(A.half().double(), W.half().double(), b.double())
    want = F.layer_norm(R.double() + lin, (N,), gam.double(), bet.double(), 1e-5).float()
    torch.testing.assert_close(Y, want, rtol=2e-4, atol=2e-4)
    if shadow:
        assert torch.equal(Y16, Y.half())


@pytest.mark.parametrize("direct", [1, 0])
@pytest.mark.parametrize("io", [15, 8, 9, 6, 14])
@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 1024, 1024, False), (3, 256, 132, True), (2, 80, 80, True), (1, 300, 1000, True)])
def test_attention_tc_half_tensors(cuda_lib, B, Lq, Lk, masked, io, direct):
    """fp16 Q / K / V / O in HBM (io bits 0..3), fused wider buffers as in the engine.  direct = 1: with fp16 K and V
    (io bits 1, 2) the kernel reads its K / V tiles from these rows by tensor copies (no pack kernel, NULL
    workspace); direct = 0: the pack kernel."""
    if direct == 0 and (io & 6) != 6:
        pytest.skip("same path as direct = 1")
    H, hd = 8, 36
    E = H * hd
    lib = cuda_lib.load()
    g = _g(Lq * 13 + Lk + io)
    q32 = torch.randn(B, Lq, E, device="cuda", generator=g)
    kv32 = torch.randn(B, Lk, 2 * E, device="cuda", generator=g)
    q = q32.half() if io & 1 else q32
    kbuf = kv32.half() if io & 2 else kv32
    vbuf = kv32.half() if io & 4 else kv32
    k, v = kbuf[..., :E], vbuf[..., E:]
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=torch.Generator().manual_seed(Lk))
        mask = (torch.arange(Lk)[None] >= lens[:, None]).cuda()
    out = torch.full((B, Lq, E), float("nan"), device="cuda", dtype=torch.float16 if io & 8 else torch.float32)
    m8 = mask.to(torch.uint8).contiguous() if masked else None
    no_ws = direct and (io & 6) == 6
    ws = None if no_ws else torch.empty(lib.bd_attention_tc_workspace_bytes(B, H, Lq, Lk, 1), dtype=torch.uint8, device="cuda")
    lib.bd_attention_tc_set_direct(direct)
    try:
        cuda_lib.call("bd_attention_tc_h", q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E,
                      Lk * 2 * E, cuda_lib.ptr(m8), out.data_ptr(), E, Lq * E, io, B, H, Lq, Lk, hd, 1.0 / math.sqrt(hd), 1,
                      cuda_lib.ptr(ws))
    finally:
        lib.bd_attention_tc_set_direct(1)
    qh = q32.reshape(B, Lq, H, hd).transpose(1, 2).double()
    kh = kv32[..., :E].reshape(B, Lk, H, hd).transpose(1, 2).double()
    vh = kv32[..., E:].reshape(B, Lk, H, hd).transpose(1, 2).double()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    want = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    torch.testing.assert_close(out.float(), want, rtol=6e-3, atol=6e-3)


@pytest.mark.parametrize("M,N,K,relu", [(151552 // 4, 288, 288, 0), (40000, 576, 288, 1), (33333, 256, 288, 1), (148 * 256, 864, 288, 1),
                                        (20000, 288, 256, 0), (19000, 64, 288, 1)])
def test_persistent_linear_equals_the_one_tile_kernel(cuda_lib, M, N, K, relu):
    """fp16 rows in and out with more row tiles than SMs: the persistent kernel (gemm_stream.cu) must give the bits
    of the one-tile-per-CTA kernel (same operand formats, MMA order and epilogue), ragged last tile included."""
    from butd_detr_b200.engine import lin_tiling, pack_weight_tc
    lib = cuda_lib.load()
    g = _g(M + N)
    A = (torch.randn(M, K, device="cuda", generator=g)).half()
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    wide, bn = lin_tiling(M, N)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, wide=wide, bn=bn)
    outs = []
    for on in (0, 1):
        lib.bd_linear_stream_set(on)
        try:
            Y = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16)
            cuda_lib.call("bd_linear_tc_h", A.data_ptr(), K, 1, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, 1, M, N, K,
                          KC, nch, BN, nsub, relu)
            torch.cuda.synchronize()
        finally:
            lib.bd_linear_stream_set(1)
        outs.append(Y)
    assert torch.equal(outs[0], outs[1])
    want = F.linear(A.double(), W.half().double(), b.double())
    if relu:
        want = want.relu()
    torch.testing.assert_close(outs[1].double(), want, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("M,N,K,shadow", [(151552 // 4, 288, 288, True), (148 * 256 + 77, 288, 256, True), (20000, 288, 288, False),
                                          (30001, 256, 320, True)])
def test_persistent_linear_layernorm_equals_the_one_tile_kernel(cuda_lib, M, N, K, shadow):
    """LayerNorm(R + A16 W^T + b) with more row tiles than SMs: the persistent kernel (two 64-row halves per tile)
    must give the bits of linear_tc_kernel<1, 0>, fp16 copy included, ragged last tile included."""
    from butd_detr_b200.engine import pack_weight_tc
    lib = cuda_lib.load()
    g = _g(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    R = torch.randn(M, N, device="cuda", generator=g)
    gam, bet = torch.rand(N, device="cuda", generator=g) + 0.5, torch.randn(N, device="cuda", generator=g) * 0.1
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, full_rows=True)
    outs = []
    for on in (0, 1):
        lib.bd_linear_stream_set(on)
        try:
            Y = torch.full((M, N), float("nan"), device="cuda")
            Y16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16) if shadow else None
            cuda_lib.call("bd_linear_ln_tc_h", A.data_ptr(), K, 1, Wp.data_ptr(), b.data_ptr(), R.data_ptr(), N, gam.data_ptr(),
                          bet.data_ptr(), 1e-5, Y.data_ptr(), N, cuda_lib.ptr(Y16), N, M, N, K, KC, nch, BN, nsub)
            torch.cuda.synchronize()
        finally:
            lib.bd_linear_stream_set(1)
        outs.append((Y, Y16))
    assert torch.equal(outs[0][0], outs[1][0])
    if shadow:
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[1][1], outs[1][0].half())
    want = F.layer_norm(R.double() + F.linear(A.double(), W.half().double(), b.double()), (N,), gam.double(), bet.double(), 1e-5).float()
    torch.testing.assert_close(outs[1][0], want, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 1024, 80, True), (3, 256, 128, False), (2, 80, 80, True), (1, 300, 33, True),
                                            (2, 130, 96, False), (1, 128, 1, False), (3, 256, 132, True), (2, 1024, 132, False),
                                            (2, 200, 136, True), (1, 64, 129, True)])
def test_attention_short_keys_kernel(cuda_lib, B, Lq, Lk, masked):
    """Lk <= 128 (+ up to 8 keys handled on the FMA pipe: Lk = 132 detected boxes) with fp16 K / V: the
    four-CTAs-per-SM kernel (P in shared memory, O over the score columns) against fp64 torch and against the general
    kernel (same arithmetic: equal up to the accumulation order)."""
    H, hd = 8, 36
    E = H * hd
    lib = cuda_lib.load()
    g = _g(Lq * 7 + Lk)
    q32 = torch.randn(B, Lq, E, device="cuda", generator=g)
    kv32 = torch.randn(B, Lk, 2 * E, device="cuda", generator=g)
    q, kv = q32.half(), kv32.half()
    k, v = kv[..., :E], kv[..., E:]
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=torch.Generator().manual_seed(Lk))
        lens[0] = Lk
        mask = (torch.arange(Lk)[None] >= lens[:, None]).cuda()
    m8 = mask.to(torch.uint8).contiguous() if masked else None
    outs = []
    for on in (1, 0):
        lib.bd_attention_tc_set_short(on)
        try:
            out = torch.full((B, Lq, E), float("nan"), device="cuda", dtype=torch.float16)
            cuda_lib.call("bd_attention_tc_h", q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E,
                          Lk * 2 * E, cuda_lib.ptr(m8), out.data_ptr(), E, Lq * E, 15, B, H, Lq, Lk, hd, 1.0 / math.sqrt(hd), 1, None)
            torch.cuda.synchronize()
        finally:
            lib.bd_attention_tc_set_short(1)
        outs.append(out)
    qh = q.double().reshape(B, Lq, H, hd).transpose(1, 2)
    kh = k.double().reshape(B, Lk, H, hd).transpose(1, 2)
    vh = v.double().reshape(B, Lk, H, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    want = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    torch.testing.assert_close(outs[0].float(), want, rtol=6e-3, atol=6e-3)
    torch.testing.assert_close(outs[0].float(), outs[1].float(), rtol=2e-3, atol=2e-3)
