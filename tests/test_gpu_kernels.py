"""GPU numerics of the fused forward kernels (C-ABI, fp32 path) against plain PyTorch fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("M,N,K,relu,add", [(131, 64, 8, True, False), (4099, 128, 132, True, False),
                                             (1024, 576, 288, False, True), (80, 288, 288, False, False),
                                             (256, 3, 288, False, False), (1024, 1, 288, False, False),
                                             (300, 288, 6, True, False), (513, 160, 768, False, False),
                                             (9000, 128, 64, True, False)])
def test_linear(cuda_lib, M, N, K, relu, add):
    g = _g(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    A2 = torch.randn(M, K, device="cuda", generator=g) if add else None
    W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
    b = torch.randn(N, device="cuda", generator=g)
    Y = torch.empty(M, N, device="cuda")
    cuda_lib.call("bd_linear_f32", A.data_ptr(), K, cuda_lib.ptr(A2), K, W.data_ptr(), b.data_ptr(), Y.data_ptr(), N,
                  M, N, K, int(relu))
    want = F.linear((A + A2 if add else A).double(), W.double(), b.double())
    want = (want.relu() if relu else want).float()
    torch.testing.assert_close(Y, want, rtol=1e-5, atol=2e-5)


def test_linear_strided_views(cuda_lib):
    """Column-slice inputs/outputs (leading dimensions), as used for fused QKV / det features."""
    g = _g(5)
    big = torch.randn(200, 864, device="cuda", generator=g)
    W = torch.randn(160, 288, device="cuda", generator=g) / 17
    out = torch.zeros(200, 288, device="cuda")
    x = big[:, 288:576]
    cuda_lib.call("bd_linear_f32", x.data_ptr(), 864, None, 0, W.data_ptr(), None, out[:, 128:].data_ptr(), 288,
                  200, 160, 288, 0)
    torch.testing.assert_close(out[:, 128:], x @ W.t(), rtol=1e-5, atol=2e-5)
    assert float(out[:, :128].abs().max()) == 0.0


@pytest.mark.parametrize("B,Lq,Lk,masked", [(2, 1024, 1024, False), (2, 80, 1024, False), (2, 1024, 80, True),
                                             (3, 256, 132, True), (1, 32, 16, True), (2, 70, 65, True)])
def test_attention(cuda_lib, B, Lq, Lk, masked):
    H, hd = 8, 36
    E = H * hd
    g = _g(Lq * 7 + Lk)
    qkv_q = torch.randn(B, Lq, 2 * E, device="cuda", generator=g)  # q lives in a wider fused buffer
    kv = torch.randn(B, Lk, 2 * E, device="cuda", generator=g)
    q, k, v = qkv_q[..., :E], kv[..., :E], kv[..., E:]
    mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=torch.Generator().manual_seed(Lk))
        mask = (torch.arange(Lk)[None] >= lens[:, None]).cuda()
    out = torch.empty(B, Lq, E, device="cuda")
    m8 = mask.to(torch.uint8).contiguous() if masked else None
    cuda_lib.call("bd_attention_f32", q.data_ptr(), 2 * E, Lq * 2 * E, k.data_ptr(), 2 * E, Lk * 2 * E,
                  v.data_ptr(), 2 * E, Lk * 2 * E, cuda_lib.ptr(m8), out.data_ptr(), E, Lq * E, B, H, Lq, Lk, hd,
                  1.0 / math.sqrt(hd))
    qh = q.reshape(B, Lq, H, hd).transpose(1, 2).double()
    kh = k.reshape(B, Lk, H, hd).transpose(1, 2).double()
    vh = v.reshape(B, Lk, H, hd).transpose(1, 2).double()
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if masked:
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    want = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E).float()
    torch.testing.assert_close(out, want, rtol=1e-4, atol=2e-5)


def test_attention_fully_masked_row_is_nan_like_reference(cuda_lib):
    H, hd, E = 8, 36, 288
    q = torch.randn(1, 4, E, device="cuda")
    kv = torch.randn(1, 5, E, device="cuda")
    mask = torch.ones(1, 5, dtype=torch.uint8, device="cuda")
    out = torch.zeros(1, 4, E, device="cuda")
    cuda_lib.call("bd_attention_f32", q.data_ptr(), E, 4 * E, kv.data_ptr(), E, 5 * E, kv.data_ptr(), E, 5 * E,
                  mask.data_ptr(), out.data_ptr(), E, 4 * E, 1, H, 4, 5, hd, 1 / 6.0)
    assert bool(torch.isnan(out).all())


@pytest.mark.parametrize("M,D,eps,res", [(1024, 288, 1e-5, True), (77, 288, 1e-12, False), (5, 64, 1e-5, True)])
def test_add_layernorm(cuda_lib, M, D, eps, res):
    g = _g(M)
    x = torch.randn(M, D, device="cuda", generator=g) * 3
    r = torch.randn(M, D, device="cuda", generator=g) if res else None
    w, b = torch.rand(D, device="cuda", generator=g) + 0.5, torch.randn(D, device="cuda", generator=g)
    y = torch.empty_like(x)
    cuda_lib.call("bd_add_layernorm_f32", x.data_ptr(), cuda_lib.ptr(r), w.data_ptr(), b.data_ptr(), y.data_ptr(),
                  M, D, eps)
    want = F.layer_norm(x + r if res else x, (D,), w, b, eps)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)


def test_topk_sigmoid_matches_torch(cuda_lib):
    g = _g(3)
    logits = torch.randn(4, 1024, device="cuda", generator=g)
    logits[1, 100] = logits[1, 7]  # an exact tie: lower index first
    idx = torch.empty(4, 256, dtype=torch.int32, device="cuda")
    cuda_lib.call("bd_topk_sigmoid", logits.data_ptr(), 4, 1024, 256, idx.data_ptr())
    s = torch.sigmoid(logits)
    want_vals = torch.topk(s, 256)[0]
    got_vals = torch.gather(s, 1, idx.long())
    assert torch.equal(got_vals, want_vals)                    # same multiset, same descending order
    assert bool((got_vals[:, :-1] >= got_vals[:, 1:]).all())
    assert len(set(idx[1].tolist())) == 256
    pos7, pos100 = (idx[1] == 7).nonzero(), (idx[1] == 100).nonzero()
    if len(pos7) and len(pos100):
        assert int(pos7) < int(pos100)


def test_small_row_ops(cuda_lib):
    g = _g(11)
    x = torch.randn(333, 64, device="cuda", generator=g)
    y = torch.empty_like(x)
    cuda_lib.call("bd_l2_normalize_rows", x.data_ptr(), y.data_ptr(), 333, 64)
    torch.testing.assert_close(y, F.normalize(x, p=2, dim=-1), rtol=1e-6, atol=1e-6)
    table = torch.randn(485, 768, device="cuda", generator=g)
    ids = torch.randint(0, 485, (200,), device="cuda", generator=g)
    out = torch.empty(200, 768, device="cuda")
    cuda_lib.call("bd_embedding_rows", table.data_ptr(), 768, ids.data_ptr(), 200, out.data_ptr(), 768)
    assert torch.equal(out, table[ids])
    a, b = torch.randn(50, 3, device="cuda", generator=g), torch.randn(50, 3, device="cuda", generator=g)
    c = torch.empty(50, 6, device="cuda")
    cuda_lib.call("bd_concat_rows", a.data_ptr(), 3, 3, b.data_ptr(), 3, 3, c.data_ptr(), 6, 50)
    assert torch.equal(c, torch.cat([a, b], -1))
    s = torch.empty(50, 3, device="cuda")
    cuda_lib.call("bd_add_rows", a.data_ptr(), 3, b.data_ptr(), 3, s.data_ptr(), 3, 50, 3)
    assert torch.equal(s, a + b)
    t_in = torch.randn(2, 100, 37, device="cuda", generator=g)
    t_out = torch.empty(2, 37, 100, device="cuda")
    cuda_lib.call("bd_transpose_rows", t_in.data_ptr(), 2, 100, 37, t_out.data_ptr())
    assert torch.equal(t_out, t_in.transpose(1, 2).contiguous())


def test_group_maxpool_fp_rows_match_oracle(cuda_lib, oracle_lib):
    """Token-major QueryAndGroup / max-pool / FP-interpolate vs the oracle's channel-major maths."""
    from oracle import model_ref
    from pointops_cases import cloud
    xyz = cloud(8, 2000, "room", 2)
    g = torch.Generator().manual_seed(4)
    feats = torch.randn(2, 7, 2000, generator=g)                       # (B,C,n) reference layout
    inds = oracle_lib.furthest_point_sampling(xyz, 100)
    new_xyz = torch.gather(xyz, 1, inds.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    want = model_ref.query_and_group(xyz, new_xyz, feats, 0.4, 16)    # (B,10,100,16)
    idx = oracle_lib.ball_query(new_xyz, xyz, 0.4, 16).cuda()
    feats_tm = feats.transpose(1, 2).contiguous().cuda()
    out = torch.empty(2 * 100 * 16, 12, device="cuda")
    xyz_d, new_xyz_d = xyz.cuda(), new_xyz.cuda()
    cuda_lib.call("bd_group_rows", xyz_d.data_ptr(), 3, feats_tm.data_ptr(), 7, 7, new_xyz_d.data_ptr(),
                  idx.data_ptr(), 2, 2000, 100, 16, 0.4, out.data_ptr(), 12)
    got = out.view(2, 100, 16, 12)
    torch.testing.assert_close(got[..., :10].permute(0, 3, 1, 2).cpu(), want, rtol=1e-6, atol=1e-6)
    assert float(got[..., 10:].abs().max()) == 0.0
    pooled = torch.empty(200, 12, device="cuda")
    cuda_lib.call("bd_maxpool_rows", out.data_ptr(), 200, 16, 12, pooled.data_ptr())
    assert torch.equal(pooled.view(2, 100, 12), got.max(2)[0])
    # FP: interpolate + concat
    unknown, known = cloud(1, 512, "room", 2), cloud(2, 256, "room", 2)
    kf, uf = torch.randn(2, 9, 256, generator=g), torch.randn(2, 5, 512, generator=g)
    d2, i3 = oracle_lib.three_nn(unknown, known)
    dist = torch.sqrt(d2)
    rec = 1.0 / (dist + 1e-8)
    w = (rec / rec.sum(2, keepdim=True)).contiguous()
    want = torch.cat([oracle_lib.three_interpolate(kf, i3, w), uf], 1)      # (B,14,512)
    x = torch.empty(2 * 512, 14, device="cuda")
    d2_d, i3_d = d2.cuda(), i3.cuda()
    kf_d, uf_d = kf.transpose(1, 2).contiguous().cuda(), uf.transpose(1, 2).contiguous().cuda()
    cuda_lib.call("bd_fp_interp_concat", d2_d.data_ptr(), i3_d.data_ptr(), kf_d.data_ptr(), 9, uf_d.data_ptr(), 5,
                  2, 512, 256, x.data_ptr())
    torch.testing.assert_close(x.view(2, 512, 14).transpose(1, 2).cpu(), want, rtol=1e-6, atol=1e-6)


def test_fp_interp_concat_rows_kernel_equals_the_element_kernel(cuda_lib, oracle_lib):
    """Channel counts that are multiples of 4 take the warp-per-row kernel (weights once per row, 16-byte vectors):
    the fp32 rows must equal the oracle's three_interpolate + concat bit for bit (same operation order), the fp16
    rows (bd_fp_interp_concat_h) must be those values rounded once."""
    from pointops_cases import cloud
    g = torch.Generator().manual_seed(4)
    B, n, m, C2, C1 = 3, 300, 77, 24, 16
    unknown, known = cloud(5, n, "room", B), cloud(6, m, "room", B)
    kf, uf = torch.randn(B, C2, m, generator=g), torch.randn(B, C1, n, generator=g)
    d2, i3 = oracle_lib.three_nn(unknown, known)
    rec = 1.0 / (torch.sqrt(d2) + 1e-8)
    w = (rec / rec.sum(2, keepdim=True)).contiguous()
    want = torch.cat([oracle_lib.three_interpolate(kf, i3, w), uf], 1).transpose(1, 2).contiguous()  # (B, n, C2+C1)
    d2_d, i3_d = d2.cuda(), i3.cuda()
    kf_d, uf_d = kf.transpose(1, 2).contiguous().cuda(), uf.transpose(1, 2).contiguous().cuda()
    x = torch.full((B * n, C2 + C1), float("nan"), device="cuda")
    cuda_lib.call("bd_fp_interp_concat", d2_d.data_ptr(), i3_d.data_ptr(), kf_d.data_ptr(), C2, uf_d.data_ptr(), C1, B, n, m,
                  x.data_ptr())
    torch.testing.assert_close(x.view(B, n, -1).cpu(), want, rtol=1e-6, atol=1e-6)
    x16 = torch.full((B * n, C2 + C1), float("nan"), device="cuda", dtype=torch.float16)
    cuda_lib.call("bd_fp_interp_concat_h", d2_d.data_ptr(), i3_d.data_ptr(), kf_d.data_ptr(), C2, uf_d.data_ptr(), C1, B, n, m,
                  x16.data_ptr(), 1)
    assert torch.equal(x16, x.half())


@pytest.mark.parametrize("M,N,K,relu,half", [(5000, 288, 3, 1, 1), (777, 288, 6, 1, 0), (129, 128, 6, 0, 1), (64, 30, 8, 1, 0)])
def test_linear_smallk(cuda_lib, M, N, K, relu, half):
    """Narrow-input layer (first layer of the position embeddings): fp32 rows equal bd_linear_f32's (same operation
    order), fp16 rows are those values rounded once."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    ref = torch.full((M, N), float("nan"), device="cuda")
    cuda_lib.call("bd_linear_f32", A.data_ptr(), K, None, 0, W.data_ptr(), b.data_ptr(), ref.data_ptr(), N, M, N, K, relu)
    Y = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16 if half else torch.float32)
    cuda_lib.call("bd_linear_smallk", A.data_ptr(), K, W.data_ptr(), b.data_ptr(), Y.data_ptr(), N, half, M, N, K, relu)
    want = F.linear(A.double(), W.double(), b.double())
    if relu:
        want = want.relu()
    torch.testing.assert_close(ref.double(), want, rtol=1e-5, atol=1e-5)
    if half:
        assert torch.equal(Y, ref.half())
    else:
        assert torch.equal(Y, ref)
