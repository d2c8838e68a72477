"""compute-sanitizer target for the kernels added in the second half of round 2 only (racecheck on the whole forward
runs into the hazard-count limit on the SA kernels' mbarrier-ordered stores): one-key-tile attention, tensor-copy
attention (head_dim 36 and 64), persistent linear / linear + LayerNorm, narrow-input linear, warp-per-row
interpolate + concat, RoBERTa embeddings, matcher.   python tools/sanitize_new_kernels.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from butd_detr_b200 import _lib  # noqa: E402
from butd_detr_b200.engine import lin_tiling, pack_weight_tc  # noqa: E402
from butd_detr_b200.matcher import HungarianMatcher  # noqa: E402

lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
H, hd = 8, 36
E = H * hd
for (B, Lq, Lk, hdim, heads) in ((2, 256, 80, 36, 8), (2, 200, 132, 36, 8), (1, 256, 300, 36, 8), (2, 80, 80, 64, 12)):
    Em = heads * hdim
    q = torch.randn(B, Lq, Em, device="cuda", generator=g).half()
    kv = torch.randn(B, Lk, 2 * Em, device="cuda", generator=g).half()
    out = torch.empty(B, Lq, Em, device="cuda", dtype=torch.float16)
    mask = (torch.arange(Lk, device="cuda")[None] >= torch.tensor([Lk, max(1, Lk - 7)], device="cuda")[:B, None]).to(torch.uint8).contiguous()
    _lib.call("bd_attention_tc_h", q.data_ptr(), Em, Lq * Em, kv.data_ptr(), 2 * Em, Lk * 2 * Em, kv[..., Em:].data_ptr(), 2 * Em,
              Lk * 2 * Em, mask.data_ptr(), out.data_ptr(), Em, Lq * Em, 15, B, heads, Lq, Lk, hdim, 1.0 / math.sqrt(hdim), 1, None)
M, N, K = 19200, 288, 288  # 150 row tiles: more than the 148 SMs -> the persistent kernels
A = torch.randn(M, K, device="cuda", generator=g).half()
W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
bias = torch.randn(N, device="cuda", generator=g)
wide, bn = lin_tiling(M, N)
Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, wide=wide, bn=bn)
Y16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
for relu in (1, 2):
    _lib.call("bd_linear_tc_h", A.data_ptr(), K, 1, None, 0, Wp.data_ptr(), bias.data_ptr(), Y16.data_ptr(), N, 1, M, N, K, KC, nch, BN, nsub, relu)
Wr, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, full_rows=True)
R, Y = torch.randn(M, N, device="cuda", generator=g), torch.empty(M, N, device="cuda")
gam, bet = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
_lib.call("bd_linear_ln_tc_h", A.data_ptr(), K, 1, Wr.data_ptr(), bias.data_ptr(), R.data_ptr(), N, gam.data_ptr(), bet.data_ptr(), 1e-5,
          Y.data_ptr(), N, Y16.data_ptr(), N, M, N, K, KC, nch, BN, nsub)
X6 = torch.randn(3000, 6, device="cuda", generator=g)
W6 = torch.randn(288, 6, device="cuda", generator=g)
Y6 = torch.empty(3000, 288, device="cuda", dtype=torch.float16)
_lib.call("bd_linear_smallk", X6.data_ptr(), 6, W6.data_ptr(), bias.data_ptr(), Y6.data_ptr(), 288, 1, 3000, 288, 6, 1)
Bf, n, m, C2, C1 = 2, 300, 77, 256, 128
d2 = torch.rand(Bf, n, 3, device="cuda", generator=g) + 0.01
i3 = torch.randint(0, m, (Bf, n, 3), device="cuda", generator=g, dtype=torch.int32)
kf, uf = torch.randn(Bf, m, C2, device="cuda", generator=g), torch.randn(Bf, n, C1, device="cuda", generator=g)
x16 = torch.empty(Bf * n, C1 + C2, device="cuda", dtype=torch.float16)
_lib.call("bd_fp_interp_concat_h", d2.data_ptr(), i3.data_ptr(), kf.data_ptr(), C2, uf.data_ptr(), C1, Bf, n, m, x16.data_ptr(), 1)
ids = torch.randint(3, 900, (3, 21), device="cuda", generator=g)
ids[1, 15:] = 1
tab = torch.randn(1000, 768, device="cuda", generator=g)
pos = torch.randn(514, 768, device="cuda", generator=g)
typ = torch.randn(1, 768, device="cuda", generator=g)
emb = torch.empty(63, 768, device="cuda")
_lib.call("bd_roberta_embed", ids.data_ptr(), tab.data_ptr(), 1000, pos.data_ptr(), 514, typ.data_ptr(), torch.ones(768, device="cuda").data_ptr(),
          torch.zeros(768, device="cuda").data_ptr(), emb.data_ptr(), 3, 21, 768, 1, 1e-5)
gc = torch.Generator().manual_seed(1)
outs = {"pred_logits": torch.randn(2, 64, 256, generator=gc).cuda(),
        "pred_boxes": torch.cat([torch.rand(2, 64, 3, generator=gc), torch.rand(2, 64, 3, generator=gc) + 0.1], -1).cuda()}
tg = [{"boxes": torch.cat([torch.rand(k, 3, generator=gc), torch.rand(k, 3, generator=gc) + 0.1], -1).cuda(),
       "positive_map": torch.rand(k, 256, generator=gc).cuda(), "labels": torch.zeros(k, dtype=torch.int64).cuda()} for k in (5, 64)]
pairs = HungarianMatcher(1, 5, 2, True)(outs, tg)
torch.cuda.synchronize()
print("new kernels finished:", bool(torch.isfinite(Y).all()), bool(torch.isfinite(out.float()).all()), [int(p[0].sum()) for p in pairs])
