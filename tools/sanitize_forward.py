"""Target of the compute-sanitizer runs (profiles/r02_sanitizer_*.log): one configs[0]-sized forward, eager and
CUDA-graph replay, plus the cell-list kernels (grid build, bucketed FPS, ball query) on a 9000-point cloud.
python tools/sanitize_forward.py [precision] [--graph]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import BeaUTyDETR, _lib, synth

precision = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "fp16"
graph = "--graph" in sys.argv
model = BeaUTyDETR(num_queries=32, num_decoder_layers=1, num_encoder_layers=1, text_encoder=None, precision=precision,
                   cuda_graph=graph)
synth.fill_state_dict_(model.state_dict(), 0)
model = model.cuda().eval()
inputs = {k: v.cuda() for k, v in synth.synth_batch(21, 2, 4096, 16, 32).items()}
for _ in range(2):
    ep = model(inputs)
torch.cuda.synchronize()
assert torch.isfinite(ep["last_sem_cls_scores"]).all()
lib = _lib.load()
B, N, m = 2, 9000, 300
xyz = torch.from_numpy(synth.synth_scene(3, N, 8)["point_clouds"][:, :3].copy())[None].repeat(B, 1, 1).contiguous().cuda()
ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8, device="cuda")
idx = torch.zeros(B, m, dtype=torch.int32, device="cuda")
_lib.call("bd_grid_build", xyz.data_ptr(), 3, B, N, 0.2, ws.data_ptr())
_lib.call("bd_fps_grid", xyz.data_ptr(), 3, B, N, m, ws.data_ptr(), scratch.data_ptr(), idx.data_ptr())
cen = torch.empty(B, m, 3, device="cuda")
_lib.call("bd_gather_rows", xyz.data_ptr(), 3, idx.data_ptr(), B, N, m, 3, cen.data_ptr(), 3)
out = torch.zeros(B, m, 64, dtype=torch.int32, device="cuda")
_lib.call("bd_ball_query_grid_query", cen.data_ptr(), xyz.data_ptr(), 3, B, N, m, 0.2, 64, out.data_ptr(), ws.data_ptr())
big = torch.from_numpy(synth.synth_scene(4, 50000, 8)["point_clouds"][:, :3].copy())[None].contiguous().cuda()
i2 = torch.zeros(1, 256, dtype=torch.int32, device="cuda")
_lib.call("bd_fps", big.data_ptr(), 3, 1, 50000, 256, None, i2.data_ptr())  # 16-CTA cluster kernel (DSMEM st.async)
torch.cuda.synchronize()
print("sanitize target finished:", precision, "graph" if graph else "eager", int(idx.sum()), int(out.sum()), int(i2.sum()))
# round-2b additions: the text side (RoBERTa engine incl. the head_dim-64 direct attention) and the device matcher
if "--no-extras" not in sys.argv:
    from transformers import RobertaConfig, RobertaModel
    from butd_detr_b200 import text_encoder
    from butd_detr_b200.matcher import HungarianMatcher
    torch.manual_seed(0)
    cfg = RobertaConfig(vocab_size=1000, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1,
                        num_hidden_layers=2)
    eng = text_encoder.from_module(RobertaModel(cfg).cuda(), precision)
    ids = torch.randint(3, 1000, (3, 21)).cuda()
    mask = torch.ones(3, 21, dtype=torch.long).cuda()
    mask[1, 15:] = 0
    ids[1, 15:] = 1
    hid = eng.forward(ids, mask)
    g = torch.Generator().manual_seed(1)
    outs = {"pred_logits": torch.randn(2, 32, 256, generator=g).cuda(),
            "pred_boxes": torch.cat([torch.rand(2, 32, 3, generator=g), torch.rand(2, 32, 3, generator=g) + 0.1], -1).cuda()}
    tg = [{"boxes": torch.cat([torch.rand(n, 3, generator=g), torch.rand(n, 3, generator=g) + 0.1], -1).cuda(),
           "positive_map": torch.rand(n, 256, generator=g).cuda(), "labels": torch.zeros(n, dtype=torch.int64).cuda()} for n in (5, 32)]
    pairs = HungarianMatcher(1, 0, 2, True)(outs, tg)
    # the persistent linear kernels only run with more row tiles than SMs: 157 tiles here
    import math
    from butd_detr_b200.engine import lin_tiling, pack_weight_tc
    M, N, K = 20000, 288, 288
    A = torch.randn(M, K, device="cuda").half()
    W = torch.randn(N, K, device="cuda") / math.sqrt(K)
    bias = torch.randn(N, device="cuda")
    wide, bn = lin_tiling(M, N)
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, wide=wide, bn=bn)
    Y16 = torch.empty(M, N, device="cuda", dtype=torch.float16)
    _lib.call("bd_linear_tc_h", A.data_ptr(), K, 1, None, 0, Wp.data_ptr(), bias.data_ptr(), Y16.data_ptr(), N, 1, M, N, K, KC, nch, BN,
              nsub, 1)
    Wr, (BN, KC, nch, nsub) = pack_weight_tc(W, 1, full_rows=True)
    R, Y = torch.randn(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    gam, bet = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    _lib.call("bd_linear_ln_tc_h", A.data_ptr(), K, 1, Wr.data_ptr(), bias.data_ptr(), R.data_ptr(), N, gam.data_ptr(), bet.data_ptr(),
              1e-5, Y.data_ptr(), N, Y16.data_ptr(), N, M, N, K, KC, nch, BN, nsub)
    torch.cuda.synchronize()
    print("extras finished:", bool(torch.isfinite(hid).all()), [int(p[0].sum()) for p in pairs], bool(torch.isfinite(Y).all()))
