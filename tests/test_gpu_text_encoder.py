"""Text side: the native RoBERTa forward (butd_detr_b200/text_encoder.py) against the
transformers module the reference calls (`/root/reference/models/bdetr.py:72-77,164-169`), same weights, same
token ids.  Random-initialised RoBERTa-base architecture (no hub access here); ragged lengths, pad tokens."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _roberta(layers=12, seed=0):
    from transformers import RobertaConfig, RobertaModel
    torch.manual_seed(seed)
    cfg = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1, num_hidden_layers=layers)
    m = RobertaModel(cfg).eval()
    # default init is N(0, 0.02) with unit LayerNorms; perturb LayerNorm / biases so that they matter
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("LayerNorm.weight"):
                p.add_(0.2 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            elif "dense.weight" in n or "query.weight" in n or "key.weight" in n or "value.weight" in n:
                p.mul_(2.0)
    return m


def _tokens(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 50265, (B, L), generator=g)
    lens = torch.randint(4, L + 1, (B,), generator=g)
    lens[0] = L
    mask = (torch.arange(L)[None] < lens[:, None]).long()
    ids = torch.where(mask.bool(), ids, torch.ones_like(ids))  # pad token id 1
    ids[:, 0] = 0  # <s>
    return ids, mask


@pytest.mark.parametrize("precision,gate", [("fp16", 1e-2), ("bf16x3", 1e-3), ("fp32", 1e-3)])
@pytest.mark.parametrize("B,L", [(3, 80), (2, 17)])
def test_roberta_engine_matches_transformers(cuda_lib, precision, gate, B, L):
    from butd_detr_b200 import text_encoder
    m = _roberta().cuda()
    ids, mask = _tokens(B, L, 5 + L)
    ids, mask = ids.cuda(), mask.cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            want = m(input_ids=ids, attention_mask=mask).last_hidden_state
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    eng = text_encoder.from_module(m, precision)
    n0 = cuda_lib.launch_count
    got = eng.forward(ids, mask)
    torch.cuda.synchronize()
    assert cuda_lib.launch_count - n0 == 1 + 12 * 7  # embeddings + 7 launches per layer, nothing else
    valid = mask.bool()
    err = (got - want).abs()[valid].max().item()
    assert err <= gate, f"{precision}: max |err| {err} over the valid tokens (gate {gate})"
    assert torch.isfinite(got).all()


def test_model_text_path_native_equals_transformers_module(cuda_lib):
    """`BeaUTyDETR` fed token ids: the native text engine and the transformers module (native_text_encoder=False)
    give the same end_points within the fp32 gate; `tokenized` is passed on for the loss (models/losses.py:574)."""
    from butd_detr_b200 import synth
    from butd_detr_b200.model import BeaUTyDETR
    te = _roberta(layers=2)
    inputs = synth.synth_batch(3, 2, 4096, 16, 32)
    ids, mask = _tokens(2, 16, 9)
    outs = []
    for native in (True, False):
        model = BeaUTyDETR(num_queries=32, num_decoder_layers=1, num_encoder_layers=1, text_encoder=te, precision="fp32",
                           native_text_encoder=native)
        synth.fill_state_dict_({k: v for k, v in model.state_dict().items() if not k.startswith("text_encoder.")}, 0)
        model = model.cuda().eval()
        inp = {k: v.cuda() for k, v in inputs.items() if k not in ("text_hidden",)}
        inp["input_ids"], inp["text_attention_mask"] = ids.cuda(), mask.cuda()
        ep = model(inp)
        assert torch.equal(ep["tokenized"]["attention_mask"], mask.cuda())
        outs.append({k: ep[k].float().clone() for k in ("text_feats", "last_sem_cls_scores", "proj_tokens")})
    for k in outs[0]:
        assert (outs[0][k] - outs[1][k]).abs().max().item() <= 1e-3, k
