"""Text side of the hot path: the frozen RoBERTa-base encoder the reference calls as
`self.text_encoder(**tokenized).last_hidden_state` (`/root/reference/models/bdetr.py:72-77,164-169`; transformers'
`RobertaModel`: embeddings -> 12 x [self-attention, out-projection + residual + LayerNorm, GELU feed-forward +
residual + LayerNorm]) as a schedule of libbutd_b200 launches.

  * embeddings: one kernel (`bd_roberta_embed`: word + position + token-type rows, position ids from the running
    count of non-pad tokens, LayerNorm);
  * every dense layer: the tensor-core linear kernel of the visual path (`bd_linear_tc` / `bd_linear_tc_h`:
    Q, K, V as ONE 768 -> 2304 GEMM, the exact erf-GELU in the epilogue of the 768 -> 3072 layer);
  * attention, 12 heads of 64: fp16 mode — the tcgen05 attention kernel reading the K / V tiles of a head from
    the QKV projection's fp16 rows by tensor copies (`bd_attention_tc_h`, head_dim 64); bf16x3 / fp32 modes — the
    fp32 flash-style kernel (`bd_attention_f32`, head_dim 64);
  * residual + LayerNorm over 768 columns: `bd_add_layernorm_f32`.

The tokenizer (host string processing, `RobertaTokenizerFast`) stays transformers': `input_ids` and
`attention_mask` are this module's inputs.  PyTorch is used for device memory only; there is no CPU path.
"""
import math

import torch

from . import _lib
from .engine import ForwardEngine


class RobertaEngine(ForwardEngine):
    """Eval forward of a transformers `RobertaModel` (any width with head_dim 64 or 36) on the sm_100a kernels."""

    def __init__(self, state_dict, config, device, precision="fp16"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CPU not supported: the text encoder engine needs a CUDA device")
        if precision not in ("fp32", "fp16", "bf16x3"):
            raise ValueError("precision must be 'fp32', 'fp16' or 'bf16x3'")
        _lib.load()
        self.precision = precision
        self.split = 3 if precision == "bf16x3" else 1
        self.half = precision == "fp16"
        self.d_model = int(config.hidden_size)
        self.n_heads = int(config.num_attention_heads)
        self.n_layers = int(config.num_hidden_layers)
        self.pad_idx = int(config.pad_token_id)
        self.eps = float(config.layer_norm_eps)
        if getattr(config, "hidden_act", "gelu") != "gelu":
            raise ValueError("the text encoder engine implements the exact (erf) GELU only")
        if getattr(config, "position_embedding_type", "absolute") != "absolute":
            raise ValueError("absolute position embeddings only")
        hd = self.d_model // self.n_heads
        if hd != 64:
            raise ValueError("head_dim must be 64 (RoBERTa-base / -large)")
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()}
        w = {"word": sd["embeddings.word_embeddings.weight"], "pos": sd["embeddings.position_embeddings.weight"],
             "type": sd["embeddings.token_type_embeddings.weight"],
             "emb.ln": (sd["embeddings.LayerNorm.weight"], sd["embeddings.LayerNorm.bias"])}
        for i in range(self.n_layers):
            p, k = f"encoder.layer.{i}.", f"l{i}"
            a = p + "attention.self."
            w[k + ".qkv"] = (torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]]).contiguous(),
                             torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]]).contiguous())
            w[k + ".o"] = (sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
            w[k + ".ln1"] = (sd[p + "attention.output.LayerNorm.weight"], sd[p + "attention.output.LayerNorm.bias"])
            w[k + ".ff1"] = (sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"])
            w[k + ".ff2"] = (sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
            w[k + ".ln2"] = (sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])
        self.W = w
        self._tc = {}
        self._shadow = {}
        self._live = []

    def attention(self, qkv, B, L, mask_u8):
        """softmax(Q K^T / 8 + key padding mask) V over the fused (B*L, 3E) projection -> (B*L, E)."""
        E, H = self.d_model, self.n_heads
        hd = E // H
        q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
        ld = qkv.stride(0)
        if self.half and qkv.dtype == torch.float16:
            out = self._empty(B * L, E, dtype=torch.float16)
            _lib.call("bd_attention_tc_h", q.data_ptr(), ld, L * ld, k.data_ptr(), ld, L * ld, v.data_ptr(), ld, L * ld,
                      _lib.ptr(mask_u8), out.data_ptr(), E, L * E, 15, B, H, L, L, hd, 1.0 / math.sqrt(hd), 1, None)
            return out
        out = self._empty(B * L, E)
        _lib.call("bd_attention_f32", q.data_ptr(), ld, L * ld, k.data_ptr(), ld, L * ld, v.data_ptr(), ld, L * ld,
                  _lib.ptr(mask_u8), out.data_ptr(), E, L * E, B, H, L, L, hd, 1.0 / math.sqrt(hd))
        return out

    @torch.no_grad()
    def forward(self, input_ids, attention_mask):
        """input_ids (B,L) int64, attention_mask (B,L) {0,1} (1 = token) -> last_hidden_state (B,L,E) fp32."""
        _lib.check_cuda(input_ids)
        B, L = input_ids.shape
        E = self.d_model
        self._live, self._shadow = [], {}
        with torch.cuda.device(self.device):
            ids = input_ids.to(torch.int64).contiguous()
            mask_u8 = attention_mask.ne(1).to(torch.uint8).contiguous()
            x = self._empty(B * L, E)
            g, b = self.W["emb.ln"]
            _lib.call("bd_roberta_embed", ids.data_ptr(), self.W["word"].data_ptr(), self.W["word"].shape[0],
                      self.W["pos"].data_ptr(), self.W["pos"].shape[0], self.W["type"].data_ptr(), g.data_ptr(),
                      b.data_ptr(), x.data_ptr(), B, L, E, self.pad_idx, self.eps)
            for i in range(self.n_layers):
                k = f"l{i}"
                qkv = self.lin(x, k + ".qkv", half_out=True)
                a = self.attention(qkv, B, L, mask_u8)
                x = self.add_ln(self.lin(a, k + ".o"), x, k + ".ln1", self.eps)
                h = self.lin(x, k + ".ff1", relu=2, half_out=True)  # 2 = exact GELU epilogue
                x = self.add_ln(self.lin(h, k + ".ff2"), x, k + ".ln2", self.eps)
        return x.view(B, L, E)


def from_module(module, precision="fp16", device=None):
    """RobertaEngine over a transformers `RobertaModel` (weights copied to the device once; the module is frozen in
    the reference, `models/bdetr.py:76-77`)."""
    device = device if device is not None else next(module.parameters()).device
    return RobertaEngine(module.state_dict(), module.config, device, precision)
