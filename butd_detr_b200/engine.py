"""Forward engine: the eval-mode `BeaUTyDETR.forward` (`/root/reference/models/bdetr.py:193-319`)
as a schedule of libbutd_b200 kernel launches on the current CUDA stream.

B200-first choices (vs. the reference's op-by-op PyTorch graph):
  * every activation is TOKEN-MAJOR (one contiguous row per point / token / query), so all
    1x1 convolutions and nn.Linear layers are the same K-major GEMM, neighbour gathers read
    contiguous rows, and the (B,C,N) tensors the reference returns are zero-copy transposed
    views of these buffers;
  * eval-mode BatchNorm is folded into the preceding weights once, at pack time;
  * packed in-projections are applied as fused QK / KV / QKV GEMMs, the positional embedding
    is added inside the GEMM's operand load, and the three prediction-head stems are one GEMM;
  * the FPS chain of levels 2-4 only depends on coordinates and runs on a side stream,
    overlapped with the set-abstraction MLPs;
  * nothing synchronises with the host: the whole forward is CUDA-graph capturable.

PyTorch is used for device memory (torch.empty) and streams only; all arithmetic is in the
C-ABI library.  There is no CPU path.
"""
import contextlib
import math

import torch

from . import _lib

NVTX = True  # stage-level NVTX ranges (backbone / encoder layer i / query generation / decoder layer i) for nsys & ncu


@contextlib.contextmanager
def _range(name):
    """NVTX range around one stage of the forward (host-side annotation; harmless under graph capture)."""
    if NVTX:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if NVTX:
            torch.cuda.nvtx.range_pop()

SA_CFG = (  # models/backbone_module.py:44-78
    ("sa1", 2048, 0.2, 64),
    ("sa2", 1024, 0.4, 32),
    ("sa3", 512, 0.8, 16),
    ("sa4", 256, 1.2, 16),
)
GRID_BALL_QUERY_MIN_POINTS = 2048  # measured at 148 scenes: the cell list also wins for SA2 (2048 points, r = 0.4, nsample = 32)
GRAPH_CACHE_SIZE = 6     # CUDA graphs kept per engine (LRU): each pins the buffers of one input-shape signature
GRAPH_TOKEN_BUCKET = 16  # text length is padded to a multiple of this before a graph is looked up
GRID_RULE_K = 200.0  # cell list when nsample >= K * radius^2 (metres): r=.2 -> 8, r=.4 -> 32, r=.8 -> 128
BN_EPS = 1e-5
LN_EPS = 1e-5


def _round_up(x, m):
    return (x + m - 1) // m * m


def grid_ball_query_rule(n, radius, nsample, m):
    """True when the cell-list ball query (bd_ball_query_grid) should replace the ordered brute-force
    scan (bd_ball_query) — identical results either way.  Shared by the engine and the
    `pointnet2._ext` drop-in.  The scan stops at the nsample-th hit, so it wins for big balls /
    small groups; the cell list tests ~27 cells per centre whatever the cloud size.  Measured on
    50k-point scenes (profiles/, configs[4]): the cell list wins for r = 0.2 at every nsample."""
    return n >= GRID_BALL_QUERY_MIN_POINTS and nsample <= 64 and nsample >= GRID_RULE_K * radius * radius - 1e-6


# ------------------------------------------------------------------------------ weight packing
def _fold_bn(W, bias, sd, bn, k_pad=None):
    """conv/linear (N,K) followed by eval BatchNorm -> (W', b') with the BN folded in."""
    W = W.reshape(W.shape[0], -1).float()
    s = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + BN_EPS)
    b0 = bias.float() if bias is not None else torch.zeros_like(s)
    Wf = W * s[:, None]
    bf = (b0 - sd[bn + ".running_mean"].float()) * s + sd[bn + ".bias"].float()
    if k_pad is not None and k_pad != Wf.shape[1]:
        Wp = Wf.new_zeros(Wf.shape[0], k_pad)
        Wp[:, :Wf.shape[1]] = Wf
        Wf = Wp
    return Wf.contiguous(), bf.contiguous()


def _plain(sd, name):
    W = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    return (W.reshape(W.shape[0], -1).float().contiguous(), None if b is None else b.float().contiguous())


FPS_GRID_MIN_POINTS = 8192  # ... and points per scene
FPS_GRID_MIN_BATCH = 38  # scenes from which SA1's FPS runs as the bucketed one-CTA-per-scene kernel over the cell list
# (bd_fps_grid: a single wave up to 148 scenes, 3.4 - 3.95 ms whatever the batch) instead of the register-resident
# cluster kernel (37 scenes per wave of 2.18 ms: 4.36 ms at 64 scenes, 8.72 ms at 128)
DECODER_SPLIT_MIN = int(__import__("os").environ.get("BUTD_DECODER_SPLIT_MIN", 1 << 30))  # scenes per half above which the decoder would run as two batch halves on two streams:
# measured at 32 scenes: 2336 vs 2400 scenes/s — no gain (the kernels of one half already fill a wave), so it is off
FUSED_SA = True  # set-abstraction levels as one kernel (bd_sa_mlp_tc); False = three GEMM launches
HALF_ACT = True  # fp16 mode: activations that are only ever GEMM operands (FFN / head / objectness hidden layers) live
# in HBM as fp16, and every LayerNorm output gets an fp16 copy that the next projections read by tensor copy straight
# into the operand layout (bd_linear_tc_h / bd_linear_ln_tc_h) — the values are the ones the fp32 path rounds to fp16
# inside the consumer anyway, the bytes are half
DIRECT_KV = True  # fp16 mode: the attention kernel fetches K / V tiles from the projection output by tensor copies (no
# pack kernel, no packed workspace); False = bd_attention_tc's pack kernel (A/B reference)
TC_KC = 64  # k-chunk of the tensor-core kernels = one 128-byte swizzle block of bf16


def tc_tiling(N, K, split=1, full_rows=False, wide=False, bn=None):
    """(BN, KC, n_chunks, n_sub) for bd_linear_tc / bd_linear_ln_tc on an (N, K) weight.
    BN <= 160 columns per accumulator (multiple of 16); a CTA computes n_sub of them (all of
    them when `full_rows`, so that the LayerNorm epilogue sees complete rows); K is consumed in
    chunks of 64 (one 128-byte-swizzle block), zero padded."""
    nt = -(-N // 160)
    BN = _round_up(-(-N // nt), 16)
    if bn is not None:  # caller-chosen accumulator width (latency tiling for small M, see lin_tiling)
        BN, nt = bn, -(-N // bn)
    # `wide`: two accumulators per CTA, so the A tile is staged once for 2 x BN columns (measured
    # 1.3x on 32768-row GEMMs); small-M calls keep one per CTA for more CTAs in flight
    n_sub = nt if full_rows else (min(nt, 2) if wide else 1)
    n_chunks = -(-K // TC_KC)
    return BN, TC_KC, n_chunks, n_sub


def lin_tiling(M, N, n_sm=148):
    """(wide, bn) for bd_linear_tc on an (M, N) output: one accumulator per CTA while the grid
    fits one wave of the 148 SMs (more CTAs in flight: 13.1 vs 16.2 us at M = 8192, N = 288), two
    accumulators per CTA (the A tile staged once for 2 x BN columns) beyond that (66 vs 87 us at
    M = 32768, N = 576)."""
    rows = -(-M // 128)
    nt = -(-N // 160)
    bn = _round_up(-(-N // nt), 16)
    if rows * nt > n_sm:
        # more than a wave anyway: wide tiles, fewer A stagings (re-measured with fp16 A rows at 148 scenes: one
        # accumulator per CTA and two CTAs per SM is 3.5 % slower over the whole forward)
        return (N > 160), None
    # (accumulators narrower than the default were measured too: no gain — a CTA's time is its
    #  operand-load chain, not its MMA / write-out width)
    return False, None


def pack_weight_tc(W, split=1, full_rows=False, wide=False, bn=None):
    """(N, K) fp32 -> bf16 blocks in the tensor-core kernels' shared-memory layout (128-byte
    swizzle, K-major; csrc/tc_common.cuh): Wp[n_group][k_chunk][part][sub][BN rows][64 k] where
    the 16-byte chunk j of row r holds k-chunk (j ^ (r % 8)); zero padded; part = {hi} or, for
    split 3, {hi, lo} with lo = bf16(W - hi)."""
    N, K = W.shape
    BN, KC, n_chunks, n_sub = tc_tiling(N, K, split, full_rows, wide, bn)
    ng = -(-N // (BN * n_sub))
    Wp = W.new_zeros(ng * n_sub * BN, n_chunks * KC)
    Wp[:N, :K] = W
    if split == 1:  # single-pass mode: fp16 operands (csrc/tc_common.cuh); both dtypes are 2 bytes
        # saturate like the kernels' activation conversion does (cvt.rn.satfinite): no inf operands
        parts = [Wp.clamp(-65504.0, 65504.0).to(torch.float16).view(torch.bfloat16)]
    else:
        hi = Wp.to(torch.bfloat16)
        parts = [hi, (Wp - hi.float()).to(torch.bfloat16)]
    P = torch.stack(parts, 0).view(len(parts), ng, n_sub, BN, n_chunks, 8, 8)  # (..., row, kc, chunk, elem)
    r = torch.arange(BN, device=W.device) % 8
    src_chunk = torch.arange(8, device=W.device)[None, :] ^ r[:, None]         # (BN, 8): logical chunk at slot j
    idx = src_chunk.view(1, 1, 1, BN, 1, 8, 1).expand(len(parts), ng, n_sub, BN, n_chunks, 8, 8)
    P = torch.gather(P, 5, idx)
    P = P.permute(1, 4, 0, 2, 3, 5, 6)  # (ng, kc, part, sub, row, chunk, elem)
    return P.contiguous(), (BN, KC, n_chunks, n_sub)


class PackedWeights:
    """Device-resident, BN-folded, GEMM-ready weights (fp32 masters)."""

    def __init__(self, sd, cfg):
        w = {}
        bb = "backbone_net."
        for name, _, _, _ in SA_CFG:
            for i in range(3):
                p = f"{bb}{name}.mlp_module.layer{i}"
                W = sd[p + ".conv.weight"]
                k = W.shape[1]
                w[f"{name}.{i}"] = _fold_bn(W, None, sd, p + ".bn.bn", _round_up(k, 8) if i == 0 else None)
        for name in ("fp1", "fp2"):
            for i in range(2):
                p = f"{bb}{name}.mlp.layer{i}"
                w[f"{name}.{i}"] = _fold_bn(sd[p + ".conv.weight"], None, sd, p + ".bn.bn")
        w["text_projector"] = _plain(sd, "text_projector.0")
        w["text_projector.ln"] = (sd["text_projector.1.weight"].float().contiguous(),
                                  sd["text_projector.1.bias"].float().contiguous())

        def posembed(prefix, key):
            h = prefix + ".position_embedding_head"
            w[key + ".0"] = _fold_bn(sd[h + ".0.weight"], sd[h + ".0.bias"], sd, h + ".1")
            w[key + ".1"] = _plain(sd, h + ".3")

        def mha(prefix, key):
            Wi, bi = sd[prefix + ".in_proj_weight"].float().contiguous(), sd[prefix + ".in_proj_bias"].float().contiguous()
            E = Wi.shape[1]
            w[key + ".q"] = (Wi[:E], bi[:E])
            w[key + ".k"] = (Wi[E:2 * E], bi[E:2 * E])
            w[key + ".v"] = (Wi[2 * E:], bi[2 * E:])
            w[key + ".qk"] = (Wi[:2 * E], bi[:2 * E])
            w[key + ".kv"] = (Wi[E:], bi[E:])
            w[key + ".qkv"] = (Wi, bi)
            w[key + ".o"] = _plain(sd, prefix + ".out_proj")

        def ln(prefix, key):
            w[key] = (sd[prefix + ".weight"].float().contiguous(), sd[prefix + ".bias"].float().contiguous())

        def three_layer_head(prefix, key):
            # center / size / sem heads share their input: fuse the three first layers
            heads = ("center_residual_head", "size_pred_head", "sem_cls_scores_head")
            l0 = [_fold_bn(sd[f"{prefix}.{h}.net.0.weight"], None, sd, f"{prefix}.{h}.net.1") for h in heads]
            w[key + ".stem"] = (torch.cat([a for a, _ in l0]).contiguous(), torch.cat([b for _, b in l0]).contiguous())
            for h, short in zip(heads, ("center", "size", "sem")):
                w[f"{key}.{short}.1"] = _fold_bn(sd[f"{prefix}.{h}.net.4.weight"], None, sd, f"{prefix}.{h}.net.5")
                w[f"{key}.{short}.2"] = _plain(sd, f"{prefix}.{h}.net.8")

        if cfg["butd"]:
            posembed("box_embeddings", "box_embeddings")
            w["class_embeddings"] = _plain(sd, "class_embeddings")
            w["butd_class_embeddings"] = (sd["butd_class_embeddings.weight"].float().contiguous(), None)
        posembed("pos_embed", "pos_embed")
        for i in range(cfg["num_encoder_layers"]):
            p, k = f"cross_encoder.layers.{i}", f"enc{i}"
            if cfg["self_attend"]:
                mha(p + ".self_attention_visual.self_attn", k + ".sv")
                ln(p + ".self_attention_visual.norm1", k + ".sv.ln")
                mha(p + ".self_attention_lang.self_attn", k + ".sl")
                ln(p + ".self_attention_lang.norm1", k + ".sl.ln")
            c = p + ".cross_layer"
            mha(c + ".cross_lv", k + ".lv")
            mha(c + ".cross_vl", k + ".vl")
            for n in ("norm_lv", "norm_lv2", "norm_vl", "norm_vl2"):
                ln(f"{c}.{n}", f"{k}.{n}")
            for f in ("ffn_lv", "ffn_vl"):
                w[f"{k}.{f}.0"] = _plain(sd, f"{c}.{f}.0")
                w[f"{k}.{f}.1"] = _plain(sd, f"{c}.{f}.3")
            if cfg["butd"]:
                mha(c + ".cross_d", k + ".d")
                ln(c + ".norm_d", k + ".norm_d")
        for n in ("conv1", "conv2"):
            w["points_obj_cls." + n] = _fold_bn(sd[f"points_obj_cls.{n}.weight"], sd[f"points_obj_cls.{n}.bias"], sd,
                                                "points_obj_cls.bn" + n[-1])
        w["points_obj_cls.conv3"] = _plain(sd, "points_obj_cls.conv3")
        w["decoder_query_proj"] = _plain(sd, "decoder_query_proj")
        three_layer_head("proposal_head", "proposal_head")
        for i in range(cfg["num_decoder_layers"]):
            p, k = f"decoder.{i}", f"dec{i}"
            mha(p + ".self_attn", k + ".self")
            mha(p + ".cross_l", k + ".l")
            mha(p + ".cross_v", k + ".v")
            for n in ("norm1", "norm_l", "norm_v", "norm2"):
                ln(f"{p}.{n}", f"{k}.{n}")
            if cfg["butd"]:
                mha(p + ".cross_d", k + ".d")
                ln(p + ".norm_d", k + ".norm_d")
            w[k + ".ffn.0"] = _plain(sd, p + ".ffn.0")
            w[k + ".ffn.1"] = _plain(sd, p + ".ffn.3")
            if cfg["self_position_embedding"] in ("loc_learned", "xyz_learned"):
                posembed(p + ".self_posembed", k + ".posembed")
            three_layer_head(f"prediction_heads.{i}", f"head{i}")
        if cfg["contrastive_align_loss"]:
            for side in ("image", "text"):
                for j, src in enumerate((0, 2, 4)):
                    w[f"contrastive.{side}.{j}"] = _plain(sd, f"contrastive_align_projection_{side}.{src}")
        self.w = w

    def __getitem__(self, k):
        return self.w[k]


# ------------------------------------------------------------------------------ engine
class ForwardEngine:
    def __init__(self, state_dict, cfg, device):
        self.cfg = dict(cfg)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CPU not supported: the BUTD-DETR B200 engine needs a CUDA device")
        _lib.load()
        sd = {k: v.detach().to(self.device) for k, v in state_dict.items() if not k.startswith("text_encoder.")}
        self.W = PackedWeights(sd, self.cfg)
        self.precision = cfg.get("precision", "fp32")
        if self.precision not in ("fp32", "fp16", "bf16x3"):
            raise ValueError("precision must be 'fp32' (SIMT), 'fp16' (tcgen05, fp16 operands, one MMA per product) "
                             "or 'bf16x3' (tcgen05, bf16 hi/lo split operands, three MMAs per product)")
        self.split = 3 if self.precision == "bf16x3" else 1
        self.half = self.precision == "fp16" and HALF_ACT
        self._shadow = {}  # data_ptr of an fp32 LayerNorm output -> its fp16 copy (this forward)
        self._tc = {}  # key -> (packed bf16 weight, (BN, KC, n_chunks)), built on first use
        self.d_model = cfg["d_model"]
        self.n_heads = 8
        self.side_stream = torch.cuda.Stream(device=self.device)
        # Independent branches of the transformer run on their own streams (inside the captured
        # graph as well): most of its kernels cover 20-64 CTAs of the 148 SMs, so concurrency is
        # nearly free.  text_stream: language half of the encoder; kv_stream: the decoder's memory
        # K/V projections (they depend on the encoder output only); head_stream: class scores and
        # contrastive projections of each decoder layer (nothing downstream waits for them).
        self.text_stream = torch.cuda.Stream(device=self.device)
        self.kv_stream = torch.cuda.Stream(device=self.device)
        self.head_stream = torch.cuda.Stream(device=self.device)
        self.dec_stream = torch.cuda.Stream(device=self.device)
        self.aux_stream = torch.cuda.Stream(device=self.device)  # short independent branches (V projection, size head)
        self._live = []  # every buffer of the current forward: nothing is recycled while streams overlap

    # ---- thin wrappers over the C-ABI (2-D row-major views; last stride must be 1)
    def _empty(self, *shape, dtype=torch.float32):
        t = torch.empty(*shape, dtype=dtype, device=self.device)
        self._live.append(t)
        return t

    def lin(self, x, key, relu=False, add=None, out=None, half_out=False):
        W, b = self.W[key]
        M, K = x.shape
        N = W.shape[0]
        assert x.stride(1) == 1 and W.shape[1] == K, (key, x.shape, W.shape)
        # 16-bit activations (fp16 mode): an fp16 input, or the fp16 copy of an fp32 LayerNorm output
        a_half = x.dtype == torch.float16
        if self.half and add is None and not a_half and x.is_contiguous():
            sh = self._shadow.get(x.data_ptr())
            if sh is not None and sh.shape == x.shape:
                x, a_half = sh, True
        assert not (a_half and add is not None), key
        y_half = bool(half_out and self.half and out is None)
        if (a_half or y_half) and K % 8 == 0 and x.stride(0) % (8 if a_half else 4) == 0 and x.data_ptr() % 16 == 0 and \
                (add is None or (add.stride(0) % 4 == 0 and add.data_ptr() % 16 == 0)):
            wide, bn = lin_tiling(M, N)
            tkey = f"{key}#{'wide' if wide else bn}"
            if tkey not in self._tc:
                self._tc[tkey] = pack_weight_tc(W, self.split, wide=wide, bn=bn)
            Wp, (BN, KC, n_chunks, n_sub) = self._tc[tkey]
            if out is None:
                out = self._empty(M, _round_up(N, 8), dtype=torch.float16)[:, :N] if y_half else self._empty(M, N)
            _lib.call("bd_linear_tc_h", x.data_ptr(), x.stride(0), int(a_half), _lib.ptr(add),
                      0 if add is None else add.stride(0), Wp.data_ptr(), _lib.ptr(b), out.data_ptr(),
                      out.stride(0), int(y_half), M, N, K, KC, n_chunks, BN, n_sub, int(relu))
            return out
        assert not a_half, key
        if K <= 8 and add is None and N % 2 == 0 and (out is None or (out.stride(0) % 2 == 0 and out.data_ptr() % 8 == 0)):
            # narrow-input layer (position embeddings): a write stream; fp16 rows when the next layer reads them
            y_half = bool(half_out and self.half and out is None)
            if out is None:
                out = self._empty(M, N, dtype=torch.float16 if y_half else torch.float32)
            _lib.call("bd_linear_smallk", x.data_ptr(), x.stride(0), W.data_ptr(), _lib.ptr(b), out.data_ptr(), out.stride(0),
                      int(y_half), M, N, K, int(bool(relu)))
            return out
        if out is None:
            out = self._empty(M, N)
        assert out.stride(1) == 1
        lda2 = 0 if add is None else add.stride(0)
        # the choice of kernel must not depend on the batch size (a batch has to equal its scenes run
        # alone, bit for bit): only K and alignment decide
        tc_ok = (K % 8 == 0 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and
                 (add is None or (add.stride(0) % 4 == 0 and add.data_ptr() % 16 == 0)))
        if self.precision != "fp32" and tc_ok:  # tensor cores (narrow heads padded to 16 columns); K = 3 / 6 inputs stay fp32
            wide, bn = lin_tiling(M, N)
            if add is not None and self.split == 3:
                wide = False  # two raw fp32 chunks (x and pos) + hi/lo weights of a 288-wide tile: no room for 2 stages
            tkey = f"{key}#{'wide' if wide else bn}"
            if tkey not in self._tc:
                self._tc[tkey] = pack_weight_tc(W, self.split, wide=wide, bn=bn)
            Wp, (BN, KC, n_chunks, n_sub) = self._tc[tkey]
            _lib.call("bd_linear_tc", x.data_ptr(), x.stride(0), _lib.ptr(add), lda2, Wp.data_ptr(), _lib.ptr(b),
                      out.data_ptr(), out.stride(0), M, N, K, KC, n_chunks, BN, n_sub, int(relu), self.split)
        else:
            _lib.call("bd_linear_f32", x.data_ptr(), x.stride(0), _lib.ptr(add), lda2, W.data_ptr(), _lib.ptr(b),
                      out.data_ptr(), out.stride(0), M, N, K, int(relu))
        return out

    def lin_ln(self, x, key, res, ln_key, add=None, eps=LN_EPS, shadow=True):
        """LayerNorm(res + linear(x [+ add])) — one kernel on the tensor-core path.  fp16 mode: also writes
        the fp16 copy of the result (`shadow`) that the next projections read as their operand."""
        W, b = self.W[key]
        M, K = x.shape
        N = W.shape[0]
        if self.half and add is None and N <= 320 and N % 4 == 0 and K % 8 == 0 and x.data_ptr() % 16 == 0 and \
                x.stride(0) % (8 if x.dtype == torch.float16 else 4) == 0:
            assert x.stride(1) == 1 and W.shape[1] == K and res.is_contiguous() and res.shape == (M, N)
            tkey = key + "#rows"
            if tkey not in self._tc:
                self._tc[tkey] = pack_weight_tc(W, self.split, full_rows=True)
            Wp, (BN, KC, n_chunks, n_sub) = self._tc[tkey]
            g, beta = self.W[ln_key]
            out = self._empty(M, N)
            sh = self._empty(M, N, dtype=torch.float16) if shadow else None
            _lib.call("bd_linear_ln_tc_h", x.data_ptr(), x.stride(0), int(x.dtype == torch.float16), Wp.data_ptr(),
                      _lib.ptr(b), res.data_ptr(), N, g.data_ptr(), beta.data_ptr(), float(eps), out.data_ptr(), N,
                      _lib.ptr(sh), N, M, N, K, KC, n_chunks, BN, n_sub)
            if sh is not None:
                self._shadow[out.data_ptr()] = sh
            return out
        assert x.dtype == torch.float32, key
        if (self.precision == "fp32" or N > 320 or N % 4 or K % 8 or x.stride(0) % 4 or x.data_ptr() % 16 or
                (add is not None and (add.stride(0) % 4 or add.data_ptr() % 16))):
            return self.add_ln(self.lin(x, key, add=add), res, ln_key, eps)
        assert x.stride(1) == 1 and W.shape[1] == K and res.is_contiguous() and res.shape == (M, N)
        tkey = key + "#rows"
        if tkey not in self._tc:
            self._tc[tkey] = pack_weight_tc(W, self.split, full_rows=True)
        Wp, (BN, KC, n_chunks, n_sub) = self._tc[tkey]
        g, beta = self.W[ln_key]
        out = self._empty(M, N)
        _lib.call("bd_linear_ln_tc", x.data_ptr(), x.stride(0), _lib.ptr(add), 0 if add is None else add.stride(0),
                  Wp.data_ptr(), _lib.ptr(b), res.data_ptr(), N, g.data_ptr(), beta.data_ptr(), float(eps),
                  out.data_ptr(), N, M, N, K, KC, n_chunks, BN, n_sub, self.split)
        return out

    def add_ln(self, x, res, key, eps=LN_EPS, out=None):
        g, b = self.W[key]
        M, D = x.shape
        assert x.is_contiguous() and (res is None or res.is_contiguous())
        if out is None:
            out = self._empty(M, D)
        _lib.call("bd_add_layernorm_f32", x.data_ptr(), _lib.ptr(res), g.data_ptr(), b.data_ptr(), out.data_ptr(),
                  M, D, float(eps))
        return out

    def attention(self, q, k, v, B, Lq, Lk, mask):
        """q (B*Lq, >=E) / k, v (B*Lk, >=E) 2-D views (row stride = leading dim) -> (B*Lq, E)."""
        E, H = self.d_model, self.n_heads
        hd = E // H
        aligned = (hd == 36 and q.stride(0) % 4 == 0 and k.stride(0) % 4 == 0 and q.data_ptr() % 16 == 0
                   and k.data_ptr() % 16 == 0)
        if self.half and aligned and v.stride(0) % 4 == 0 and v.data_ptr() % 8 == 0:
            # 16-bit tensors in HBM: fp16 q / k / v where the projections wrote them so, fp16 output (the
            # out-projection's operand)
            out = self._empty(B * Lq, E, dtype=torch.float16)
            io = (int(q.dtype == torch.float16) | int(k.dtype == torch.float16) << 1 | int(v.dtype == torch.float16) << 2 | 8)
            direct = (DIRECT_KV and k.dtype == v.dtype == torch.float16 and k.stride(0) % 8 == 0 and v.stride(0) % 8 == 0
                      and v.data_ptr() % 16 == 0)
            ws = None if direct else self._empty(_lib.load().bd_attention_tc_workspace_bytes(B, H, Lq, Lk, self.split),
                                                 dtype=torch.uint8)
            _lib.call("bd_attention_tc_h", q.data_ptr(), q.stride(0), Lq * q.stride(0), k.data_ptr(), k.stride(0),
                      Lk * k.stride(0), v.data_ptr(), v.stride(0), Lk * v.stride(0), _lib.ptr(mask), out.data_ptr(), E, Lq * E,
                      io, B, H, Lq, Lk, hd, 1.0 / math.sqrt(hd), self.split, _lib.ptr(ws))
            return out
        assert q.dtype == k.dtype == v.dtype == torch.float32
        out = self._empty(B * Lq, E)
        args = (q.data_ptr(), q.stride(0), Lq * q.stride(0), k.data_ptr(), k.stride(0), Lk * k.stride(0),
                v.data_ptr(), v.stride(0), Lk * v.stride(0), _lib.ptr(mask), out.data_ptr(), E, Lq * E, B, H, Lq, Lk,
                hd, 1.0 / math.sqrt(hd))
        if self.precision != "fp32" and aligned:
            ws = self._empty(_lib.load().bd_attention_tc_workspace_bytes(B, H, Lq, Lk, self.split), dtype=torch.uint8)
            _lib.call("bd_attention_tc", *args, self.split, ws.data_ptr())
        else:
            _lib.call("bd_attention_f32", *args)
        return out

    def pack_kv(self, kv, B, Lq, Lk):
        """K / V^T operand tiles of a fused (B*Lk, 2E) key-value projection, packed ahead of time (the
        decoder's memory K / V: on kv_stream, off the critical path).  None when the tensor-core
        attention does not apply."""
        E, H = self.d_model, self.n_heads
        hd = E // H
        k, v = kv[:, :E], kv[:, E:]
        if self.precision == "fp32" or hd != 36 or k.stride(0) % 4 or k.data_ptr() % 16 or v.data_ptr() % 16:
            return None
        if self.half and DIRECT_KV and kv.dtype == torch.float16 and k.stride(0) % 8 == 0:
            return None  # nothing to pack: the attention kernel reads these rows by tensor copies
        ws = self._empty(_lib.load().bd_attention_tc_workspace_bytes(B, H, Lq, Lk, self.split), dtype=torch.uint8)
        _lib.call("bd_attention_tc_pack_kv", k.data_ptr(), k.stride(0), Lk * k.stride(0), v.data_ptr(), v.stride(0),
                  Lk * v.stride(0), int(kv.dtype == torch.float16), B, H, Lq, Lk, hd, self.split, ws.data_ptr())
        return ws

    def attention_packed(self, q, ws, B, Lq, Lk, mask):
        """attention() over K / V tiles packed by pack_kv (same B, Lq, Lk)."""
        E, H = self.d_model, self.n_heads
        hd = E // H
        out = self._empty(B * Lq, E, dtype=torch.float16 if self.half else torch.float32)
        io = int(q.dtype == torch.float16) | (8 if self.half else 0)
        _lib.call("bd_attention_tc_packed", q.data_ptr(), q.stride(0), Lq * q.stride(0), _lib.ptr(mask), out.data_ptr(), E,
                  Lq * E, io, B, H, Lq, Lk, hd, 1.0 / math.sqrt(hd), self.split, ws.data_ptr())
        return out

    def mha(self, key, x_q, pos_q, x_kv, pos_k, B, Lq, Lk, mask, self_attn=False, res=None, ln_key=None, kv=None,
            packed=None):
        """nn.MultiheadAttention (eval) incl. in/out projections.  q = x_q (+pos_q);
        k = x_kv (+pos_k); v = x_kv.  Returns LayerNorm(res + out_proj(attention)) — the
        post-LN residual block every call site of the reference wraps around the attention."""
        E = self.d_model
        if self_attn and pos_q is None:  # q = k = v = x : one fused QKV GEMM
            qkv = self.lin(x_q, key + ".qkv", half_out=True)
            q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
        elif self_attn:  # q = k = x + pos, v = x : the V projection runs beside the Q/K one
            cur = torch.cuda.current_stream()
            aux = self.aux_stream
            aux.wait_event(cur.record_event())
            with torch.cuda.stream(aux):
                v = self.lin(x_q, key + ".v", half_out=True)
                v_ready = aux.record_event()
            qk = self.lin(x_q, key + ".qk", add=pos_q, half_out=True)
            q, k = qk[:, :E], qk[:, E:]
            cur.wait_event(v_ready)
        else:  # cross attention: k = v = memory (no positional term in this model)
            q = self.lin(x_q, key + ".q", add=pos_q, half_out=True)
            assert pos_k is None
            if packed is not None and q.stride(0) % 4 == 0 and q.data_ptr() % 16 == 0:
                # memory K / V projected AND packed earlier on kv_stream
                return self.lin_ln(self.attention_packed(q, packed, B, Lq, Lk, mask), key + ".o", res, ln_key)
            if kv is None:  # else: projected earlier on kv_stream
                kv = self.lin(x_kv, key + ".kv", half_out=True)
            k, v = kv[:, :E], kv[:, E:]
        o = self.attention(q, k, v, B, Lq, Lk, mask)
        return self.lin_ln(o, key + ".o", res, ln_key)

    def ffn(self, x, key, ln_key):
        """LayerNorm(x + W2 relu(W1 x))  (encoder_decoder_layers.py:52-58,96,122)."""
        return self.lin_ln(self.lin(x, key + ".0", relu=True, half_out=True), key + ".1", x, ln_key)

    def posembed(self, x, key):
        return self.lin(self.lin(x, key + ".0", relu=True, half_out=True), key + ".1")

    def head(self, feats, base_xyz, key, out):
        """ClsAgnosticPredictHead.forward (models/modules.py:135-180) on token-major rows (n, E).
        `out`: dict of row views {center (n,3), size (n,3), sem (n,C)} of the full-batch outputs.
        Class scores run on head_stream: nothing downstream on the calling stream needs them."""
        E = self.d_model
        stem = self.lin(feats, key + ".stem", relu=True, half_out=True)  # (n, 3E)
        cur = torch.cuda.current_stream()
        self.head_stream.wait_event(cur.record_event())
        with torch.cuda.stream(self.head_stream):
            h = self.lin(stem[:, 2 * E:3 * E], f"{key}.sem.1", relu=True, half_out=True)
            self.lin(h, f"{key}.sem.2", out=out["sem"])
        aux = self.aux_stream
        aux.wait_event(cur.record_event())
        with torch.cuda.stream(aux):  # the size branch beside the centre branch
            h2 = self.lin(stem[:, E:2 * E], f"{key}.size.1", relu=True, half_out=True)
            self.lin(h2, f"{key}.size.2", out=out["size"])
            size_ready = aux.record_event()
        h = self.lin(stem[:, :E], f"{key}.center.1", relu=True, half_out=True)
        delta = self.lin(h, f"{key}.center.2")
        n = feats.shape[0]
        _lib.call("bd_add_rows", base_xyz.data_ptr(), 3, delta.data_ptr(), 3, out["center"].data_ptr(), 3, n, 3)
        cur.wait_event(size_ready)
        return out["center"], out["size"]

    def contrastive(self, x, side, B, L, out=None):
        h = self.lin(x, f"contrastive.{side}.0", relu=True, half_out=True)
        h = self.lin(h, f"contrastive.{side}.1", relu=True, half_out=True)
        h = self.lin(h, f"contrastive.{side}.2", out=out)
        _lib.call("bd_l2_normalize_rows", h.data_ptr(), h.data_ptr(), h.shape[0], h.shape[1])
        return h.view(B, L, -1)

    # ---- point ops
    def fps(self, xyz, ld, B, N, m, out=None):
        idx = out if out is not None else self._empty(B, m, dtype=torch.int32)
        tmp = self._empty(B, N) if N > _lib.load().bd_fps_resident_capacity() else None
        _lib.call("bd_fps", xyz.data_ptr(), ld, B, N, m, _lib.ptr(tmp), idx.data_ptr())
        return idx

    def gather_rows(self, src, ld_src, idx, B, n_src, m, w, out=None):
        out = out if out is not None else self._empty(B, m, w)
        _lib.call("bd_gather_rows", src.data_ptr(), ld_src, idx.data_ptr(), B, n_src, m, w, out.data_ptr(), w)
        return out

    def sa_level(self, name, xyz, ld_xyz, feats, ld_feats, C, new_xyz, B, n, m, radius, ns):
        """QueryAndGroup + SharedMLP + max-pool (pointnet2_modules.py:243-257), token-major."""
        idx = self._empty(B, m, ns, dtype=torch.int32)
        grid = getattr(self, "_grid", None)
        if grid is not None and grid[1] == n and grid[2] == float(radius):  # cell list already built (backbone)
            _lib.call("bd_ball_query_grid_query", new_xyz.data_ptr(), xyz.data_ptr(), ld_xyz, B, n, m, float(radius), ns,
                      idx.data_ptr(), grid[0].data_ptr())
            self._grid = None
        elif grid_ball_query_rule(n, radius, ns, m):  # cell-list search (identical output, ~100x fewer distance tests)
            ws = self._empty(_lib.load().bd_ball_query_grid_workspace_bytes(B, n), dtype=torch.uint8)
            _lib.call("bd_ball_query_grid", new_xyz.data_ptr(), xyz.data_ptr(), ld_xyz, B, n, m, float(radius), ns,
                      idx.data_ptr(), ws.data_ptr())
        else:
            _lib.call("bd_ball_query", new_xyz.data_ptr(), xyz.data_ptr(), ld_xyz, B, n, m, float(radius), ns,
                      idx.data_ptr())
        if self.precision != "fp32":
            return self._sa_level_tc(name, idx, xyz, ld_xyz, feats, ld_feats, C, new_xyz, B, n, m, radius, ns)
        kp = self.W[name + ".0"][0].shape[1]
        g = self._empty(B * m * ns, kp)
        _lib.call("bd_group_rows", xyz.data_ptr(), ld_xyz, feats.data_ptr(), ld_feats, C, new_xyz.data_ptr(),
                  idx.data_ptr(), B, n, m, ns, float(radius), g.data_ptr(), kp)
        h = self.lin(g, name + ".0", relu=True)
        h = self.lin(h, name + ".1", relu=True)
        h = self.lin(h, name + ".2", relu=True)
        cout = h.shape[1]
        out = self._empty(B, m, cout)
        _lib.call("bd_maxpool_rows", h.data_ptr(), B * m, ns, cout, out.data_ptr())
        return out

    def _sa_level_fused(self, name, idx, xyz, ld_xyz, feats, ld_feats, C, new_xyz, B, n, m, radius, ns):
        """Whole SA level in one kernel (bd_sa_mlp_tc); None when the level's shapes do not fit it."""
        Ws = [self.W[f"{name}.{i}"] for i in range(3)]
        N = [w.shape[0] for w, _ in Ws]
        if not (FUSED_SA and 128 % ns == 0 and N[0] in (64, 128) and N[1] in (64, 128) and N[2] in (64, 128, 256)
                and Ws[1][0].shape[1] == N[0] and Ws[2][0].shape[1] == N[1]):
            return None
        key = name + "#fused"
        if key not in self._tc:  # layer 0: reorder K [xyz(3) | feats(C)] -> [feats(C) | xyz(3)], pad to 8
            W0 = Ws[0][0]
            Wr = torch.cat([W0[:, 3:3 + C], W0[:, :3]], 1)
            Wr = torch.nn.functional.pad(Wr, (0, _round_up(C + 3, 8) - (C + 3)))
            self._tc[key] = [pack_weight_tc(w, self.split, full_rows=True)[0] for w in (Wr, Ws[1][0], Ws[2][0])]
        Wp = self._tc[key]
        out = self._empty(B, m, N[2])
        if self.half:
            # 16-bit feature rows between the levels: gather the previous level's fp16 copy (half the bytes, no
            # conversion) and leave an fp16 copy of this level's pooled rows for the next one
            f16 = self._shadow.get(feats.data_ptr()) if torch.is_tensor(feats) else None
            use16 = f16 is not None and C % 8 == 0 and f16.shape[-1] == C
            out16 = self._empty(B, m, N[2], dtype=torch.float16) if name != SA_CFG[-1][0] else None
            src = f16 if use16 else feats
            _lib.call("bd_sa_mlp_tc_h", idx.data_ptr(), src.data_ptr(), C if use16 else ld_feats, C, int(use16),
                      xyz.data_ptr(), ld_xyz, new_xyz.data_ptr(), B, n, m, ns, float(radius), Wp[0].data_ptr(),
                      Ws[0][1].data_ptr(), N[0], Wp[1].data_ptr(), Ws[1][1].data_ptr(), N[1], Wp[2].data_ptr(),
                      Ws[2][1].data_ptr(), N[2], out.data_ptr(), N[2], _lib.ptr(out16), N[2], self.split)
            if out16 is not None:
                self._shadow[out.data_ptr()] = out16
            return out
        _lib.call("bd_sa_mlp_tc", idx.data_ptr(), feats.data_ptr(), ld_feats, C, xyz.data_ptr(), ld_xyz,
                  new_xyz.data_ptr(), B, n, m, ns, float(radius), Wp[0].data_ptr(), Ws[0][1].data_ptr(), N[0],
                  Wp[1].data_ptr(), Ws[1][1].data_ptr(), N[1], Wp[2].data_ptr(), Ws[2][1].data_ptr(), N[2],
                  out.data_ptr(), N[2], self.split)
        return out

    def _sa_level_tc(self, name, idx, xyz, ld_xyz, feats, ld_feats, C, new_xyz, B, n, m, radius, ns):
        """Tensor-core SA level: one fused kernel, or (FUSED_SA = False) [gather fused into layer 0]
        -> layer 1 -> [layer 2 + max-pool fused]."""
        out = self._sa_level_fused(name, idx, xyz, ld_xyz, feats, ld_feats, C, new_xyz, B, n, m, radius, ns)
        if out is not None:
            return out
        if C % 8 == 0 and ld_feats % 4 == 0:
            # wide feature rows (SA2-4): gather fused into the GEMM's operand staging
            k0 = name + ".0#gather"
            if k0 not in self._tc:  # reorder K: [xyz(3) | feats(C)] -> [feats(C) | xyz(3)], pad to 8
                W, _ = self.W[name + ".0"]
                Wr = torch.cat([W[:, 3:3 + C], W[:, :3]], 1)
                self._tc[k0] = pack_weight_tc(torch.nn.functional.pad(Wr, (0, _round_up(C + 3, 8) - (C + 3))), self.split)
            Wp, (BN, KC, n_chunks, n_sub) = self._tc[k0]
            b0 = self.W[name + ".0"][1]
            N0 = b0.shape[0]
            h = self._empty(B * m * ns, N0)
            _lib.call("bd_sa_group_linear_tc", idx.data_ptr(), feats.data_ptr(), ld_feats, C, xyz.data_ptr(), ld_xyz,
                      new_xyz.data_ptr(), B, n, m, ns, float(radius), Wp.data_ptr(), b0.data_ptr(), h.data_ptr(), N0,
                      N0, KC, n_chunks, BN, n_sub, self.split)
        else:
            # SA1 (3 colour channels, 8 floats per grouped row): a streaming gather kernel is cheaper
            kp = self.W[name + ".0"][0].shape[1]
            g = self._empty(B * m * ns, kp)
            _lib.call("bd_group_rows", xyz.data_ptr(), ld_xyz, feats.data_ptr(), ld_feats, C, new_xyz.data_ptr(),
                      idx.data_ptr(), B, n, m, ns, float(radius), g.data_ptr(), kp)
            h = self.lin(g, name + ".0", relu=True)
        h = self.lin(h, name + ".1", relu=True)
        W2, b2 = self.W[name + ".2"]
        k2 = name + ".2#wide"
        if k2 not in self._tc:
            self._tc[k2] = pack_weight_tc(W2, self.split, wide=True)
        Wp2, (BN, KC, n_chunks, n_sub) = self._tc[k2]
        out = self._empty(B, m, W2.shape[0])
        _lib.call("bd_linear_pool_tc", h.data_ptr(), h.stride(0), Wp2.data_ptr(), b2.data_ptr(), out.data_ptr(),
                  W2.shape[0], B * m * ns, W2.shape[0], W2.shape[1], KC, n_chunks, BN, n_sub, ns, self.split)
        return out

    def fp_level(self, name, unknown, known, unknown_feats, known_feats, B, n, m):
        """PointnetFPModule.forward (pointnet2_modules.py:371-416), token-major."""
        dist2 = self._empty(B, n, 3)
        idx = self._empty(B, n, 3, dtype=torch.int32)
        _lib.call("bd_three_nn", unknown.data_ptr(), known.data_ptr(), B, n, m, dist2.data_ptr(), idx.data_ptr())
        C2, C1 = known_feats.shape[-1], unknown_feats.shape[-1]
        if self.half and C1 % 8 == 0 and C2 % 8 == 0:  # fp16 rows: the operand of the first layer; fp16 hidden layer
            x = self._empty(B * n, C1 + C2, dtype=torch.float16)
            _lib.call("bd_fp_interp_concat_h", dist2.data_ptr(), idx.data_ptr(), known_feats.data_ptr(), C2,
                      unknown_feats.data_ptr(), C1, B, n, m, x.data_ptr(), 1)
        else:
            x = self._empty(B * n, C1 + C2)
            _lib.call("bd_fp_interp_concat", dist2.data_ptr(), idx.data_ptr(), known_feats.data_ptr(), C2,
                      unknown_feats.data_ptr(), C1, B, n, m, x.data_ptr())
        h = self.lin(x, name + ".0", relu=True, half_out=True)
        return self.lin(h, name + ".1", relu=True).view(B, n, -1)

    def backbone(self, pc, ep):
        """Pointnet2Backbone.forward (models/backbone_module.py:92-144)."""
        B, N, ld = pc.shape
        C_in = ld - 3
        main = torch.cuda.current_stream()
        xyz, ld_xyz, n = pc, ld, N
        feats, ld_feats, C = pc[..., 3:], ld, C_in
        self._grid = None
        lib = _lib.load()
        if FPS_GRID_MIN_POINTS <= N <= lib.bd_fps_grid_capacity() and B >= FPS_GRID_MIN_BATCH:
            # the cell list of SA1's ball query is built first: its cell order also drives the
            # bucketed furthest-point sampling (bd_fps_grid)
            ws = self._empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8)
            _lib.call("bd_grid_build", pc.data_ptr(), ld, B, N, float(SA_CFG[0][2]), ws.data_ptr())
            self._grid = (ws, N, float(SA_CFG[0][2]))
            inds1 = self._empty(B, SA_CFG[0][1], dtype=torch.int32)
            scratch = self._empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8)
            _lib.call("bd_fps_grid", pc.data_ptr(), ld, B, N, SA_CFG[0][1], ws.data_ptr(), scratch.data_ptr(),
                      inds1.data_ptr())
        else:
            inds1 = self.fps(pc, ld, B, N, SA_CFG[0][1])
        xyz1 = self.gather_rows(pc, ld, inds1, B, N, SA_CFG[0][1], 3)
        # coordinates of levels 2-4 only depend on xyz: run their (serial) FPS chain on a side
        # stream while the main stream does ball query + MLPs
        levels = {"sa1": (xyz1, inds1)}
        for name, m, _, _ in SA_CFG[1:]:  # allocated on `main`, written on the side stream
            levels[name] = (self._empty(B, m, 3), self._empty(B, m, dtype=torch.int32))
        ready = {}
        self.side_stream.wait_stream(main)
        with torch.cuda.stream(self.side_stream):
            prev, n_prev = xyz1, SA_CFG[0][1]
            for name, m, _, _ in SA_CFG[1:]:
                nxt, inds = levels[name]
                self.fps(prev, 3, B, n_prev, m, out=inds)
                self.gather_rows(prev, 3, inds, B, n_prev, m, 3, out=nxt)
                ready[name] = self.side_stream.record_event()
                prev, n_prev = nxt, m
        for name, m, radius, ns in SA_CFG:
            if name in ready:
                main.wait_event(ready[name])
            new_xyz, inds = levels[name]
            with _range("butd/" + name):
                f = self.sa_level(name, xyz, ld_xyz, feats, ld_feats, C, new_xyz, B, n, m, radius, ns)
            ep[name + "_xyz"] = new_xyz
            ep[name + "_features_tm"] = f
            if name in ("sa1", "sa2"):
                ep[name + "_inds"] = inds
            xyz, ld_xyz, n = new_xyz, 3, m
            feats, ld_feats, C = f, f.shape[-1], f.shape[-1]
        main.wait_stream(self.side_stream)  # join (also required to end a CUDA-graph capture)
        f = self.fp_level("fp1", ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features_tm"], ep["sa4_features_tm"], B, 512, 256)
        f = self.fp_level("fp2", ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features_tm"], f, B, 1024, 512)
        for name in ("sa1", "sa2", "sa3", "sa4"):
            ep[name + "_features"] = ep.pop(name + "_features_tm").transpose(1, 2)
        ep["fp2_features"] = f.transpose(1, 2)
        ep["fp2_xyz"] = ep["sa2_xyz"]
        ep["fp2_inds"] = ep["sa1_inds"][:, :ep["fp2_xyz"].shape[1]]
        return f  # (B, 1024, 288) token-major seed features

    # ---- CUDA-graph replay of the whole forward
    @torch.no_grad()
    def forward_graphed(self, inputs):
        """Same result as forward(), replayed from a CUDA graph captured per input-shape
        signature.  Inputs are copied into static buffers; the returned tensors are the graph's
        static outputs and are overwritten by the next call with the same shapes."""
        keys = [k for k in ("point_clouds", "seed_features", "seed_xyz", "seed_inds", "text_hidden",
                            "text_attention_mask", "det_boxes", "det_bbox_label_mask", "det_class_ids") if k in inputs]
        # the tokenizer pads to the longest utterance of the batch, so L varies from batch to batch:
        # pad it to a multiple of 16 (padding tokens are masked keys: they contribute exactly nothing)
        # so that a handful of graphs serves every length; the L-shaped outputs are sliced back
        L = inputs["text_hidden"].shape[1]
        Lp = _round_up(L, GRAPH_TOKEN_BUCKET)
        if Lp != L:
            inputs = dict(inputs)
            inputs["text_hidden"] = torch.nn.functional.pad(inputs["text_hidden"], (0, 0, 0, Lp - L))
            inputs["text_attention_mask"] = torch.nn.functional.pad(inputs["text_attention_mask"], (0, Lp - L))
        sig = tuple((k, tuple(inputs[k].shape), inputs[k].dtype) for k in keys)
        if not hasattr(self, "_graphs"):
            self._graphs = {}  # insertion-ordered: least recently used first
        entry = self._graphs.pop(sig, None)
        if entry is None:
            while len(self._graphs) >= GRAPH_CACHE_SIZE:  # every graph pins its own pool of intermediates
                del self._graphs[next(iter(self._graphs))]
            static_in = {k: inputs[k].detach().clone().contiguous() for k in keys}
            with torch.cuda.device(self.device):
                warm = torch.cuda.Stream()
                warm.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(warm):  # warm-up off the default stream (lazy module loads etc.)
                    self.forward(static_in)
                torch.cuda.current_stream().wait_stream(warm)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count
                with torch.cuda.graph(graph):
                    static_out = self.forward(static_in)
                entry = (graph, static_in, static_out, _lib.launch_count - n0)
        self._graphs[sig] = entry  # (re-)inserted last = most recently used
        graph, static_in, static_out, n_launch = entry
        for k in keys:
            static_in[k].copy_(inputs[k], non_blocking=True)
        graph.replay()
        _lib.launch_count += n_launch
        out = dict(static_out)
        if Lp != L:
            for k in ("text_feats", "text_attention_mask", "text_memory", "proj_tokens"):
                if k in out:
                    out[k] = out[k][:, :L]
        return out

    # ---- whole forward
    @torch.no_grad()
    def forward(self, inputs, overrides=None):
        """inputs: point_clouds (B,N,3+C) f32, text_hidden (B,L,768) f32, text_attention_mask
        (B,L) {0,1} (HF convention), det_boxes (B,D,6) f32, det_bbox_label_mask (B,D) bool,
        det_class_ids (B,D) i64 — all on the engine's device.  Returns the end_points dict of
        SURVEY.md Appendix B.  `overrides` (tests): {'sample_inds': (B,Q) i32} teacher-forces
        the query selection."""
        cfg = self.cfg
        ov = overrides or {}
        E = self.d_model
        if "seed_features" in inputs and "seed" not in ov:  # attention-only entry (BASELINE.json configs[3])
            ov = dict(ov)
            ov["seed"] = {"features": inputs["seed_features"].transpose(1, 2), "xyz": inputs["seed_xyz"],
                          "inds": inputs["seed_inds"]}
        pc = None
        if "seed" not in ov:
            pc = inputs["point_clouds"].contiguous().float()
            _lib.check_cuda(pc)
        _lib.check_cuda(inputs["text_hidden"])
        B = inputs["text_hidden"].shape[0]
        ep = {}
        self._live = []
        self._shadow = {}
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            if "seed" in ov:  # backbone output supplied: (B,V,E) token-major features, (B,V,3), (B,V) i32
                sd = ov["seed"]
                vis = sd["features"].contiguous().float()
                ep["fp2_xyz"], ep["fp2_inds"] = sd["xyz"].contiguous().float(), sd["inds"]
                ep["fp2_features"] = vis.transpose(1, 2)
            else:
                with _range("butd/backbone"):
                    vis = self.backbone(pc, ep)  # (B,V,E)
            V = vis.shape[1]
            vis = vis.reshape(B * V, E)
            ep["seed_inds"], ep["seed_xyz"] = ep["fp2_inds"], ep["fp2_xyz"]
            xyz = ep["fp2_xyz"]
            # text projector: Linear(768,E) + LayerNorm(eps=1e-12)  (models/bdetr.py:79-83)
            th = inputs["text_hidden"].contiguous().float()
            L = th.shape[1]
            text = self.add_ln(self.lin(th.view(B * L, -1), "text_projector"), None, "text_projector.ln", eps=1e-12)
            text_mask = inputs["text_attention_mask"].ne(1)
            tmask_u8 = text_mask.to(torch.uint8).contiguous()
            ep["text_feats"], ep["text_attention_mask"] = text.view(B, L, E), text_mask
            det = dmask_u8 = None
            D = 0
            if cfg["butd"]:
                boxes = inputs["det_boxes"].contiguous().float()
                D = boxes.shape[1]
                dmask_u8 = (~inputs["det_bbox_label_mask"]).to(torch.uint8).contiguous()
                det = self._empty(B * D, E)
                h = self.lin(boxes.view(B * D, 6), "box_embeddings.0", relu=True, half_out=True)
                self.lin(h, "box_embeddings.1", out=det[:, :128])
                table = self.W["butd_class_embeddings"][0]
                emb = self._empty(B * D, table.shape[1])
                ids = inputs["det_class_ids"].contiguous().to(torch.int64)
                _lib.call("bd_embedding_rows", table.data_ptr(), table.shape[1], ids.data_ptr(), B * D,
                          emb.data_ptr(), table.shape[1])
                self.lin(emb, "class_embeddings", out=det[:, 128:])
            pos = self.posembed(xyz.reshape(B * V, 3), "pos_embed")
            # ---- BiEncoder (encoder_decoder_layers.py:225-255, 75-124)
            # visual half on the current stream, language half on text_stream; they exchange
            # their self-attended states once per layer (cross_lv needs vis, cross_vl needs text)
            ts = self.text_stream
            ts.wait_event(main.record_event())
            for i in range(cfg["num_encoder_layers"]):
                k = f"enc{i}"
                torch.cuda.nvtx.range_push(f"butd/encoder_layer{i}") if NVTX else None
                if cfg["self_attend"]:
                    vis = self.mha(k + ".sv", vis, pos, vis, pos, B, V, V, None, True, vis, k + ".sv.ln")
                    with torch.cuda.stream(ts):
                        text = self.mha(k + ".sl", text, None, text, None, B, L, L, tmask_u8, True, text, k + ".sl.ln")
                ev_vis, ev_text = main.record_event(), ts.record_event()
                main.wait_event(ev_text)
                ts.wait_event(ev_vis)
                text_kv = text  # cross_vl attends to the text BEFORE the cross_lv update (:84)
                with torch.cuda.stream(ts):
                    text = self.mha(k + ".lv", text, None, vis, None, B, L, V, None, False, text, k + ".norm_lv")
                    text = self.ffn(text, k + ".ffn_lv", k + ".norm_lv2")
                vis = self.mha(k + ".vl", vis, pos, text_kv, None, B, V, L, tmask_u8, False, vis, k + ".norm_vl")
                if cfg["butd"]:
                    vis = self.mha(k + ".d", vis, None, det, None, B, V, D, dmask_u8, False, vis, k + ".norm_d")
                vis = self.ffn(vis, k + ".ffn_vl", k + ".norm_vl2")
                torch.cuda.nvtx.range_pop() if NVTX else None
            main.wait_stream(ts)
            # memory-side K/V projections of every decoder layer: functions of the encoder output
            # only, so they start now and run beside query generation and the earlier layers
            kvs = self.kv_stream
            kvs.wait_event(main.record_event())
            mem_kv = []
            with torch.cuda.stream(kvs):
                for i in range(cfg["num_decoder_layers"]):
                    k = f"dec{i}"
                    kv_l = self.lin(text, k + ".l.kv", half_out=True)
                    kv_d = self.lin(det, k + ".d.kv", half_out=True) if cfg["butd"] else None
                    kv_v = self.lin(vis, k + ".v.kv", half_out=True)
                    # ... and their K / V^T operand tiles: the decoder's attention calls then start at the
                    # attention kernel itself (whole batch only: the tile image is not sliceable by scene)
                    Qn = cfg["num_queries"]
                    packs = (None, None, None)
                    if B < 2 * DECODER_SPLIT_MIN:
                        packs = (self.pack_kv(kv_l, B, Qn, L), self.pack_kv(kv_d, B, Qn, D) if cfg["butd"] else None,
                                 self.pack_kv(kv_v, B, Qn, V))
                    mem_kv.append((kv_l, kv_d, kv_v, kvs.record_event(), packs))
            ep["text_memory"] = text.view(B, L, E)
            ep["seed_features"] = vis.view(B, V, E).transpose(1, 2)
            if cfg["contrastive_align_loss"]:  # nothing downstream reads it: off the critical path
                self.head_stream.wait_event(main.record_event())
                with torch.cuda.stream(self.head_stream):
                    ep["proj_tokens"] = self.contrastive(text, "text", B, L)
            # ---- query generation (models/bdetr.py:177-191)
            torch.cuda.nvtx.range_push("butd/query_generation") if NVTX else None
            h = self.lin(vis, "points_obj_cls.conv1", relu=True, half_out=True)
            h = self.lin(h, "points_obj_cls.conv2", relu=True, half_out=True)
            logits = self.lin(h, "points_obj_cls.conv3")  # (B*V, 1)
            ep["seeds_obj_cls_logits"] = logits.view(B, V, 1).transpose(1, 2)
            Q = cfg["num_queries"]
            if "sample_inds" in ov:
                sample_inds = ov["sample_inds"].to(self.device, torch.int32).contiguous()
            else:
                sample_inds = self._empty(B, Q, dtype=torch.int32)
                _lib.call("bd_topk_sigmoid", logits.data_ptr(), B, V, Q, sample_inds.data_ptr())
            cluster_xyz = self.gather_rows(xyz, 3, sample_inds, B, V, Q, 3).view(B * Q, 3)
            cluster_feat = self.gather_rows(vis, E, sample_inds, B, V, Q, E).view(B * Q, E)
            ep["query_points_xyz"] = cluster_xyz.view(B, Q, 3)
            ep["query_points_feature"] = cluster_feat.view(B, Q, E).transpose(1, 2)
            ep["query_points_sample_inds"] = sample_inds
            query = self.lin(cluster_feat, "decoder_query_proj")
            torch.cuda.nvtx.range_pop() if NVTX else None
            nd = cfg["num_decoder_layers"]
            spe = cfg["self_position_embedding"]
            hs = self.head_stream
            n_cls = self.W["proposal_head.sem.2"][0].shape[0]
            prefixes = ["proposal_"] + ["last_" if i == nd - 1 else f"{i}head_" for i in range(nd)]
            outs = {}
            for pf in prefixes:  # full-batch outputs; the decoder parts below fill their rows
                outs[pf] = {"center": self._empty(B * Q, 3), "size": self._empty(B * Q, 3),
                            "sem": self._empty(B * Q, n_cls)}
                if cfg["contrastive_align_loss"]:
                    outs[pf]["proj"] = self._empty(B * Q, self.W["contrastive.image.2"][0].shape[0])

            def rows(t, b0, b1, per):
                return None if t is None else t[b0 * per:b1 * per]

            def decode(b0, b1):
                """Proposal head + the decoder layers for scenes [b0, b1) on the current stream
                (models/bdetr.py:261-317, encoder_decoder_layers.py:340-406)."""
                cur = torch.cuda.current_stream()
                nb = b1 - b0
                q = rows(query, b0, b1, Q)
                cxyz = rows(cluster_xyz, b0, b1, Q)
                tm, dm = rows(tmask_u8, b0, b1, 1), rows(dmask_u8, b0, b1, 1)
                o = {k2: rows(v2, b0, b1, Q) for k2, v2 in outs["proposal_"].items()}
                if cfg["contrastive_align_loss"]:
                    hs.wait_event(cur.record_event())
                    with torch.cuda.stream(hs):
                        self.contrastive(q, "image", nb, Q, out=o["proj"])
                base_xyz, base_size = self.head(rows(cluster_feat, b0, b1, Q), cxyz, "proposal_head", o)
                for i in range(nd):
                    k = f"dec{i}"
                    torch.cuda.nvtx.range_push(f"butd/decoder_layer{i}") if NVTX else None
                    kv_l, kv_d, kv_v, kv_ready, (pk_l, pk_d, pk_v) = mem_kv[i]
                    if spe == "loc_learned":  # query_pos = cat(base_xyz, base_size)  (bdetr.py:287)
                        qp_in = self._empty(nb * Q, 6)
                        _lib.call("bd_concat_rows", base_xyz.data_ptr(), 3, 3, base_size.data_ptr(), 3, 3,
                                  qp_in.data_ptr(), 6, nb * Q)
                    elif spe == "xyz_learned":
                        qp_in = base_xyz
                    else:
                        qp_in = None
                    qpos = self.posembed(qp_in, k + ".posembed") if qp_in is not None else None
                    q = self.mha(k + ".self", q, qpos, q, qpos, nb, Q, Q, None, True, q, k + ".norm1")
                    cur.wait_event(kv_ready)
                    q = self.mha(k + ".l", q, qpos, None, None, nb, Q, L, tm, False, q, k + ".norm_l",
                                 kv=rows(kv_l, b0, b1, L), packed=pk_l)
                    if cfg["butd"]:
                        q = self.mha(k + ".d", q, qpos, None, None, nb, Q, D, dm, False, q, k + ".norm_d",
                                     kv=rows(kv_d, b0, b1, D), packed=pk_d)
                    q = self.mha(k + ".v", q, qpos, None, None, nb, Q, V, None, False, q, k + ".norm_v",
                                 kv=rows(kv_v, b0, b1, V), packed=pk_v)
                    q = self.ffn(q, k + ".ffn", k + ".norm2")
                    o = {k2: rows(v2, b0, b1, Q) for k2, v2 in outs[prefixes[i + 1]].items()}
                    if cfg["contrastive_align_loss"]:
                        hs.wait_event(cur.record_event())
                        with torch.cuda.stream(hs):
                            self.contrastive(q, "image", nb, Q, out=o["proj"])
                    base_xyz, base_size = self.head(q, cxyz, f"head{i}", o)
                    torch.cuda.nvtx.range_pop() if NVTX else None

            # The decoder's kernels are short (256 queries per scene) and latency-bound; two halves
            # of the batch on two streams overlap each other's dependent chains.
            halves = [(0, B)] if B < 2 * DECODER_SPLIT_MIN else [(0, B // 2), (B // 2, B)]
            ds = self.dec_stream
            if len(halves) == 2:
                ds.wait_event(main.record_event())
                with torch.cuda.stream(ds):
                    decode(*halves[1])
            decode(*halves[0])
            if len(halves) == 2:
                main.wait_stream(ds)
            main.wait_stream(kvs)
            main.wait_stream(hs)
            for i, pf in enumerate(prefixes):
                o = outs[pf]
                ep[pf + "base_xyz"] = cluster_xyz.view(B, Q, 3)  # every head is called with base_xyz = cluster_xyz
                ep[pf + "center"] = o["center"].view(B, Q, 3)
                ep[pf + "pred_size"] = o["size"].view(B, Q, 3)
                ep[pf + "sem_cls_scores"] = o["sem"].view(B, Q, -1)
                if cfg["contrastive_align_loss"]:
                    ep[pf + "proj_queries"] = o["proj"].view(B, Q, -1)
        return ep
