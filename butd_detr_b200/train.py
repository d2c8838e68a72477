"""Training step of `BeaUTyDETR`: train-mode forward with autograd, and the
ONE gradient exchange of data-parallel training.

What is native and what is not (said plainly):
  * the nine PointNet++ operators — forward AND backward (`gather_points_grad`, `group_points_grad`,
    `three_interpolate_grad`: scatter-add kernels of libbutd_b200, replacing
    `/root/reference/pointnet2/_ext_src/src/{sampling,group_points,interpolate}_gpu.cu`) — run on this
    package's sm_100a kernels through `torch.autograd.Function`s with the reference's signatures
    (`pointnet2/pointnet2_utils.py:97-114,164-203,214-254`);
  * the dense layers of the training step (1x1 convolutions, batch-statistics BatchNorm, attention,
    LayerNorm, dropout) and their backward are PyTorch operators on the module's own parameters — the
    tcgen05 forward kernels of the eval path have no backward yet.  The maths is the reference's
    (`models/bdetr.py:193-319`, `models/encoder_decoder_layers.py`, `models/modules.py`,
    `pointnet2/pointnet2_modules.py`), dropout placement included;
  * gradients live in ONE flat fp32 arena (`GradArena`): every `p.grad` is a view into it, autograd
    accumulates in place, and the data-parallel exchange is a single `all_reduce` over the arena (NCCL
    over NVLink; the reference's DDP issues one all-reduce per 25 MB bucket, `main_utils.py:310-313`).
"""
import math

import torch
import torch.nn.functional as F

from . import pointnet2_ext as _ext

SA_CFG = (("sa1", 2048, 0.2, 64), ("sa2", 1024, 0.4, 32), ("sa3", 512, 0.8, 16), ("sa4", 256, 1.2, 16))
BN_MOMENTUM = 0.1   # models/bdetr.py:321-325 (init_bn_momentum)
DROPOUT = 0.1       # transformer layers, text projector (models/bdetr.py:83,100-131)
HEAD_DROPOUT = 0.3  # ThreeLayerMLP (models/modules.py:89-108)


# ------------------------------------------------------------------ the point operators with autograd
class _GatherPoints(torch.autograd.Function):  # pointnet2_utils.py:97-114
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return _ext.gather_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n), None


class _GroupPoints(torch.autograd.Function):  # pointnet2_utils.py:214-254
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.shape[2]
        return _ext.group_points(features.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


class _ThreeInterpolate(torch.autograd.Function):  # pointnet2_utils.py:164-203
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m = features.shape[2]
        return _ext.three_interpolate(features.contiguous(), idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m), None, None


gather_points, group_points, three_interpolate = _GatherPoints.apply, _GroupPoints.apply, _ThreeInterpolate.apply


class _Weights:
    """Parameters / buffers of the module by their reference state_dict names."""

    def __init__(self, model):
        self.t = dict(model.named_parameters())
        self.t.update(dict(model.named_buffers()))

    def __getitem__(self, k):
        return self.t[k]

    def get(self, k):
        return self.t.get(k)


class _Fwd:
    """One train-mode forward: functional layers over the module's own tensors."""

    def __init__(self, model, dropout=True):
        self.w = _Weights(model)
        self.cfg = model.cfg
        self.p_drop = DROPOUT if dropout else 0.0
        self.p_head = HEAD_DROPOUT if dropout else 0.0

    # ---- layers
    def bn(self, x, pre):
        w = self.w
        w[pre + ".num_batches_tracked"].add_(1)
        return F.batch_norm(x, w[pre + ".running_mean"], w[pre + ".running_var"], w[pre + ".weight"], w[pre + ".bias"],
                            True, BN_MOMENTUM, 1e-5)

    def ln(self, x, pre, eps=1e-5):
        return F.layer_norm(x, (x.shape[-1],), self.w[pre + ".weight"], self.w[pre + ".bias"], eps)

    def lin(self, x, pre):
        return F.linear(x, self.w[pre + ".weight"], self.w.get(pre + ".bias"))

    def conv1(self, x, pre):
        return F.conv1d(x, self.w[pre + ".weight"], self.w.get(pre + ".bias"))

    def drop(self, x, p=None):
        p = self.p_drop if p is None else p
        return F.dropout(x, p, True) if p > 0 else x

    def mha(self, pre, q, k, v, mask=None):
        """nn.MultiheadAttention.forward (seq-first), dropout on the attention weights, `[0]` taken."""
        w = self.w
        E = q.shape[-1]
        out, _ = F.multi_head_attention_forward(
            q, k, v, E, 8, w[pre + ".in_proj_weight"], w[pre + ".in_proj_bias"], None, None, False, self.p_drop,
            w[pre + ".out_proj.weight"], w[pre + ".out_proj.bias"], training=True, key_padding_mask=mask,
            need_weights=False)
        return out

    def pos_embed(self, pre, xyz):  # PositionEmbeddingLearned: (B,N,c) -> (B,F,N)
        h = pre + ".position_embedding_head"
        x = F.relu(self.bn(self.conv1(xyz.transpose(1, 2).contiguous(), h + ".0"), h + ".1"))
        return self.conv1(x, h + ".3")

    def three_layer_mlp(self, pre, x):  # models/modules.py:89-108
        n = pre + ".net"
        x = self.drop(F.relu(self.bn(self.conv1(x, n + ".0"), n + ".1")), self.p_head)
        x = self.drop(F.relu(self.bn(self.conv1(x, n + ".4"), n + ".5")), self.p_head)
        return self.conv1(x, n + ".8")

    def head(self, pre, features, base_xyz, ep, prefix):  # ClsAgnosticPredictHead (models/modules.py:135-180)
        center = base_xyz + self.three_layer_mlp(pre + ".center_residual_head", features).transpose(2, 1)
        size = self.three_layer_mlp(pre + ".size_pred_head", features).transpose(2, 1)
        sem = self.three_layer_mlp(pre + ".sem_cls_scores_head", features).transpose(2, 1)
        ep[prefix + "base_xyz"], ep[prefix + "center"] = base_xyz, center
        ep[prefix + "pred_size"], ep[prefix + "sem_cls_scores"] = size, sem
        return center, size

    def contrastive(self, pre, x):  # models/bdetr.py:137-151
        x = F.relu(self.lin(x, pre + ".0"))
        x = F.relu(self.lin(x, pre + ".2"))
        return F.normalize(self.lin(x, pre + ".4"), p=2, dim=-1)

    # ---- backbone (pointnet2_modules.py:210-272, 371-416; backbone_module.py:92-144)
    def shared_mlp(self, pre, x, n_layers):
        for i in range(n_layers):
            p = f"{pre}.layer{i}"
            x = F.relu(self.bn(F.conv2d(x, self.w[p + ".conv.weight"]), p + ".bn.bn"))
        return x

    def sa_module(self, pre, xyz, features, npoint, radius, nsample):
        inds = _ext.furthest_point_sampling(xyz.contiguous(), npoint)
        xyz_t = xyz.transpose(1, 2).contiguous()
        new_xyz = gather_points(xyz_t, inds).transpose(1, 2).contiguous()
        idx = _ext.ball_query(new_xyz, xyz.contiguous(), radius, nsample)
        g_xyz = (group_points(xyz_t, idx) - new_xyz.transpose(1, 2).unsqueeze(-1)) / radius
        grouped = torch.cat([g_xyz, group_points(features, idx)], dim=1)
        new_feat = self.shared_mlp(pre + ".mlp_module", grouped, 3)
        new_feat = F.max_pool2d(new_feat, kernel_size=[1, new_feat.size(3)]).squeeze(-1)
        return new_xyz, new_feat, inds

    def fp_module(self, pre, unknown, known, unknown_feats, known_feats):
        dist2, idx = _ext.three_nn(unknown.contiguous(), known.contiguous())
        dist_recip = 1.0 / (torch.sqrt(dist2) + 1e-8)
        weight = (dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)).contiguous()
        interp = three_interpolate(known_feats, idx, weight)
        new = torch.cat([interp, unknown_feats], dim=1).unsqueeze(-1)
        return self.shared_mlp(pre + ".mlp", new, 2).squeeze(-1)

    def backbone(self, pc, ep):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous()
        for name, npoint, radius, nsample in SA_CFG:
            xyz, features, inds = self.sa_module(f"backbone_net.{name}", xyz, features, npoint, radius, nsample)
            if name in ("sa1", "sa2"):
                ep[name + "_inds"] = inds
            ep[name + "_xyz"], ep[name + "_features"] = xyz, features
        f = self.fp_module("backbone_net.fp1", ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features"], ep["sa4_features"])
        f = self.fp_module("backbone_net.fp2", ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features"], f)
        ep["fp2_features"], ep["fp2_xyz"] = f, ep["sa2_xyz"]
        ep["fp2_inds"] = ep["sa1_inds"][:, 0:ep["fp2_xyz"].shape[1]]

    # ---- transformer (encoder_decoder_layers.py)
    def ffn(self, x, pre):
        return self.drop(self.lin(self.drop(F.relu(self.lin(x, pre + ".0"))), pre + ".3"))

    def encoder_layer(self, pre, vis, pos, text, text_mask, det, det_mask):
        if self.cfg["self_attend"]:
            s, p = vis.transpose(0, 1), pos.transpose(0, 1)
            a = pre + ".self_attention_visual"
            s = self.ln(s + self.drop(self.mha(a + ".self_attn", s + p, s + p, s)), a + ".norm1")
            vis = s.transpose(0, 1)
            t = text.transpose(0, 1)
            a = pre + ".self_attention_lang"
            t = self.ln(t + self.drop(self.mha(a + ".self_attn", t, t, t, text_mask)), a + ".norm1")
            text = t.transpose(0, 1)
        c = pre + ".cross_layer"
        qv, kt = vis + pos, text
        t2 = self.mha(c + ".cross_lv", text.transpose(0, 1), vis.transpose(0, 1), vis.transpose(0, 1)).transpose(0, 1)
        text = self.ln(text + self.drop(t2), c + ".norm_lv")
        text = self.ln(text + self.ffn(text, c + ".ffn_lv"), c + ".norm_lv2")
        v2 = self.mha(c + ".cross_vl", qv.transpose(0, 1), kt.transpose(0, 1), kt.transpose(0, 1), text_mask).transpose(0, 1)
        vis = self.ln(vis + self.drop(v2), c + ".norm_vl")
        if det is not None:
            v2 = self.mha(c + ".cross_d", vis.transpose(0, 1), det.transpose(0, 1), det.transpose(0, 1), det_mask).transpose(0, 1)
            vis = self.ln(vis + self.drop(v2), c + ".norm_d")
        vis = self.ln(vis + self.ffn(vis, c + ".ffn_vl"), c + ".norm_vl2")
        return vis, text

    def decoder_layer(self, pre, query, vis, text, query_pos, text_mask, det, det_mask):
        if query_pos is not None:
            qp = self.pos_embed(pre + ".self_posembed", query_pos).transpose(1, 2).contiguous()
        else:
            qp = torch.zeros_like(query)
        q, qp = query.transpose(0, 1), qp.transpose(0, 1)
        q = self.ln(q + self.drop(self.mha(pre + ".self_attn", q + qp, q + qp, q)), pre + ".norm1")
        lt = text.transpose(0, 1)
        q = self.ln(q + self.drop(self.mha(pre + ".cross_l", q + qp, lt, lt, text_mask)), pre + ".norm_l")
        if det is not None:
            dt = det.transpose(0, 1)
            q = self.ln(q + self.drop(self.mha(pre + ".cross_d", q + qp, dt, dt, det_mask)), pre + ".norm_d")
        vt = vis.transpose(0, 1)
        q = self.ln(q + self.drop(self.mha(pre + ".cross_v", q + qp, vt, vt)), pre + ".norm_v")
        q = self.ln(q + self.ffn(q, pre + ".ffn"), pre + ".norm2")
        return q.transpose(0, 1).contiguous()


def forward_train(model, inputs, text_hidden, hf_mask, dropout=True):
    """Train-mode `BeaUTyDETR.forward` (models/bdetr.py:193-319): BatchNorm on batch statistics
    (running statistics updated in place), dropout active (`dropout=False` switches every dropout off —
    the deterministic configuration the gradient-parity tests use)."""
    f = _Fwd(model, dropout)
    cfg, w = model.cfg, f.w
    ep = {}
    f.backbone(inputs["point_clouds"].float(), ep)
    ep["seed_inds"], ep["seed_xyz"], ep["seed_features"] = ep["fp2_inds"], ep["fp2_xyz"], ep["fp2_features"]
    text = f.drop(f.ln(f.lin(text_hidden, "text_projector.0"), "text_projector.1", eps=1e-12))
    text_mask = hf_mask.ne(1).bool()
    ep["text_feats"], ep["text_attention_mask"] = text, text_mask
    xyz, feats = ep["fp2_xyz"], ep["fp2_features"]
    det = det_mask = None
    if cfg["butd"]:  # models/bdetr.py:217-225
        det_mask = ~inputs["det_bbox_label_mask"]
        cls = f.lin(F.embedding(inputs["det_class_ids"], w["butd_class_embeddings.weight"]), "class_embeddings")
        det = torch.cat([f.pos_embed("box_embeddings", inputs["det_boxes"].float()), cls.transpose(1, 2)], 1)
        det = det.transpose(1, 2).contiguous()
    vis = feats.transpose(1, 2).contiguous()
    pos = f.pos_embed("pos_embed", xyz).transpose(1, 2).contiguous()
    for i in range(cfg["num_encoder_layers"]):
        vis, text = f.encoder_layer(f"cross_encoder.layers.{i}", vis, pos, text, text_mask, det, det_mask)
    feats = vis.transpose(1, 2).contiguous()
    ep["text_memory"], ep["seed_features"] = text, feats
    if cfg["contrastive_align_loss"]:
        ep["proj_tokens"] = f.contrastive("contrastive_align_projection_text", text)
    # query generation (models/bdetr.py:177-191)
    h = F.relu(f.bn(f.conv1(feats, "points_obj_cls.conv1"), "points_obj_cls.bn1"))
    h = F.relu(f.bn(f.conv1(h, "points_obj_cls.conv2"), "points_obj_cls.bn2"))
    logits = f.conv1(h, "points_obj_cls.conv3")
    ep["seeds_obj_cls_logits"] = logits
    Q = cfg["num_queries"]
    sample_inds = torch.topk(torch.sigmoid(logits).squeeze(1), Q)[1].int()
    cluster_xyz = gather_points(xyz.transpose(1, 2).contiguous(), sample_inds).transpose(1, 2).contiguous()
    cluster_feat = gather_points(feats, sample_inds).contiguous()
    ep["query_points_xyz"], ep["query_points_feature"] = cluster_xyz, cluster_feat
    ep["query_points_sample_inds"] = sample_inds
    query = f.conv1(cluster_feat, "decoder_query_proj").transpose(1, 2).contiguous()
    if cfg["contrastive_align_loss"]:
        ep["proposal_proj_queries"] = f.contrastive("contrastive_align_projection_image", query)
    base_xyz, base_size = f.head("proposal_head", cluster_feat, cluster_xyz, ep, "proposal_")
    base_xyz, base_size = base_xyz.detach().clone(), base_size.detach().clone()  # models/bdetr.py:271-272
    nd, spe = cfg["num_decoder_layers"], cfg["self_position_embedding"]
    for i in range(nd):
        prefix = "last_" if i == nd - 1 else f"{i}head_"
        if spe == "loc_learned":
            query_pos = torch.cat([base_xyz, base_size], -1)
        elif spe == "xyz_learned":
            query_pos = base_xyz
        else:
            query_pos = None
        query = f.decoder_layer(f"decoder.{i}", query, vis, text, query_pos, text_mask, det, det_mask)
        if cfg["contrastive_align_loss"]:
            ep[prefix + "proj_queries"] = f.contrastive("contrastive_align_projection_image", query)
        base_xyz, base_size = f.head(f"prediction_heads.{i}", query.transpose(1, 2).contiguous(), cluster_xyz, ep, prefix)
        base_xyz, base_size = base_xyz.detach().clone(), base_size.detach().clone()  # models/bdetr.py:314-315
    return ep


class GradArena:
    """All gradients of a module in ONE flat fp32 buffer: every `p.grad` is a view into it, so
    autograd accumulates in place and data-parallel training needs a single collective."""

    def __init__(self, module, skip=("text_encoder.",)):
        self.params = [(n, p) for n, p in module.named_parameters()
                       if p.requires_grad and not any(n.startswith(s) for s in skip)]
        total = sum(p.numel() for _, p in self.params)
        dev = self.params[0][1].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for _, p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def check_views(self):
        """True while every p.grad still aliases the arena (an optimizer's `zero_grad(set_to_none=True)`
        would break that — use arena.zero() instead)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for _, p in self.params)

    def all_reduce(self, group=None, average=True):
        """The one exchange of the data-parallel step: sum (or mean) of the arena over the ranks."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            self.flat.div_(dist.get_world_size(group))
