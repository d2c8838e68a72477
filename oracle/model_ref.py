"""TEST INFRASTRUCTURE — plain-PyTorch fp32 CPU restatement of the reference's eval-mode
`BeaUTyDETR.forward` (`/root/reference/models/bdetr.py:193-319`) as a function of a
`state_dict` with the reference's own key names, running on `oracle/point_ops` (the C
restatement of the CUDA-only point ops).

Why it exists: `/root/reference` (Python) cannot travel to the GPU box, and its native ops
have no CPU path, so the checker used there — and the "reference CPU forward" that
`bench.py --impl reference` / `cpu_baseline` time — must be a port.  It is pinned against
the reference's real modules: `tests/golden/make_model_golden.py` imports the reference here
(`oracle/ref_loader.py`), runs both on the same seeded inputs/weights and commits the
reference's outputs; `tests/test_oracle_model.py` checks this file against them.

Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may import this module.
Every function cites the reference lines it restates.
"""
import math

import torch
import torch.nn.functional as F

from . import point_ops as P


# ----------------------------------------------------------------------------- helpers
def _bn(x, sd, pre, eps=1e-5):
    """nn.BatchNorm{1,2}d in eval mode (running statistics)."""
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"],
                        sd[pre + ".weight"], sd[pre + ".bias"], False, 0.0, eps)


def _ln(x, sd, pre, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + ".weight"], sd[pre + ".bias"], eps)


def _lin(x, sd, pre):
    return F.linear(x, sd[pre + ".weight"], sd.get(pre + ".bias"))


def _conv1(x, sd, pre):
    """nn.Conv1d(kernel_size=1) on (B, C, N)."""
    return F.conv1d(x, sd[pre + ".weight"], sd.get(pre + ".bias"))


def mha(sd, pre, query, key, value, key_padding_mask=None, n_heads=8):
    """nn.MultiheadAttention.forward, batch_first=False, eval (dropout off), returning only
    the attention output (call sites take `[0]`), e.g. encoder_decoder_layers.py:87-93.
    query (Lq,B,E), key/value (Lk,B,E), key_padding_mask (B,Lk) bool, True = ignore.
    Maths of torch.nn.functional.multi_head_attention_forward: packed in-proj, q scaled by
    1/sqrt(head_dim) before QK^T, -inf additive mask, softmax, PV, out-proj."""
    Lq, B, E = query.shape
    Lk = key.shape[0]
    hd = E // n_heads
    W, b = sd[pre + ".in_proj_weight"], sd[pre + ".in_proj_bias"]
    q = F.linear(query, W[:E], b[:E])
    k = F.linear(key, W[E:2 * E], b[E:2 * E])
    v = F.linear(value, W[2 * E:], b[2 * E:])
    q = q.reshape(Lq, B * n_heads, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(Lk, B * n_heads, hd).transpose(0, 1)
    v = v.reshape(Lk, B * n_heads, hd).transpose(0, 1)
    attn = torch.bmm(q, k.transpose(1, 2))  # (B*H, Lq, Lk)
    if key_padding_mask is not None:
        m = key_padding_mask.view(B, 1, 1, Lk).expand(B, n_heads, Lq, Lk).reshape(B * n_heads, Lq, Lk)
        attn = attn.masked_fill(m, float("-inf"))
    attn = torch.softmax(attn, dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(Lq, B, E)
    return F.linear(out, sd[pre + ".out_proj.weight"], sd[pre + ".out_proj.bias"])


def pos_embed_learned(sd, pre, xyz):
    """PositionEmbeddingLearned.forward (models/modules.py:52-67 and
    encoder_decoder_layers.py:19-34): (B,N,c) -> (B,F,N)."""
    h = pre + ".position_embedding_head"
    x = xyz.transpose(1, 2).contiguous()
    x = F.relu(_bn(_conv1(x, sd, h + ".0"), sd, h + ".1"))
    return _conv1(x, sd, h + ".3")


def three_layer_mlp(sd, pre, x):
    """ThreeLayerMLP (models/modules.py:89-108), eval: dropout off."""
    n = pre + ".net"
    x = F.relu(_bn(_conv1(x, sd, n + ".0"), sd, n + ".1"))
    x = F.relu(_bn(_conv1(x, sd, n + ".4"), sd, n + ".5"))
    return _conv1(x, sd, n + ".8")


def predict_head(sd, pre, features, base_xyz, end_points, prefix):
    """ClsAgnosticPredictHead.forward (models/modules.py:135-180), objectness/heading off."""
    center = base_xyz + three_layer_mlp(sd, pre + ".center_residual_head", features).transpose(2, 1)
    size = three_layer_mlp(sd, pre + ".size_pred_head", features).transpose(2, 1)
    sem = three_layer_mlp(sd, pre + ".sem_cls_scores_head", features).transpose(2, 1)
    end_points[prefix + "base_xyz"] = base_xyz
    end_points[prefix + "center"] = center
    end_points[prefix + "pred_size"] = size
    end_points[prefix + "sem_cls_scores"] = sem
    return center, size


def contrastive_proj(sd, pre, x):
    """contrastive_align_projection_{image,text} (models/bdetr.py:137-151) + F.normalize."""
    x = F.relu(_lin(x, sd, pre + ".0"))
    x = F.relu(_lin(x, sd, pre + ".2"))
    return F.normalize(_lin(x, sd, pre + ".4"), p=2, dim=-1)


# ----------------------------------------------------------------------------- backbone
def shared_mlp(sd, pre, x, n_layers):
    """pt_utils.SharedMLP: k x [Conv2d 1x1 (no bias) -> BN2d -> ReLU] (pytorch_utils.py:11-36)."""
    for i in range(n_layers):
        p = f"{pre}.layer{i}"
        x = F.conv2d(x, sd[p + ".conv.weight"])
        x = F.relu(_bn(x, sd, p + ".bn.bn"))
    return x


def query_and_group(xyz, new_xyz, features, radius, nsample):
    """QueryAndGroup.forward (pointnet2_utils.py:317-376), use_xyz, normalize_xyz."""
    idx = P.ball_query(new_xyz, xyz, radius, nsample)
    g_xyz = P.group_points(xyz.transpose(1, 2).contiguous(), idx)
    g_xyz = g_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    g_xyz = g_xyz / radius
    g_feat = P.group_points(features.contiguous(), idx)
    return torch.cat([g_xyz, g_feat], dim=1)


def sa_module(sd, pre, xyz, features, npoint, radius, nsample):
    """PointnetSAModuleVotes.forward (pointnet2_modules.py:210-272), max pooling."""
    inds = P.furthest_point_sampling(xyz.contiguous(), npoint)
    new_xyz = P.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    grouped = query_and_group(xyz, new_xyz, features, radius, nsample)
    new_feat = shared_mlp(sd, pre + ".mlp_module", grouped, 3)
    new_feat = new_feat.max(dim=3)[0]
    return new_xyz, new_feat, inds


def fp_module(sd, pre, unknown, known, unknown_feats, known_feats):
    """PointnetFPModule.forward (pointnet2_modules.py:371-416)."""
    dist2, idx = P.three_nn(unknown.contiguous(), known.contiguous())
    dist = torch.sqrt(dist2)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = (dist_recip / norm).contiguous()
    interp = P.three_interpolate(known_feats.contiguous(), idx, weight)
    new = torch.cat([interp, unknown_feats], dim=1).unsqueeze(-1)
    return shared_mlp(sd, pre + ".mlp", new, 2).squeeze(-1)


SA_CFG = (("sa1", 2048, 0.2, 64), ("sa2", 1024, 0.4, 32), ("sa3", 512, 0.8, 16), ("sa4", 256, 1.2, 16))


def backbone(sd, pointcloud, pre="backbone_net"):
    """Pointnet2Backbone.forward (models/backbone_module.py:92-144)."""
    ep = {}
    xyz = pointcloud[..., 0:3].contiguous()
    features = pointcloud[..., 3:].transpose(1, 2).contiguous()
    for name, npoint, radius, nsample in SA_CFG:
        xyz, features, inds = sa_module(sd, f"{pre}.{name}", xyz, features, npoint, radius, nsample)
        if name in ("sa1", "sa2"):
            ep[name + "_inds"] = inds
        ep[name + "_xyz"] = xyz
        ep[name + "_features"] = features
    f = fp_module(sd, pre + ".fp1", ep["sa3_xyz"], ep["sa4_xyz"], ep["sa3_features"], ep["sa4_features"])
    f = fp_module(sd, pre + ".fp2", ep["sa2_xyz"], ep["sa3_xyz"], ep["sa2_features"], f)
    ep["fp2_features"] = f
    ep["fp2_xyz"] = ep["sa2_xyz"]
    ep["fp2_inds"] = ep["sa1_inds"][:, 0:ep["fp2_xyz"].shape[1]]
    return ep


# ----------------------------------------------------------------------------- transformer
def encoder_layer(sd, pre, vis, pos, text, text_mask, det, det_mask):
    """BiEncoderLayer.forward (encoder_decoder_layers.py:225-255) incl. CrossAttentionLayer
    (:75-124).  vis/pos (B,V,F), text (B,L,F), det (B,D,F); masks True = pad."""
    # vis self-attention, q = k = src + pos, v = src (:166-186)
    s, p = vis.transpose(0, 1), pos.transpose(0, 1)
    s = _ln(s + mha(sd, pre + ".self_attention_visual.self_attn", s + p, s + p, s, None),
            sd, pre + ".self_attention_visual.norm1")
    vis = s.transpose(0, 1)
    # text self-attention (:137-156)
    t = text.transpose(0, 1)
    t = _ln(t + mha(sd, pre + ".self_attention_lang.self_attn", t, t, t, text_mask),
            sd, pre + ".self_attention_lang.norm1")
    text = t.transpose(0, 1)
    c = pre + ".cross_layer"
    qv = vis + pos
    kt = text  # the text BEFORE the cross_lv update (:84)
    # language <- vision (:87-96)
    t2 = mha(sd, c + ".cross_lv", text.transpose(0, 1), vis.transpose(0, 1), vis.transpose(0, 1), None).transpose(0, 1)
    text = _ln(text + t2, sd, c + ".norm_lv")
    ff = _lin(F.relu(_lin(text, sd, c + ".ffn_lv.0")), sd, c + ".ffn_lv.3")
    text = _ln(text + ff, sd, c + ".norm_lv2")
    # vision <- language (:99-107)
    v2 = mha(sd, c + ".cross_vl", qv.transpose(0, 1), kt.transpose(0, 1), kt.transpose(0, 1), text_mask).transpose(0, 1)
    vis = _ln(vis + v2, sd, c + ".norm_vl")
    # vision <- boxes (:110-119)
    if det is not None:
        v2 = mha(sd, c + ".cross_d", vis.transpose(0, 1), det.transpose(0, 1), det.transpose(0, 1), det_mask).transpose(0, 1)
        vis = _ln(vis + v2, sd, c + ".norm_d")
    ff = _lin(F.relu(_lin(vis, sd, c + ".ffn_vl.0")), sd, c + ".ffn_vl.3")
    vis = _ln(vis + ff, sd, c + ".norm_vl2")
    return vis, text


def decoder_layer(sd, pre, query, vis, text, query_pos6, text_mask, det, det_mask):
    """BiDecoderLayer.forward (encoder_decoder_layers.py:340-406), loc_learned pos-embed."""
    qp = pos_embed_learned(sd, pre + ".self_posembed", query_pos6).transpose(1, 2).contiguous()
    q, qp = query.transpose(0, 1), qp.transpose(0, 1)
    q = _ln(q + mha(sd, pre + ".self_attn", q + qp, q + qp, q, None), sd, pre + ".norm1")
    lt = text.transpose(0, 1)
    q = _ln(q + mha(sd, pre + ".cross_l", q + qp, lt, lt, text_mask), sd, pre + ".norm_l")
    if det is not None:
        dt = det.transpose(0, 1)
        q = _ln(q + mha(sd, pre + ".cross_d", q + qp, dt, dt, det_mask), sd, pre + ".norm_d")
    vt = vis.transpose(0, 1)
    q = _ln(q + mha(sd, pre + ".cross_v", q + qp, vt, vt, None), sd, pre + ".norm_v")
    ff = _lin(F.relu(_lin(q, sd, pre + ".ffn.0")), sd, pre + ".ffn.3")
    q = _ln(q + ff, sd, pre + ".norm2")
    return q.transpose(0, 1).contiguous()


# ----------------------------------------------------------------------------- whole forward
@torch.no_grad()
def forward(sd, inputs, num_queries=256, num_decoder_layers=6, num_encoder_layers=3, butd=True,
            stage_overrides=None):
    """Eval-mode BeaUTyDETR.forward (models/bdetr.py:193-319) from the output of the frozen
    text encoder on: `inputs` = {point_clouds (B,N,3+C), text_hidden (B,L,768) =
    RoBERTa last_hidden_state, text_attention_mask (B,L) 1 = token (HF convention),
    det_boxes (B,D,6), det_bbox_label_mask (B,D) bool, det_class_ids (B,D) int64}.
    `stage_overrides` (tests only) teacher-forces intermediate results, e.g.
    {'backbone': end_points_dict} or {'sample_inds': tensor}."""
    so = stage_overrides or {}
    ep = dict(so["backbone"]) if "backbone" in so else backbone(sd, inputs["point_clouds"])
    ep["seed_inds"], ep["seed_xyz"], ep["seed_features"] = ep["fp2_inds"], ep["fp2_xyz"], ep["fp2_features"]
    # text projector (models/bdetr.py:79-83,168): Linear + LayerNorm(eps=1e-12)
    text = _ln(_lin(inputs["text_hidden"], sd, "text_projector.0"), sd, "text_projector.1", eps=1e-12)
    text_mask = inputs["text_attention_mask"].ne(1).bool()
    ep["text_feats"], ep["text_attention_mask"] = text, text_mask
    xyz, feats = ep["fp2_xyz"], ep["fp2_features"]
    # box stream (:217-225)
    det = det_mask = None
    if butd:
        det_mask = ~inputs["det_bbox_label_mask"]
        cls = _lin(F.embedding(inputs["det_class_ids"], sd["butd_class_embeddings.weight"]), sd, "class_embeddings")
        det = torch.cat([pos_embed_learned(sd, "box_embeddings", inputs["det_boxes"]),
                         cls.transpose(1, 2)], 1).transpose(1, 2).contiguous()
    # cross-encoder (:231-246)
    vis = feats.transpose(1, 2).contiguous()
    pos = pos_embed_learned(sd, "pos_embed", xyz).transpose(1, 2).contiguous()
    for i in range(num_encoder_layers):
        vis, text = encoder_layer(sd, f"cross_encoder.layers.{i}", vis, pos, text, text_mask, det, det_mask)
    feats = vis.transpose(1, 2).contiguous()
    ep["text_memory"], ep["seed_features"] = text, feats
    ep["proj_tokens"] = contrastive_proj(sd, "contrastive_align_projection_text", text)
    # query generation (:177-191)
    h = F.relu(_bn(_conv1(feats, sd, "points_obj_cls.conv1"), sd, "points_obj_cls.bn1"))
    h = F.relu(_bn(_conv1(h, sd, "points_obj_cls.conv2"), sd, "points_obj_cls.bn2"))
    logits = _conv1(h, sd, "points_obj_cls.conv3")
    ep["seeds_obj_cls_logits"] = logits
    if "sample_inds" in so:
        sample_inds = so["sample_inds"]
    else:
        sample_inds = torch.topk(torch.sigmoid(logits).squeeze(1), num_queries)[1].int()
    cluster_xyz = P.gather_points(xyz.transpose(1, 2).contiguous(), sample_inds).transpose(1, 2).contiguous()
    cluster_feat = P.gather_points(feats, sample_inds).contiguous()
    ep["query_points_xyz"], ep["query_points_feature"] = cluster_xyz, cluster_feat
    ep["query_points_sample_inds"] = sample_inds
    query = _conv1(cluster_feat, sd, "decoder_query_proj").transpose(1, 2).contiguous()
    ep["proposal_proj_queries"] = contrastive_proj(sd, "contrastive_align_projection_image", query)
    base_xyz, base_size = predict_head(sd, "proposal_head", cluster_feat, cluster_xyz, ep, "proposal_")
    # decoder (:278-317)
    for i in range(num_decoder_layers):
        prefix = "last_" if i == num_decoder_layers - 1 else f"{i}head_"
        query_pos = torch.cat([base_xyz, base_size], -1)
        query = decoder_layer(sd, f"decoder.{i}", query, vis, text, query_pos, text_mask, det, det_mask)
        ep[prefix + "proj_queries"] = contrastive_proj(sd, "contrastive_align_projection_image", query)
        base_xyz, base_size = predict_head(sd, f"prediction_heads.{i}", query.transpose(1, 2).contiguous(),
                                           cluster_xyz, ep, prefix)
    return ep
