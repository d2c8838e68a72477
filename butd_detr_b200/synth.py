"""Deterministic synthetic inputs and weights (numpy PCG64; identical on every machine).

There is no ScanNet / ReferIt3D data, no checkpoint and no RoBERTa weights offline, so the
bench, the tests and the golden-vector generator all draw from here:

* `synth_scene`   — a "ScanNet-shaped" room scan with the schema `Joint3DDataset.__getitem__`
  feeds the model (`/root/reference/src/joint_det_dataset.py:738-790`,
  `train_dist_mod.py:102-110`): `point_clouds (N,6)` = xyz + (rgb - mean_rgb)
  (`joint_det_dataset.py:68,415`), detected boxes / class ids / mask (`MAX_NUM_OBJ`, `:33`).
* `synth_text`    — stands in for RoBERTa's `last_hidden_state` + HF attention mask
  (`models/bdetr.py:164-171`).
* `fill_state_dict_` — per-key seeded values for every tensor of a `BeaUTyDETR`
  state_dict (keys as in the reference), with non-trivial BatchNorm running statistics so
  that eval-mode BN folding is exercised.
"""
import zlib

import numpy as np
import torch

MEAN_RGB = np.array([109.8, 97.2, 83.8], dtype=np.float64) / 256.0  # joint_det_dataset.py:68


def _rng(seed, tag):
    return np.random.Generator(np.random.PCG64([int(seed) & 0x7FFFFFFF, zlib.crc32(tag.encode())]))


def synth_scene(seed, n_points=50000, n_boxes=132, dup_frac=0.01):
    """One scene. Returns dict of numpy arrays:
    point_clouds (N,6) f32, det_boxes (D,6) f32, det_class_ids (D,) i64, det_bbox_label_mask (D,) bool."""
    g = _rng(seed, "scene")
    W, Dp, H = 6.0, 5.0, 2.7  # room centred at the origin, floor at z = 0
    n_obj = int(g.integers(10, 21))
    obj_c = np.stack([g.uniform(-W / 2 + 0.8, W / 2 - 0.8, n_obj), g.uniform(-Dp / 2 + 0.8, Dp / 2 - 0.8, n_obj),
                      np.zeros(n_obj)], 1)
    obj_s = g.uniform(0.3, 1.5, (n_obj, 3))
    obj_c[:, 2] = obj_s[:, 2] / 2
    # surfaces: floor, 4 walls, then the 6 faces of each box; sample proportional to area
    areas = [W * Dp, W * H, W * H, Dp * H, Dp * H]
    for s in obj_s:
        areas += [s[0] * s[1]] * 2 + [s[0] * s[2]] * 2 + [s[1] * s[2]] * 2
    areas = np.asarray(areas)
    face = g.choice(len(areas), size=n_points, p=areas / areas.sum())
    u, v = g.uniform(-0.5, 0.5, n_points), g.uniform(-0.5, 0.5, n_points)
    xyz = np.zeros((n_points, 3))
    for f in range(5):
        m = face == f
        if f == 0:
            xyz[m] = np.stack([u[m] * W, v[m] * Dp, np.zeros(m.sum())], 1)
        elif f in (1, 2):
            xyz[m] = np.stack([u[m] * W, np.full(m.sum(), (Dp / 2) * (1 if f == 1 else -1)), (v[m] + 0.5) * H], 1)
        else:
            xyz[m] = np.stack([np.full(m.sum(), (W / 2) * (1 if f == 3 else -1)), u[m] * Dp, (v[m] + 0.5) * H], 1)
    for o in range(n_obj):
        c, s = obj_c[o], obj_s[o]
        for k in range(6):
            m = face == 5 + 6 * o + k
            if not m.any():
                continue
            ax = k // 2  # fixed axis: 0 -> z faces, 1 -> y faces, 2 -> x faces
            sign = 1.0 if k % 2 == 0 else -1.0
            p = np.zeros((m.sum(), 3))
            if ax == 0:
                p[:, 0], p[:, 1], p[:, 2] = u[m] * s[0], v[m] * s[1], sign * s[2] / 2
            elif ax == 1:
                p[:, 0], p[:, 2], p[:, 1] = u[m] * s[0], v[m] * s[2], sign * s[1] / 2
            else:
                p[:, 1], p[:, 2], p[:, 0] = u[m] * s[1], v[m] * s[2], sign * s[0] / 2
            xyz[m] = p + c
    xyz += g.normal(0, 0.004, xyz.shape)  # sensor noise
    rgb = g.uniform(0, 1, (n_points, 3)) - MEAN_RGB
    pc = np.concatenate([xyz, rgb], 1).astype(np.float32)
    # ScanNet scans with < 50k vertices are re-sampled WITH replacement
    # (visual_data_handlers.py:114-118): exact duplicate points (and exact FPS ties) do occur.
    n_dup = int(n_points * dup_frac)
    if n_dup > 0:
        dst = g.choice(n_points, n_dup, replace=False)
        src = g.choice(n_points, n_dup, replace=True)
        pc[dst] = pc[src]
    pc = pc[g.permutation(n_points)]
    # detected boxes: first n_valid are real (centre, size), rest zero-padded (joint_det_dataset.py:626-700)
    D = n_boxes
    n_valid = int(g.integers(min(16, D), D + 1))
    boxes = np.zeros((D, 6), np.float32)
    boxes[:n_valid, 0] = g.uniform(-W / 2, W / 2, n_valid)
    boxes[:n_valid, 1] = g.uniform(-Dp / 2, Dp / 2, n_valid)
    boxes[:n_valid, 2] = g.uniform(0, H, n_valid)
    boxes[:n_valid, 3:] = g.uniform(0.2, 1.5, (n_valid, 3))
    cls = np.zeros(D, np.int64)
    cls[:n_valid] = g.integers(0, 485, n_valid)
    mask = np.zeros(D, bool)
    mask[:n_valid] = True
    return {"point_clouds": pc, "det_boxes": boxes, "det_class_ids": cls, "det_bbox_label_mask": mask}


def synth_text(seed, n_tokens=80, ragged=True, hidden=768):
    """Stand-in for RoBERTa output: (L,768) f32 hidden states and (L,) int64 HF mask (1 = token)."""
    g = _rng(seed, "text")
    h = g.normal(0, 1, (n_tokens, hidden)).astype(np.float32)
    n_valid = int(g.integers(max(2, n_tokens // 4), n_tokens + 1)) if ragged else n_tokens
    m = np.zeros(n_tokens, np.int64)
    m[:n_valid] = 1
    return {"text_hidden": h, "text_attention_mask": m}


def synth_batch(seed, batch, n_points=50000, n_tokens=80, n_boxes=132, ragged_text=True, device="cpu"):
    """Batch of scenes as torch tensors keyed like the reference's model inputs
    (+ `text_hidden` / `text_attention_mask` instead of raw strings; at least one scene
    uses the full token length, mirroring `padding="longest"`, models/bdetr.py:164-166)."""
    scenes = [synth_scene(seed * 1000 + b, n_points, n_boxes) for b in range(batch)]
    texts = [synth_text(seed * 1000 + b, n_tokens, ragged=ragged_text and b > 0) for b in range(batch)]
    out = {k: torch.from_numpy(np.stack([s[k] for s in scenes])) for k in scenes[0]}
    out.update({k: torch.from_numpy(np.stack([t[k] for t in texts])) for k in texts[0]})
    return {k: v.to(device) for k, v in out.items()}


def synth_value(name, shape, dtype=torch.float32, seed=0):
    """Seeded value for ONE state_dict entry, chosen from its name/shape only."""
    g = _rng(seed, name)
    shape = tuple(shape)
    if name.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.int64)
    if name.endswith("running_var"):
        a = g.uniform(0.5, 1.5, shape)
    elif name.endswith("running_mean"):
        a = g.normal(0, 0.1, shape)
    elif name.startswith("butd_class_embeddings"):
        a = g.normal(0.019, 0.40, shape)  # statistics of data/class_embeddings3d.npy
    elif len(shape) <= 1:
        a = g.uniform(0.5, 1.5, shape) if name.endswith("weight") else g.normal(0, 0.05, shape)
    else:
        fan_in = int(np.prod(shape[1:]))
        gain = 2.0 if len(shape) == 4 else 1.0  # SharedMLP convs feed ReLUs
        a = g.normal(0, np.sqrt(gain / fan_in), shape)
    return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(dtype)


def fill_state_dict_(sd, seed=0, skip_prefixes=("text_encoder.",)):
    """In-place seeded fill of every tensor in `sd` (name -> tensor), except skipped prefixes."""
    for name, t in sd.items():
        if any(name.startswith(p) for p in skip_prefixes):
            continue
        t.copy_(synth_value(name, t.shape, t.dtype, seed))
    return sd
