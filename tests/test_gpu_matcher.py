"""Device-side Hungarian matcher against the reference's algorithm: the cost formulas of
`/root/reference/models/losses.py:28-91,296-313` restated in torch, scipy.optimize.linear_sum_assignment (the
third-party solver the reference calls, losses.py:316-319) as the assignment oracle, and — when the reference Python
is installed under baseline/_ref — the reference's own `HungarianMatcher` / `SetCriterion`."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _corners(x):  # losses.py:28-39
    c, s = x[..., :3], torch.clamp(x[..., 3:], min=1e-6)
    return torch.cat([c - 0.5 * s, c + 0.5 * s], -1)


def _giou(a, b):  # losses.py:42-91
    lt, rb = torch.max(a[:, None, :3], b[None, :, :3]), torch.min(a[:, None, 3:], b[None, :, 3:])
    inter = (rb - lt).clamp(min=0).prod(-1)
    va, vb = (a[:, 3:] - a[:, :3]).prod(-1), (b[:, 3:] - b[:, :3]).prod(-1)
    union = va[:, None] + vb[None] - inter
    wh = (torch.max(a[:, None, 3:], b[None, :, 3:]) - torch.min(a[:, None, :3], b[None, :, :3])).clamp(min=0)
    vol = wh.prod(-1)
    return inter / union - (vol - union) / vol


def _reference_cost(logits, boxes, tgt, w, soft):  # losses.py:283-313, one scene, (Q, T)
    prob = logits.softmax(-1)
    if soft:
        cc = -prob @ tgt["positive_map"][:, :prob.shape[-1]].t()
    else:
        cc = -prob[:, tgt["labels"]]
    return w[1] * torch.cdist(boxes, tgt["boxes"], p=1) + w[0] * cc + w[2] * -_giou(_corners(boxes), _corners(tgt["boxes"]))


def _scene(B, Q, C, sizes, seed):
    g = torch.Generator().manual_seed(seed)
    out = {"pred_logits": torch.randn(B, Q, C, generator=g) * 2,
           "pred_boxes": torch.cat([torch.rand(B, Q, 3, generator=g) * 4 - 2, torch.rand(B, Q, 3, generator=g) * 1.5 + 0.05], -1)}
    targets = []
    for n in sizes:
        pm = torch.zeros(n, 256)
        for t in range(n):
            s = int(torch.randint(0, C - 4, (1,), generator=g))
            k = int(torch.randint(1, 4, (1,), generator=g))
            pm[t, s:s + k] = 1.0 / k
        targets.append({"labels": torch.randint(0, C, (n,), generator=g),
                        "boxes": torch.cat([torch.rand(n, 3, generator=g) * 4 - 2, torch.rand(n, 3, generator=g) * 1.5 + 0.05], -1),
                        "positive_map": pm})
    return out, targets


@pytest.mark.parametrize("soft", [True, False])
@pytest.mark.parametrize("B,Q,C,sizes,w", [
    (4, 256, 256, [3, 132, 0, 17], (1, 0, 2)),      # the reference's training setting (main_utils.py:241-253)
    (2, 256, 256, [256, 1], (1, 5, 2)),             # as many targets as queries
    (3, 32, 19, [5, 32, 2], (2, 5, 1)),             # narrow logits, Q not the CTA width
    (2, 100, 256, [40, 7], (1, 5, 2)),              # Q not a multiple of 32
])
def test_matcher_equals_scipy_on_the_reference_costs(cuda_lib, B, Q, C, sizes, w, soft):
    from scipy.optimize import linear_sum_assignment
    from butd_detr_b200.matcher import HungarianMatcher
    out, targets = _scene(B, Q, C, sizes, 11 + Q + C)
    out_d = {k: v.cuda() for k, v in out.items()}
    tg_d = [{k: v.cuda() for k, v in t.items()} for t in targets]
    m = HungarianMatcher(*w, soft_token=soft)
    got = m(out_d, tg_d)
    torch.cuda.synchronize()
    assert int(m.last_status) == 0
    off = 0
    for b, n in enumerate(sizes):
        iq, it = got[b]
        assert iq.dtype == torch.int64 and iq.is_cuda and len(iq) == len(it) == n
        if n == 0:
            continue
        want_c = _reference_cost(out["pred_logits"][b], out["pred_boxes"][b], targets[b], w, soft)  # (Q, n) fp32, CPU
        ours_c = m.last_cost[off:off + n].t().cpu()
        off += n
        assert (ours_c - want_c).abs().max().item() <= 2e-5, "cost matrix differs from the reference formulas"
        # the solver: scipy on OUR cost matrix must return the same pairs (continuous costs: unique optimum)
        ri, ci = linear_sum_assignment(ours_c.numpy())
        assert np.array_equal(iq.cpu().numpy(), ri) and np.array_equal(it.cpu().numpy(), ci)
        # ... and on the reference-formula matrix the same total cost (the two matrices differ by rounding only)
        r2, c2 = linear_sum_assignment(want_c.numpy())
        assert abs(float(want_c[r2, c2].sum()) - float(want_c[iq.cpu(), it.cpu()].sum())) <= 1e-3


def test_matcher_flags_scenes_without_a_finite_assignment(cuda_lib):
    from butd_detr_b200.matcher import HungarianMatcher
    out, targets = _scene(2, 32, 19, [4, 3], 5)
    out["pred_boxes"][1] = float("nan")
    m = HungarianMatcher(1, 5, 2, soft_token=True)
    got = m({k: v.cuda() for k, v in out.items()}, [{k: v.cuda() for k, v in t.items()} for t in targets])
    torch.cuda.synchronize()
    assert int(m.last_status) == 1 and bool((got[1][0] == -1).all()) and bool((got[0][0] >= 0).all())


def test_reference_criterion_with_the_device_matcher(cuda_lib):
    """The reference's own SetCriterion (losses.py:333-543) driven by the device matcher gives the losses it gives
    with its own matcher (scipy on the host)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference Python not installed (python baseline/install_ref.py in the build container)")
    from butd_detr_b200 import pointnet2_ext
    from butd_detr_b200.matcher import HungarianMatcher
    _, models = ref_loader, ref_loader.import_reference(ext=pointnet2_ext)
    import torch.distributed as dist
    import os
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        B, Q, L = 3, 256, 20
        out, targets = _scene(B, Q, 256, [4, 31, 9], 23)
        g = torch.Generator().manual_seed(3)
        out["proj_tokens"] = torch.nn.functional.normalize(torch.randn(B, L, 64, generator=g), dim=-1)
        out["proj_queries"] = torch.nn.functional.normalize(torch.randn(B, Q, 64, generator=g), dim=-1)
        out_d = {k: v.cuda() for k, v in out.items()}
        out_d["tokenized"] = {"attention_mask": torch.ones(B, L, dtype=torch.long, device="cuda")}
        tg_d = [{k: v.cuda() for k, v in t.items()} for t in targets]
        res = []
        for matcher in (models.HungarianMatcher(1, 0, 2, True), HungarianMatcher(1, 0, 2, True)):
            crit = models.SetCriterion(matcher=matcher, losses=["boxes", "labels", "contrastive_align"], eos_coef=0.1,
                                       temperature=0.07).cuda()
            losses, indices = crit(out_d, tg_d)
            res.append(({k: float(v) for k, v in losses.items()}, [(i.cpu(), j.cpu()) for i, j in indices]))
        (l_ref, i_ref), (l_dev, i_dev) = res
        for (a, b_), (c, d) in zip(i_ref, i_dev):
            assert torch.equal(a, c) and torch.equal(b_, d)
        for k in l_ref:
            assert abs(l_ref[k] - l_dev[k]) <= 1e-5 * max(1.0, abs(l_ref[k])), (k, l_ref[k], l_dev[k])
    finally:
        if created:
            dist.destroy_process_group()
