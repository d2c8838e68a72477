"""The outer drop-in boundary, exercised with the reference's OWN code on the GPU box
(SURVEY.md §8b): the unmodified reference Python (git-ignored copy under baseline/_ref, made by
baseline/install_ref.py) is imported and
  * its `BeaUTyDETR` — torch CUDA layers + its PointNet++ modules running on THIS package's
    `pointnet2._ext` drop-in — loads the same state_dict (strict) and must produce the same
    `end_points` as the B200 engine;
  * its loss (`models/losses.py:546-617`, Hungarian matcher, contrastive alignment reading
    `end_points['tokenized']`) and its `GroundingEvaluator` (`src/grounding_evaluator.py:99-242`)
    consume this package's `end_points` unchanged, after the batch keys are merged in the way
    `main_utils.py:421-426` does (collision assert included), and report what they report on the
    reference model's own output;
  * the module survives the DistributedDataParallel wrap of `main_utils.py:310-313`."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(n_points=4096, num_queries=32, n_tokens=16, n_boxes=32, dec=2, batch=2, seed=17)


@pytest.fixture(scope="module")
def ref(cuda_lib):
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference Python not installed (python baseline/install_ref.py in the build container)")
    from butd_detr_b200 import pointnet2_ext
    return ref_loader, ref_loader.import_reference(ext=pointnet2_ext)


@pytest.fixture(scope="module")
def process_group():
    """One-rank NCCL group: the reference loss all-reduces `num_boxes` (models/losses.py:532-534) and
    DDP needs one."""
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield dist
    if created:
        dist.destroy_process_group()


@pytest.fixture(scope="module")
def pair(ref):
    """(our model, reference model with the same weights, device inputs, our end_points, reference end_points)."""
    from butd_detr_b200 import BeaUTyDETR, pointnet2_ext, synth
    ref_loader, _ = ref
    # the reference layers must compute in fp32 like the reference's pinned torch 1.10 did for matmuls;
    # cuDNN's TF32 convolutions (on by default) would cost it 1e-3
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ours = BeaUTyDETR(num_queries=CFG["num_queries"], num_decoder_layers=CFG["dec"], text_encoder=None)
    sd = synth.fill_state_dict_(ours.state_dict(), 0)
    ours = ours.cuda().eval()
    theirs = ref_loader.build_reference_model(ext=pointnet2_ext, num_queries=CFG["num_queries"],
                                              num_decoder_layers=CFG["dec"])
    missing = theirs.load_state_dict({k: v for k, v in sd.items()}, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("text_encoder.") for k in missing.missing_keys), missing
    theirs = theirs.cuda().eval()
    inputs = {k: v.cuda() for k, v in synth.synth_batch(CFG["seed"], CFG["batch"], CFG["n_points"], CFG["n_tokens"],
                                                        CFG["n_boxes"]).items()}
    with torch.no_grad():
        ep_ours = ours(inputs)
        ep_ref = ref_loader.run_reference(theirs, inputs)
    torch.cuda.synchronize()
    return ours, theirs, inputs, ep_ours, ep_ref


def test_reference_model_on_the_drop_in_ext_equals_the_engine(pair):
    """Same weights, same inputs: reference modules (cuDNN / cuBLAS + our nine point ops) vs the engine."""
    _, _, _, ep_ours, ep_ref = pair
    worst = {}
    for k, w in ep_ref.items():
        if not torch.is_tensor(w):
            continue
        assert k in ep_ours, f"end_points['{k}'] missing"
        g = ep_ours[k]
        assert tuple(g.shape) == tuple(w.shape), (k, g.shape, w.shape)
        if w.dtype.is_floating_point:
            worst[k] = float((g.float() - w.float()).abs().max())
        else:
            assert torch.equal(g.to(w.dtype), w), k
    print("engine vs reference model on the drop-in ext: max abs err", max(worst.values()), "over", len(worst), "tensors")
    bad = {k: v for k, v in worst.items() if not v <= 1e-3}
    assert not bad, bad
    assert ep_ours["tokenized"]["attention_mask"].shape == ep_ref["tokenized"]["attention_mask"].shape


def _ground_truth(inputs, n_points, G=132):
    """Synthetic annotations with the schema of Joint3DDataset.__getitem__ (joint_det_dataset.py:738-790)."""
    B = inputs["point_clouds"].shape[0]
    g = torch.Generator().manual_seed(5)
    n_obj = 3
    gt = {
        "center_label": torch.zeros(B, G, 3), "size_gts": torch.zeros(B, G, 3),
        "sem_cls_label": torch.zeros(B, G, dtype=torch.int64), "box_label_mask": torch.zeros(B, G),
        "positive_map": torch.zeros(B, G, 256), "point_instance_label": -torch.ones(B, n_points, dtype=torch.int64),
        "is_view_dep": torch.zeros(B, dtype=torch.bool), "is_hard": torch.zeros(B, dtype=torch.bool),
        "is_unique": torch.ones(B, dtype=torch.bool),
    }
    pc = inputs["point_clouds"].cpu()
    for b in range(B):
        centres = pc[b, torch.randint(0, n_points, (n_obj,), generator=g), :3]
        sizes = torch.rand(n_obj, 3, generator=g) * 0.8 + 0.4
        gt["center_label"][b, :n_obj], gt["size_gts"][b, :n_obj] = centres, sizes
        gt["sem_cls_label"][b, :n_obj] = torch.randint(0, 18, (n_obj,), generator=g)
        gt["box_label_mask"][b, :n_obj] = 1
        for o in range(n_obj):
            gt["positive_map"][b, o, 1 + 2 * o:3 + 2 * o] = 0.5  # two tokens mention the object
            inside = ((pc[b, :, :3] - centres[o]).abs() < sizes[o] / 2).all(-1)
            gt["point_instance_label"][b, inside] = o
    return {k: v.cuda() for k, v in gt.items()}


def test_reference_loss_and_evaluator_consume_end_points_unchanged(pair, ref, process_group):
    _, models = ref
    import sys
    from src.grounding_evaluator import GroundingEvaluator
    _, _, inputs, ep_ours, ep_ref = pair
    gt = _ground_truth(inputs, CFG["n_points"])
    set_criterion = models.SetCriterion(matcher=models.HungarianMatcher(1, 0, 2, True),
                                        losses=["boxes", "labels", "contrastive_align"], eos_coef=0.1,
                                        temperature=0.07).cuda()  # main_utils.py:241-253
    prefixes = ["last_", "proposal_"] + [f"{i}head_" for i in range(CFG["dec"] - 1)]
    results = []
    for ep in (dict(ep_ours), dict(ep_ref)):
        for key in gt:  # main_utils.py:424-426
            assert key not in ep
            ep[key] = gt[key]
        with torch.no_grad():
            loss, ep = models.compute_hungarian_loss(ep, CFG["dec"], set_criterion, query_points_obj_topk=4)
        for key in list(ep):
            if "pred_size" in key:
                ep[key] = torch.clamp(ep[key], min=1e-6)  # main_utils.py:486-488
        ev = GroundingEvaluator(only_root=True, thresholds=[0.25, 0.5], topks=[1, 5, 10], prefixes=prefixes)
        for prefix in prefixes:
            ev.evaluate(ep, prefix)
        results.append((float(loss), {k: float(ep[k]) for k in ("loss_ce", "loss_bbox", "loss_giou",
                                                                 "loss_constrastive_align",
                                                                 "query_points_generation_loss")}, dict(ev.dets), dict(ev.gts)))
    (l0, parts0, dets0, gts0), (l1, parts1, dets1, gts1) = results
    print("loss on our end_points", l0, "on the reference model's", l1)
    assert np.isfinite(l0) and abs(l0 - l1) <= 1e-3 * max(1.0, abs(l1)), (l0, l1)
    for k in parts0:
        assert abs(parts0[k] - parts1[k]) <= 1e-3 * max(1.0, abs(parts1[k])), (k, parts0[k], parts1[k])
    assert gts0 == gts1 and dets0 == dets1


def test_ddp_wrapped_eval_forward(pair, process_group):
    """DistributedDataParallel(model, device_ids=[gpu], broadcast_buffers=False) as in
    main_utils.py:310-313 (one process here), then the eval forward through the wrapper."""
    ours, _, inputs, ep_ours, _ = pair
    ddp = torch.nn.parallel.DistributedDataParallel(ours, device_ids=[0], broadcast_buffers=False,
                                                    find_unused_parameters=True)
    ddp.eval()
    with torch.no_grad():
        ep = ddp(inputs)
    for k in ("last_center", "last_sem_cls_scores", "sa1_inds", "proj_tokens"):
        assert torch.equal(ep[k], ep_ours[k]), k
