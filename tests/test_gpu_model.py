"""GPU end-to-end parity of the B200 `BeaUTyDETR` forward against golden vectors produced by the
UNMODIFIED reference model (tests/golden/make_model_golden.py) — tolerance 1e-3 (fp32 gate of
BASELINE.json north_star) — plus the module-surface contract."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = {  # must mirror tests/golden/make_model_golden.py
    "c1": dict(n_points=4096, num_queries=32, n_tokens=16, n_boxes=32, enc=1, dec=1, batch=2, seed=11),
    "c2": dict(n_points=50000, num_queries=256, n_tokens=80, n_boxes=132, enc=3, dec=6, batch=1, seed=12),
}
TOL = 1e-3


def build(name, cuda_lib, precision="fp32"):
    from butd_detr_b200 import BeaUTyDETR, synth
    c = CFG[name]
    model = BeaUTyDETR(num_queries=c["num_queries"], num_decoder_layers=c["dec"], num_encoder_layers=c["enc"],
                       text_encoder=None, precision=precision)
    synth.fill_state_dict_(model.state_dict(), 0)
    model = model.cuda().eval()
    inputs = synth.synth_batch(c["seed"], c["batch"], c["n_points"], c["n_tokens"], c["n_boxes"])
    return model, inputs


def checksum(inputs):
    crc = 0
    for k in sorted(inputs):
        crc = zlib.crc32(inputs[k].contiguous().numpy().tobytes(), crc)
    return crc


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_forward_matches_reference_golden(name, cuda_lib, golden_dir):
    gold = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    model, inputs = build(name, cuda_lib)
    assert checksum(inputs) == int(gold["__input_crc32"]), "synthetic input generator drifted"
    ep = model({k: v.cuda() for k, v in inputs.items()})
    torch.cuda.synchronize()
    # integer outputs: bit exact
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        if k in gold.files:
            assert np.array_equal(ep[k].cpu().numpy(), gold[k]), k
    # the top-k ordering can legitimately flip on ~1e-7 score differences; compare as sets and
    # teacher-force the reference's selection for the per-query tensors if they differ
    want_inds = gold["query_points_sample_inds"]
    got_inds = ep["query_points_sample_inds"].cpu().numpy()
    if not np.array_equal(got_inds, want_inds):
        agree = np.mean([len(set(a) & set(b)) / len(a) for a, b in zip(got_inds, want_inds)])
        assert agree > 0.98, f"query selection agreement {agree}"
        ep = model({k: v.cuda() for k, v in inputs.items()},
                   overrides={"sample_inds": torch.from_numpy(want_inds)})
    worst = {}
    for k in gold.files:
        if k.startswith("__") or gold[k].dtype.kind in "iub":
            continue
        got = ep[k].float().cpu().numpy()
        assert got.shape == gold[k].shape, (k, got.shape, gold[k].shape)
        worst[k] = float(np.abs(got - gold[k]).max())
    bad = {k: v for k, v in worst.items() if not v <= TOL}
    print(name, "max abs err over", len(worst), "tensors:", max(worst.values()))
    assert not bad, bad
    assert torch.equal(ep["text_attention_mask"].cpu(), inputs["text_attention_mask"].ne(1))


def test_module_surface(cuda_lib, golden_dir):
    from butd_detr_b200 import BeaUTyDETR
    model = BeaUTyDETR(text_encoder=None)
    spec = json.load(open(os.path.join(golden_dir, "state_dict_spec.json")))
    sd = model.state_dict()
    assert set(sd) == set(spec["tensors"])
    for k, (shape, dtype) in spec["tensors"].items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    model.eval()
    with pytest.raises(RuntimeError, match="CPU not supported"):
        model({"point_clouds": torch.zeros(1, 1024, 6), "text_hidden": torch.zeros(1, 4, 768),
               "text_attention_mask": torch.ones(1, 4, dtype=torch.long)})
    model.train()  # the training forward exists (tests/test_gpu_train.py); the attention-only entry is eval only
    with pytest.raises(NotImplementedError):
        model({"seed_features": torch.zeros(1, 288, 1024).cuda(), "text_hidden": torch.zeros(1, 4, 768).cuda(),
               "text_attention_mask": torch.ones(1, 4, dtype=torch.long).cuda()})


def test_batch_rows_are_independent(cuda_lib):
    """Scenes never interact in the forward: a batch equals its scenes run alone."""
    from butd_detr_b200 import BeaUTyDETR, synth
    model = BeaUTyDETR(num_queries=32, num_decoder_layers=1, num_encoder_layers=1, text_encoder=None)
    synth.fill_state_dict_(model.state_dict(), 0)
    model = model.cuda().eval()
    inputs = {k: v.cuda() for k, v in synth.synth_batch(3, 3, 4096, 16, 32, ragged_text=False).items()}
    full = model(inputs)
    one = model({k: v[1:2] for k, v in inputs.items()})
    for k in ("last_center", "last_pred_size", "last_sem_cls_scores", "seeds_obj_cls_logits"):
        torch.testing.assert_close(full[k][1:2], one[k], rtol=0, atol=1e-5)
    assert torch.equal(full["sa1_inds"][1:2], one["sa1_inds"])


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("fp16", 1e-2)])
@pytest.mark.parametrize("name", ["c1", "c2"])
def test_tensor_core_forward_matches_reference_golden(name, precision, tol, cuda_lib, golden_dir):
    """Tensor-core modes (BASELINE.json configs[1]).  'bf16x3' (bf16 hi/lo split operands, three
    tcgen05 MMAs per product, fp32 accumulation) is the shipped mode and must meet even the fp32
    gate (1e-3).  'fp16' (fp16 operands, one MMA per product) must meet the reduced-precision gate
    (1e-2); plain bf16 operands did not (1.2-2.2e-2 after ~30 stacked post-LN GEMM layers), which
    is why the single-pass mode uses fp16's 11 significant bits.  Query selection is
    teacher-forced with the reference's indices (SURVEY.md §7 hard part 3: 1e-2 perturbations
    reorder the near-tied top-k scores, which permutes per-query outputs without changing their
    values); the selection agreement is reported."""
    gold = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    model, inputs = build(name, cuda_lib, precision=precision)
    dev_in = {k: v.cuda() for k, v in inputs.items()}
    free = model(dev_in)
    want_inds = gold["query_points_sample_inds"]
    got_inds = free["query_points_sample_inds"].cpu().numpy()
    agree = np.mean([len(set(a) & set(b)) / len(a) for a, b in zip(got_inds, want_inds)])
    for k in ("sa1_inds", "sa2_inds"):
        assert np.array_equal(free[k].cpu().numpy(), gold[k]), k      # point ops stay exact
    ep = model(dev_in, overrides={"sample_inds": torch.from_numpy(want_inds)})
    worst = {}
    for k in gold.files:
        if k.startswith("__") or gold[k].dtype.kind in "iub":
            continue
        worst[k] = float(np.abs(ep[k].float().cpu().numpy() - gold[k]).max())
    graded = {k: v for k, v in worst.items() if k.endswith(("center", "pred_size", "sem_cls_scores", "proj_queries"))}
    print(name, precision, ": top-k agreement %.3f, max abs err graded %.2e, all %.2e" % (
        agree, max(graded.values()), max(worst.values())))
    bad = {k: v for k, v in graded.items() if not v <= tol}
    assert not bad, bad
    assert agree > (0.98 if precision == "bf16x3" else 0.8)
