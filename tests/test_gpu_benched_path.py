"""Parity of the configuration bench.py actually times: CUDA-graph replay (multi-stream capture),
large batches (the B > 8 FPS variants, multi-wave kernels), fp16 / bf16x3 tensor-core operands —
against the eager path, the per-scene runs and the reference's golden outputs.  Plus configs[3]
(attention-only entry) against the oracle port and the fully-masked row on the tcgen05 attention."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GRADED = ("center", "pred_size", "sem_cls_scores", "proj_queries")
C2 = dict(n_points=50000, num_queries=256, n_tokens=80, n_boxes=132, enc=3, dec=6, seed=12)  # tests/golden/make_model_golden.py


def _model(precision, cuda_graph, num_queries=256, enc=3, dec=6):
    from butd_detr_b200 import BeaUTyDETR, synth
    model = BeaUTyDETR(num_queries=num_queries, num_decoder_layers=dec, num_encoder_layers=enc, text_encoder=None,
                       precision=precision, cuda_graph=cuda_graph)
    synth.fill_state_dict_(model.state_dict(), 0)
    return model.cuda().eval()


def _assert_same(a, b, what, atol=1e-5):
    """Two end_points dicts: integer / bool tensors bit-equal, floats within atol."""
    assert set(a) == set(b), what
    for k in sorted(a):
        x, y = a[k], b[k]
        if not torch.is_tensor(x):
            continue
        assert x.shape == y.shape, (what, k, x.shape, y.shape)
        if x.dtype.is_floating_point:
            err = float((x.float() - y.float()).abs().max())
            assert err <= atol, f"{what}: {k} differs by {err}"
        else:
            assert torch.equal(x, y), f"{what}: {k} differs"


def _snapshot(ep):
    return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in ep.items()}


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16x3"])
def test_graph_replay_equals_eager(cuda_lib, precision):
    """forward_graphed (one CUDA graph capturing six streams with PDL edges, static buffers) returns
    what the eager launch sequence returns — on the capture inputs, on a replay with other inputs of
    the same shapes, and again on the first inputs (no state leaks between replays)."""
    from butd_detr_b200 import synth
    eager = _model(precision, False, 32, 1, 1)
    graphed = _model(precision, True, 32, 1, 1)
    a = {k: v.cuda() for k, v in synth.synth_batch(31, 3, 4096, 16, 32).items()}
    b = {k: v.cuda() for k, v in synth.synth_batch(32, 3, 4096, 16, 32).items()}
    want_a, want_b = _snapshot(eager(a)), _snapshot(eager(b))
    got_a = _snapshot(graphed(a))
    got_b = _snapshot(graphed(b))
    got_a2 = _snapshot(graphed(a))
    torch.cuda.synchronize()
    _assert_same(got_a, want_a, f"{precision} capture inputs", 1e-6)
    _assert_same(got_b, want_b, f"{precision} replay with new inputs", 1e-6)
    _assert_same(got_a2, want_a, f"{precision} replay of the first inputs", 1e-6)
    assert len(graphed.engine()._graphs) == 1


@pytest.mark.parametrize("precision,B", [("fp16", 16), ("bf16x3", 16), ("fp16", 128)])
def test_large_batch_graph_forward_matches_golden_and_per_scene_runs(cuda_lib, golden_dir, precision, B):
    """The benched path itself (configs[1], CUDA graph, B scenes, tensor-core operands): scene 0 is the
    input of the reference's golden run (model_c2.npz) — its point-op indices must equal the
    reference's and its query-independent outputs must meet the precision's gate; EVERY scene must
    equal its own eager single-scene run (bit-equal indices, 1e-5 on floats: scenes never interact,
    SURVEY.md §8e)."""
    from butd_detr_b200 import synth
    gold = np.load(os.path.join(golden_dir, "model_c2.npz"))
    first = synth.synth_batch(C2["seed"], 1, C2["n_points"], C2["n_tokens"], C2["n_boxes"])
    rest = synth.synth_batch(4242, B - 1, C2["n_points"], C2["n_tokens"], C2["n_boxes"])
    batch = {k: torch.cat([first[k], rest[k]]).cuda() for k in first}
    graphed = _model(precision, True)
    eager = _model(precision, False)
    ep = graphed(batch)
    torch.cuda.synchronize()
    # scene 0 against the reference's own outputs
    for k in ("sa1_inds", "sa2_inds", "fp2_inds"):
        assert np.array_equal(ep[k][0:1].cpu().numpy(), gold[k]), k
    gate = 1e-2 if precision == "fp16" else 1e-3
    for k in ("proj_tokens", "seeds_obj_cls_logits", "text_memory"):
        err = float(np.abs(ep[k][0:1].float().cpu().numpy() - gold[k]).max())
        print(f"B={B} {precision} scene 0 {k}: max abs err vs reference {err:.2e}")
        if k == "proj_tokens":
            assert err <= gate, (k, err)
    want_inds = gold["query_points_sample_inds"][0]
    got_inds = ep["query_points_sample_inds"][0].cpu().numpy()
    agree = len(set(got_inds) & set(want_inds)) / len(want_inds)
    print(f"B={B} {precision}: top-k agreement of scene 0 with the reference {agree:.3f}")
    if np.array_equal(got_inds, want_inds):  # same queries in the same order: the graded tensors compare directly
        for pf in ["proposal_"] + [f"{i}head_" for i in range(5)] + ["last_"]:
            for g in GRADED:
                err = float(np.abs(ep[pf + g][0:1].float().cpu().numpy() - gold[pf + g]).max())
                assert err <= gate, (pf + g, err)
    # every scene against its own eager B = 1 run
    worst = 0.0
    for s in range(B):
        one = eager({k: v[s:s + 1] for k, v in batch.items()})
        for k, v in one.items():
            if not torch.is_tensor(v):
                continue
            w = ep[k][s:s + 1]
            if v.dtype.is_floating_point:
                err = float((v.float() - w.float()).abs().max())
                worst = max(worst, err)
                assert err <= 1e-5, f"scene {s}: {k} differs from its single-scene run by {err}"
            else:
                assert torch.equal(v, w), f"scene {s}: {k} differs from its single-scene run"
    print(f"B={B} {precision}: batch vs per-scene eager runs, max abs diff {worst:.2e}")


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16x3", 1e-3), ("fp16", 1e-2)])
def test_attention_only_entry_matches_oracle(cuda_lib, oracle_lib, precision, tol):
    """BASELINE.json configs[3]: seed features / coordinates supplied through the module's inputs
    (`seed_features (B,288,V)`, `seed_xyz`, `seed_inds`), 256 detected-box tokens, 80 text tokens —
    no FPS / ball query.  Checked against the oracle port entered at the same stage."""
    from butd_detr_b200 import synth
    from oracle import model_ref
    B, V, D = 2, 1024, 256
    model = _model(precision, False)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    inputs = synth.synth_batch(53, B, 2048, 80, D)
    g = torch.Generator().manual_seed(7)
    seed_feats = torch.randn(B, 288, V, generator=g)
    seed_xyz = torch.rand(B, V, 3, generator=g) * torch.tensor([6.0, 5.0, 2.7]) - torch.tensor([3.0, 2.5, 0.0])
    seed_inds = torch.arange(V, dtype=torch.int32)[None].expand(B, -1).contiguous()
    stage = {"fp2_features": seed_feats, "fp2_xyz": seed_xyz, "fp2_inds": seed_inds}
    want = model_ref.forward(sd, inputs, 256, 6, 3, stage_overrides={"backbone": stage})
    dev_in = {k: v.cuda() for k, v in inputs.items() if k != "point_clouds"}
    dev_in.update(seed_features=seed_feats.cuda(), seed_xyz=seed_xyz.cuda(), seed_inds=seed_inds.cuda())
    ep = model(dev_in)
    assert "sa1_inds" not in ep  # the backbone did not run
    got_inds, want_inds = ep["query_points_sample_inds"].cpu(), want["query_points_sample_inds"]
    agree = np.mean([len(set(a.tolist()) & set(b.tolist())) / len(a) for a, b in zip(got_inds, want_inds)])
    print(f"configs[3] {precision}: top-k agreement {agree:.3f}")
    if not torch.equal(got_inds, want_inds):
        ep = model(dev_in, overrides={"sample_inds": want_inds})
    worst = {}
    for k, w in want.items():
        if not torch.is_tensor(w) or not w.dtype.is_floating_point or k not in ep:
            continue
        worst[k] = float((ep[k].float().cpu() - w).abs().max())
    graded = {k: v for k, v in worst.items() if k.endswith(GRADED) or k in ("proj_tokens", "text_memory")}
    print(f"configs[3] {precision}: max abs err graded {max(graded.values()):.2e}, all {max(worst.values()):.2e}")
    bad = {k: v for k, v in (worst if precision == "fp32" else graded).items() if not v <= tol}
    assert not bad, bad


@pytest.mark.parametrize("split", [1, 3])
def test_attention_tc_fully_masked_row_is_nan_like_reference(cuda_lib, split):
    """softmax over a row whose keys are all padding is NaN in the reference (nn.MultiheadAttention);
    the tcgen05 kernel must say the same, and leave the other scene of the batch untouched."""
    H, hd, E, Lq, Lk = 8, 36, 288, 130, 80
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(2, Lq, E, device="cuda", generator=g)
    kv = torch.randn(2, Lk, 2 * E, device="cuda", generator=g)
    mask = torch.zeros(2, Lk, dtype=torch.uint8, device="cuda")
    mask[0] = 1          # scene 0: every key masked
    mask[1, 50:] = 1
    out = torch.zeros(2, Lq, E, device="cuda")
    lib = cuda_lib.load()
    ws = torch.empty(lib.bd_attention_tc_workspace_bytes(2, H, Lq, Lk, split), dtype=torch.uint8, device="cuda")
    cuda_lib.call("bd_attention_tc", q.data_ptr(), E, Lq * E, kv.data_ptr(), 2 * E, Lk * 2 * E,
                  kv[..., E:].data_ptr(), 2 * E, Lk * 2 * E, mask.data_ptr(), out.data_ptr(), E, Lq * E, 2, H, Lq, Lk, hd,
                  1.0 / math.sqrt(hd), split, ws.data_ptr())
    assert bool(torch.isnan(out[0]).all())
    qh = q[1].reshape(Lq, H, hd).transpose(0, 1).double()
    kh = kv[1, :50, :E].reshape(50, H, hd).transpose(0, 1).double()
    vh = kv[1, :50, E:].reshape(50, H, hd).transpose(0, 1).double()
    want = ((qh @ kh.transpose(-1, -2) / math.sqrt(hd)).softmax(-1) @ vh).transpose(0, 1).reshape(Lq, E).float()
    tol = 5e-3 if split == 1 else 1e-4
    torch.testing.assert_close(out[1], want, rtol=tol, atol=tol)
