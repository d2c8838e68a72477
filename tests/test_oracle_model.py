"""CPU: the torch-fp32 port of the forward (oracle/model_ref.py) reproduces the golden vectors
that tests/golden/make_model_golden.py generated from the UNMODIFIED reference model."""
import os
import zlib

import numpy as np
import pytest
import torch

from butd_detr_b200 import synth
from butd_detr_b200.model import BeaUTyDETR
from oracle import model_ref

CFG = {
    "c1": dict(n_points=4096, num_queries=32, n_tokens=16, n_boxes=32, enc=1, dec=1, batch=2, seed=11),
    "c2": dict(n_points=50000, num_queries=256, n_tokens=80, n_boxes=132, enc=3, dec=6, batch=1, seed=12),
}


def synthetic_state_dict(c):
    model = BeaUTyDETR(num_queries=c["num_queries"], num_decoder_layers=c["dec"], num_encoder_layers=c["enc"],
                       text_encoder=None)
    return synth.fill_state_dict_(model.state_dict(), 0)


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_port_matches_reference_golden(name, golden_dir, oracle_lib):
    c = CFG[name]
    gold = np.load(os.path.join(golden_dir, f"model_{name}.npz"))
    inputs = synth.synth_batch(c["seed"], c["batch"], c["n_points"], c["n_tokens"], c["n_boxes"])
    crc = 0
    for k in sorted(inputs):
        crc = zlib.crc32(inputs[k].contiguous().numpy().tobytes(), crc)
    assert crc == int(gold["__input_crc32"]), "synthetic generator drifted from the golden inputs"
    sd = synthetic_state_dict(c)
    assert len(sd) == int(gold["__n_state_tensors"])
    torch.set_num_threads(os.cpu_count())
    ep = model_ref.forward(sd, inputs, c["num_queries"], c["dec"], c["enc"])
    for k in gold.files:
        if k.startswith("__"):
            continue
        got = ep[k].numpy()
        if gold[k].dtype.kind in "iub":
            assert np.array_equal(got, gold[k]), k
        else:
            assert float(np.abs(got - gold[k]).max()) <= 2e-5, k


def test_module_state_dict_matches_reference_schema(golden_dir):
    import json
    spec = json.load(open(os.path.join(golden_dir, "state_dict_spec.json")))
    model = BeaUTyDETR(text_encoder=None)
    sd = model.state_dict()
    assert set(sd) == set(spec["tensors"])
    for k, (shape, dtype) in spec["tensors"].items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    params = {k for k, _ in model.named_parameters()}
    assert sorted(set(sd) - params) == spec["buffers"]
    # the optimiser's param-group split by name (main_utils.py:258-280) still works
    assert any("backbone_net" in k for k in params)


def test_module_refuses_cpu_tensors_in_both_modes():
    """No CPU path: the eval engine and the training forward (point operators on the CUDA kernels) both
    refuse host tensors the way the reference's ops do ("CPU not supported", ball_query.cpp:32-34)."""
    model = BeaUTyDETR(text_encoder=None, num_decoder_layers=1, num_encoder_layers=1).eval()
    x = {"point_clouds": torch.zeros(1, 1024, 6), "text_hidden": torch.zeros(1, 4, 768),
         "text_attention_mask": torch.ones(1, 4, dtype=torch.long)}
    with pytest.raises(RuntimeError, match="CPU not supported"):
        model(x)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        model.train()(x)
