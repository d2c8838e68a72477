"""Dumps the reference BeaUTyDETR's state_dict schema (names, shapes, dtypes; RoBERTa excluded)
to tests/golden/state_dict_spec.json.  Container-only (needs /root/reference)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

model = ref_loader.build_reference_model()
spec = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()
        if not k.startswith("text_encoder.")}
params = {k for k, _ in model.named_parameters()}
out = {"tensors": spec, "buffers": sorted(k for k in spec if k not in params)}
with open(os.path.join(ROOT, "tests", "golden", "state_dict_spec.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print(len(spec), "tensors,", len(out["buffers"]), "buffers")
