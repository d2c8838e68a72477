"""Generates tests/golden/model_{c1,c2}.npz from the UNMODIFIED reference model.

Run in the build container only (needs /root/reference):
    python tests/golden/make_model_golden.py [c1] [c2]

For each config: seeded synthetic inputs + seeded weights (butd_detr_b200/synth.py) are fed
to the reference's own `BeaUTyDETR` (imported by oracle/ref_loader.py; its CUDA-only point
ops are served by the C restatement oracle/point_ops_ref.c) in eval mode, fp32, on CPU, and
the resulting `end_points` tensors are stored.  The inputs/weights are NOT stored — they are
regenerated from the seeds; a checksum guards against generator drift.
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from butd_detr_b200 import synth  # noqa: E402
from oracle import model_ref, ref_loader  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0]: 4096 pts / 32 queries / 16 tokens, 1 encoder + 1 decoder layer
    "c1": dict(n_points=4096, num_queries=32, n_tokens=16, n_boxes=32, enc=1, dec=1, batch=2, seed=11),
    # BASELINE.json configs[1]: 50k pts / 256 queries / 80 tokens / 132 boxes, 3 enc + 6 dec
    "c2": dict(n_points=50000, num_queries=256, n_tokens=80, n_boxes=132, enc=3, dec=6, batch=1, seed=12),
}
WEIGHT_SEED = 0

C2_KEYS_EXTRA = ("sa1_inds", "sa2_inds", "fp2_inds", "seeds_obj_cls_logits", "query_points_sample_inds",
                 "proj_tokens", "text_memory", "query_points_xyz")


def checksum(inputs):
    c = 0
    for k in sorted(inputs):
        c = zlib.crc32(inputs[k].contiguous().numpy().tobytes(), c)
    return c


def main(names):
    torch.set_num_threads(os.cpu_count())
    for name in names:
        cfg = CONFIGS[name]
        model = ref_loader.build_reference_model(num_queries=cfg["num_queries"], num_decoder_layers=cfg["dec"])
        model.cross_encoder.layers = model.cross_encoder.layers[:cfg["enc"]]
        model.cross_encoder.num_layers = cfg["enc"]
        sd = model.state_dict()
        synth.fill_state_dict_(sd, WEIGHT_SEED)
        model.load_state_dict(sd)
        model.eval()
        inputs = synth.synth_batch(cfg["seed"], cfg["batch"], cfg["n_points"], cfg["n_tokens"], cfg["n_boxes"])
        ep = ref_loader.run_reference(model, inputs)
        # the restatement must agree with the reference before we trust either
        sd2 = {k: v for k, v in sd.items() if not k.startswith("text_encoder.")}
        ep2 = model_ref.forward(sd2, inputs, cfg["num_queries"], cfg["dec"], cfg["enc"])
        out = {}
        for k, v in ep.items():
            if not torch.is_tensor(v):
                continue
            if name == "c2" and not (k in C2_KEYS_EXTRA or any(k.endswith(s) for s in (
                    "center", "pred_size", "sem_cls_scores", "proj_queries"))):
                continue
            out[k] = v.detach().cpu().numpy()
            a, b = v.float(), ep2[k].float()
            print(f"{name} {k:32s} {tuple(v.shape)} max|ref-port|={float((a - b).abs().max()):.3e}")
        out["__input_crc32"] = np.int64(checksum(inputs))
        out["__n_state_tensors"] = np.int64(len(sd2))
        path = os.path.join(ROOT, "tests", "golden", f"model_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or ["c1", "c2"])
