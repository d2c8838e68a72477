"""Prints max-abs / relative error per end_points tensor of the bf16 mode vs the reference golden."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model import build, CFG
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
gold = np.load(f"tests/golden/model_{name}.npz")
model, inputs = build(name, None, precision=sys.argv[2] if len(sys.argv) > 2 else "fp16")
ep = model({k: v.cuda() for k, v in inputs.items()}, overrides={"sample_inds": torch.from_numpy(gold["query_points_sample_inds"])})
for k in gold.files:
    if k.startswith("__") or gold[k].dtype.kind in "iub":
        continue
    g = gold[k]; e = np.abs(ep[k].float().cpu().numpy() - g)
    print(f"{k:28s} absmax {np.abs(g).max():7.3f}  max_err {e.max():.3e}  mean_err {e.mean():.3e}  rms_rel {np.sqrt((e**2).mean())/g.std():.3e}")
