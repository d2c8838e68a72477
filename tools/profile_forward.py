"""Minimal eager forward loop for ncu (launch list / --set full captures).
    ncu ... python tools/profile_forward.py --batch 1 --iters 2
Never quote a time printed under a profiler as a bench number."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from butd_detr_b200 import _lib, synth  # noqa: E402
from butd_detr_b200.model import BeaUTyDETR  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--points", type=int, default=50000)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--warmup", type=int, default=0)
args = ap.parse_args()
model = BeaUTyDETR(text_encoder=None, precision=args.precision)
synth.fill_state_dict_(model.state_dict(), 0)
model = model.cuda().eval()
inputs = {k: v.cuda() for k, v in synth.synth_batch(7, args.batch, args.points, 80, 132).items()}
for _ in range(args.warmup):  # outside the profiled range (ncu --profile-from-start off): weight packing happens here
    model(inputs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
n0 = _lib.launch_count
for _ in range(args.iters):
    ep = model(inputs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches per forward:", (_lib.launch_count - n0) // args.iters)
