"""Per-entry-point CUDA-event profile of one eager forward (no ncu): python tools/event_profile.py --batch 32"""
import argparse, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib, synth
from butd_detr_b200.model import BeaUTyDETR
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--full", action="store_true", help="one line per distinct call signature")
args = ap.parse_args()
model = BeaUTyDETR(text_encoder=None, precision=args.precision, cuda_graph=False)
synth.fill_state_dict_(model.state_dict(), 0)
model = model.cuda().eval()
inputs = {k: v.cuda() for k, v in synth.synth_batch(7, args.batch, 50000, 80, 132).items()}
for _ in range(2): model(inputs)
torch.cuda.synchronize()
prof = _lib.Profiler()
with prof:
    model(inputs)
torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for name, a, e0, e1 in prof.records:
    key = name if args.full else name.split("(")[0]
    tot[key][0] += e0.elapsed_time(e1) * 1e3
    tot[key][1] += 1
total = sum(v[0] for v in tot.values())
print(f"total {total:.0f} us over {sum(v[1] for v in tot.values())} launches (eager, batch {args.batch})")
for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{t:9.1f} us {100 * t / total:5.1f}%  n={n:3d}  mean {t / n:8.1f}  {k[:150]}")
