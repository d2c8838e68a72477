"""Where the training step (BASELINE.json configs[2], 8 scenes) spends its time: CUDA time per kernel (torch profiler),
total CUDA time against the wall clock of the step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from butd_detr_b200 import synth  # noqa: E402
from butd_detr_b200.model import BeaUTyDETR  # noqa: E402
from butd_detr_b200.train import GradArena  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = BeaUTyDETR(text_encoder=None)
synth.fill_state_dict_(model.state_dict(), 0)
model = model.cuda().train()
arena = GradArena(model)
opt = torch.optim.SGD([p for _, p in arena.params], lr=1e-5)
batch = {k: v.cuda() for k, v in synth.synth_batch(7, B, 50000, 80, 132).items()}


def step():
    arena.zero()
    ep = model(batch)
    loss = ep["proj_tokens"].square().mean()
    for k, v in ep.items():
        if k.endswith(("center", "pred_size", "sem_cls_scores", "proj_queries")):
            loss = loss + v.square().mean()
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    step()
torch.cuda.synchronize()
print("wall ms per step", (time.time() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in ev)
print("total CUDA ms", tot / 1e3, "launches", sum(e.count for e in ev if e.self_device_time_total > 0))
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:40]:
    print(f"{e.self_device_time_total / 1e3:9.3f} ms  n={e.count:5d}  {e.key[:110]}")
