"""BASELINE.json configs[3] (attention-only "cls" path: 256 box tokens + 80 text tokens, no FPS / ball
query) and configs[4] (ball-query sweep: 50k points, nsample x radius, GB/s vs HBM roofline).
Writes a markdown table to stdout (committed under profiles/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from butd_detr_b200 import _lib, synth
from butd_detr_b200.model import BeaUTyDETR

lib = _lib.load()
PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def bench(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


PARTS = sys.argv[1:] or ["3", "4"]
print("## configs[3] — attention-only path (256 box tokens, 80 text tokens, 1024 seeds, 256 queries, 3 enc + 6 dec)\n")
print("| batch | precision | ms/step | scenes/s |\n|---|---|---|---|")
for prec in (("fp16", "bf16x3", "fp32") if "3" in PARTS else ()):
    model = BeaUTyDETR(text_encoder=None, precision=prec)
    synth.fill_state_dict_(model.state_dict(), 0)
    model = model.cuda().eval()
    eng = model.engine()
    for B in (1, 8, 32, 128):
        inp = {k: v.cuda() for k, v in synth.synth_batch(5, B, 2048, 80, 256).items()}
        g = torch.Generator(device="cuda").manual_seed(B)
        seed = {"features": torch.randn(B, 1024, 288, device="cuda", generator=g),
                "xyz": torch.rand(B, 1024, 3, device="cuda", generator=g) * 4 - 2,
                "inds": torch.arange(1024, device="cuda", dtype=torch.int32)[None].expand(B, -1).contiguous()}
        graph = torch.cuda.CUDAGraph()
        eng.forward(inp, {"seed": seed})
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            out = eng.forward(inp, {"seed": seed})
        ms = bench(graph.replay, 20)
        assert torch.isfinite(out["last_sem_cls_scores"]).all()
        print(f"| {B} | {prec} | {ms:.3f} | {B / ms * 1e3:.0f} |")

print("\n## configs[4] — ball-query sweep, N = 50 000 points, m = 2 048 FPS centres (idx only)\n")
print("algorithmic bytes = B (12 N + 12 m + 4 m nsample); peak = %.1f GB/s (measured)\n" % PEAK)
print("| B | radius | nsample | ordered scan us | GB/s | frac | cell list us (build + query) | query only us | GB/s | frac | identical | rule picks |\n|---|---|---|---|---|---|---|---|---|---|---|---|")
from butd_detr_b200.engine import grid_ball_query_rule
for B in ((1, 8, 64, 128) if "4" in PARTS else ()):
    pcs = torch.from_numpy(np.stack([synth.synth_scene(50 + b)["point_clouds"] for b in range(B)])).cuda()
    inds = torch.zeros(B, 2048, dtype=torch.int32, device="cuda")
    _lib.call("bd_fps", pcs.data_ptr(), 6, B, 50000, 2048, None, inds.data_ptr())
    cen = torch.empty(B, 2048, 3, device="cuda")
    _lib.call("bd_gather_rows", pcs.data_ptr(), 6, inds.data_ptr(), B, 50000, 2048, 3, cen.data_ptr(), 3)
    ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, 50000), dtype=torch.uint8, device="cuda")
    for r in (0.2, 0.4, 0.8):
        for ns in (16, 32, 64):
            o1 = torch.zeros(B, 2048, ns, dtype=torch.int32, device="cuda")
            o2 = torch.zeros_like(o1)
            t1 = bench(lambda: _lib.call("bd_ball_query", cen.data_ptr(), pcs.data_ptr(), 6, B, 50000, 2048, r, ns, o1.data_ptr()), 5, 1) * 1e3
            t2 = bench(lambda: _lib.call("bd_ball_query_grid", cen.data_ptr(), pcs.data_ptr(), 6, B, 50000, 2048, r, ns, o2.data_ptr(), ws.data_ptr()), 5, 1) * 1e3
            t3 = bench(lambda: _lib.call("bd_ball_query_grid_query", cen.data_ptr(), pcs.data_ptr(), 6, B, 50000, 2048, r, ns, o2.data_ptr(), ws.data_ptr()), 5, 1) * 1e3
            by = B * (12 * 50000 + 12 * 2048 + 4 * 2048 * ns)
            pick = "cell list" if grid_ball_query_rule(50000, r, ns, 2048) else "scan"
            print(f"| {B} | {r} | {ns} | {t1:.1f} | {by / t1 / 1e3:.1f} | {by / t1 / 1e3 / PEAK:.5f} | {t2:.1f} | {t3:.1f} | {by / t2 / 1e3:.1f} | {by / t2 / 1e3 / PEAK:.5f} | {torch.equal(o1, o2)} | {pick} |", flush=True)
