"""FPS variants on 50k-point scenes: register-resident cluster kernel (bd_fps) vs the bucketed
one-CTA-per-scene kernel over the cell list (bd_grid_build + bd_fps_grid, 8 / 16 / 32 warps).
python tools/fps_bench.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from butd_detr_b200 import _lib, synth

lib = _lib.load()
N, m = 50000, 2048


def bench(fn, n=5, warm=2):
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


print("| B | bd_fps (cluster) ms | grid build ms | bd_fps_grid 16w ms | bd_fps_grid 32w ms | bd_fps_grid 8w x 2 CTAs/SM ms | identical |")
print("|---|---|---|---|---|---|---|")
for B in [int(a) for a in sys.argv[1:]] or [1, 4, 8, 16, 32, 64, 128, 148]:
    pcs = torch.from_numpy(np.stack([synth.synth_scene(50 + b)["point_clouds"] for b in range(B)])).cuda()
    a = torch.zeros(B, m, dtype=torch.int32, device="cuda")
    b = torch.zeros_like(a)
    ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
    scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8, device="cuda")
    t0 = bench(lambda: _lib.call("bd_fps", pcs.data_ptr(), 6, B, N, m, None, a.data_ptr()))
    tb = bench(lambda: _lib.call("bd_grid_build", pcs.data_ptr(), 6, B, N, 0.2, ws.data_ptr()))
    ts = []
    same = True
    for w in (16, 32, 8):
        lib.bd_fps_grid_set_warps(w)
        b.zero_()
        ts.append(bench(lambda: _lib.call("bd_fps_grid", pcs.data_ptr(), 6, B, N, m, ws.data_ptr(), scratch.data_ptr(), b.data_ptr())))
        same = same and bool(torch.equal(a, b))
    lib.bd_fps_grid_set_warps(0)
    print(f"| {B} | {t0:.3f} | {tb:.3f} | {ts[0]:.3f} | {ts[1]:.3f} | {ts[2]:.3f} | {same} |", flush=True)

# pruning statistics of the bucket kernel (counting variant)
import ctypes
B = 8
pcs = torch.from_numpy(np.stack([synth.synth_scene(50 + b)["point_clouds"] for b in range(B)])).cuda()
ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8, device="cuda")
out = torch.zeros(B, m, dtype=torch.int32, device="cuda")
_lib.call("bd_grid_build", pcs.data_ptr(), 6, B, N, 0.2, ws.data_ptr())
lib.bd_fps_grid_stats(1, None)
cnt = (ctypes.c_ulonglong * 16)()
lib.bd_fps_grid_stats(1, ctypes.cast(cnt, ctypes.c_void_p))
_lib.call("bd_fps_grid", pcs.data_ptr(), 6, B, N, m, ws.data_ptr(), scratch.data_ptr(), out.data_ptr())
torch.cuda.synchronize()
lib.bd_fps_grid_stats(0, ctypes.cast(cnt, ctypes.c_void_p))
v, bt = cnt[0], cnt[1]
nb = (N + 31) // 32
print(f"bucket visits per scene-round: {v / B / (m - 1):.1f} of {nb} buckets ({100 * v / B / (m - 1) / nb:.2f} %), "
      f"equivalent full sweeps per scene: {v / B / nb:.1f}; visit batches per warp-round: {bt / B / 16 / (m - 1):.2f}")
ph = [cnt[8 + i] / (m - 1) for i in range(7)]
print("warp 0 of scene 0, cycles per round: tests+append %.0f | list barrier %.0f | visits %.0f | post+arrive %.0f | "
      "mbarrier wait %.0f | CTA reduce %.0f | total %.0f" % (ph[0], ph[1], ph[2], ph[3], ph[5], ph[6], sum(ph)))
