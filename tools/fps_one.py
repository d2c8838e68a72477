"""One bd_grid_build + bd_fps_grid over B 50k-point scenes (ncu target). python tools/fps_one.py [B] [warps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from butd_detr_b200 import _lib, synth
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2:
    lib.bd_fps_grid_set_warps(int(sys.argv[2]))
N, m = 50000, 2048
pcs = torch.from_numpy(np.stack([synth.synth_scene(50 + b)["point_clouds"] for b in range(B)])).cuda()
ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
scratch = torch.empty(lib.bd_fps_grid_scratch_bytes(B, N), dtype=torch.uint8, device="cuda")
out = torch.zeros(B, m, dtype=torch.int32, device="cuda")
for _ in range(2):
    _lib.call("bd_grid_build", pcs.data_ptr(), 6, B, N, 0.2, ws.data_ptr())
    _lib.call("bd_fps_grid", pcs.data_ptr(), 6, B, N, m, ws.data_ptr(), scratch.data_ptr(), out.data_ptr())
torch.cuda.synchronize()
