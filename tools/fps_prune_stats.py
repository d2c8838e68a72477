"""How often bd_fps_ordered sweeps (library built with -DFPS_DEBUG_COUNT): python tools/fps_prune_stats.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from butd_detr_b200 import _lib
from pointops_cases import cloud
lib = _lib.load()
B, N, m = 32, 50000, 2048
xyz = cloud(5, N, "room", B).cuda()
ws = torch.empty(lib.bd_ball_query_grid_workspace_bytes(B, N), dtype=torch.uint8, device="cuda")
_lib.call("bd_grid_build", xyz.data_ptr(), 3, B, N, 0.2, ws.data_ptr())
got = torch.zeros(B, m, dtype=torch.int32, device="cuda")
_lib.call("bd_fps_ordered", xyz.data_ptr(), 3, B, N, m, lib.bd_grid_order(ws.data_ptr(), B, N), None, got.data_ptr())
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
lib.bd_fps_debug_counters(out)
print("thread sweeps %d / %d = %.3f ; warp sweeps %d / %d = %.3f" % (out[0], out[2], out[0] / max(out[2], 1), out[1], out[3], out[1] / max(out[3], 1)))
st = [out[8 + i] for i in range(8)]
print("round 1000, CTA 0 thread 0 (cycles): " + ", ".join(f"{n} +{st[i + 1] - st[i]}" for i, n in enumerate(["sweep", "redux", "syncthreads", "cta reduce", "send", "mbar wait", "decode"])))
want = torch.zeros(B, m, dtype=torch.int32, device="cuda")
for name, fn in (("bd_fps", lambda: _lib.call("bd_fps", xyz.data_ptr(), 3, B, N, m, None, want.data_ptr())),
                 ("bd_fps_ordered", lambda: _lib.call("bd_fps_ordered", xyz.data_ptr(), 3, B, N, m, lib.bd_grid_order(ws.data_ptr(), B, N), None, got.data_ptr()))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.bd_fps_debug_counters(out)
    st = [out[8 + i] for i in range(8)]
    print(name, f"{e0.elapsed_time(e1) * 1e3:.0f} us; round 1000, CTA 0 thread 0 (cycles): " + ", ".join(f"{n} +{st[i + 1] - st[i]}" for i, n in enumerate(["sweep", "redux", "syncthreads", "cta reduce", "send", "mbar wait", "decode"])))
print("equal:", bool(torch.equal(got, want)))
