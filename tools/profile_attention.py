"""One attention shape, a few launches — the command wrapped by ncu (tools/profile_attention.py B Lq Lk [split] [impl])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
_lib.load()
B, Lq, Lk = (int(x) for x in sys.argv[1:4])
split = int(sys.argv[4]) if len(sys.argv) > 4 else 3
impl = int(sys.argv[5]) if len(sys.argv) > 5 else 1
H, hd, E = 8, 36, 288
_lib.load().bd_attention_tc_select(impl)
q = torch.randn(B, Lq, E, device="cuda"); kv = torch.randn(B, Lk, 2 * E, device="cuda"); o = torch.empty(B, Lq, E, device="cuda")
k, v = kv[..., :E], kv[..., E:]
ws = torch.empty(_lib.load().bd_attention_tc_workspace_bytes(B, H, Lq, Lk, split), dtype=torch.uint8, device="cuda")
args = (q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E, Lk * 2 * E, None, o.data_ptr(), E, Lq * E, B, H, Lq, Lk, hd, 1 / 6.0)
for _ in range(4):
    _lib.call("bd_attention_tc", *args, split, ws.data_ptr())
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record()
for _ in range(20):
    _lib.call("bd_attention_tc", *args, split, ws.data_ptr())
e[1].record(); torch.cuda.synchronize()
print(f"B={B} Lq={Lq} Lk={Lk} split={split} impl={impl}: {e[0].elapsed_time(e[1]) / 20 * 1e3:.1f} us per call (pack + main)")
