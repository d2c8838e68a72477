"""Timing experiments on the warp-specialised attention kernel (debug toggles; results are garbage)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
lib = _lib.load()
H, hd, E = 8, 36, 288
def run(B, Lq, Lk, sel, split=3, n=20):
    lib.bd_attention_tc_select(sel)
    q = torch.randn(B, Lq, E, device="cuda"); kv = torch.randn(B, Lk, 2 * E, device="cuda"); o = torch.empty(B, Lq, E, device="cuda")
    k, v = kv[..., :E], kv[..., E:]
    ws = torch.empty(lib.bd_attention_tc_workspace_bytes(B, H, Lq, Lk, split), dtype=torch.uint8, device="cuda")
    args = (q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E, Lk * 2 * E, None, o.data_ptr(), E, Lq * E, B, H, Lq, Lk, hd, 1 / 6.0)
    for _ in range(3): _lib.call("bd_attention_tc", *args, split, ws.data_ptr())
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(n): _lib.call("bd_attention_tc", *args, split, ws.data_ptr())
    e[1].record(); torch.cuda.synchronize()
    lib.bd_attention_tc_select(1)
    return e[0].elapsed_time(e[1]) / n * 1e3
for (B, Lq, Lk) in [(32, 1024, 1024), (32, 256, 1024)]:
    for split in (3, 1):
        base = run(B, Lq, Lk, 1, split)
        print(f"B={B} Lq={Lq} Lk={Lk} split={split}: full {base:.1f} us | no softmax {run(B, Lq, Lk, 1 | (1 << 4), split):.1f} | no softmax, no PV {run(B, Lq, Lk, 1 | (3 << 4), split):.1f} | no softmax, no QK {run(B, Lq, Lk, 1 | (5 << 4), split):.1f} | no MMA, no softmax {run(B, Lq, Lk, 1 | (7 << 4), split):.1f} | softmax only (no MMA) {run(B, Lq, Lk, 1 | (6 << 4), split):.1f}", flush=True)
