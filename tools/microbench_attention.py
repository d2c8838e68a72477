"""Steady-state timing of the attention kernels at the model's shapes (graph of back-to-back launches)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
_lib.load()
def bench(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
H, hd, E = 8, 36, 288
for B in (1, 8, 32):
    for (Lq, Lk, name) in [(1024, 1024, "vis self"), (256, 1024, "dec cross_v"), (1024, 132, "enc cross_d"), (1024, 80, "enc cross_vl"), (80, 1024, "enc cross_lv"), (256, 256, "dec self"), (256, 80, "dec cross_l")]:
        q = torch.randn(B, Lq, E, device="cuda"); kv = torch.randn(B, Lk, 2 * E, device="cuda"); o = torch.empty(B, Lq, E, device="cuda")
        k, v = kv[..., :E], kv[..., E:]
        args = (q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E, Lk * 2 * E, None, o.data_ptr(), E, Lq * E, B, H, Lq, Lk, hd, 1 / 6.0)
        t0 = bench(lambda: _lib.call("bd_attention_f32", *args))
        ws = torch.empty(_lib.load().bd_attention_tc_workspace_bytes(B, H, Lq, Lk, 3), dtype=torch.uint8, device="cuda")
        _lib.load().bd_attention_tc_select(0)
        l1 = bench(lambda: _lib.call("bd_attention_tc", *args, 1, ws.data_ptr()))
        l3 = bench(lambda: _lib.call("bd_attention_tc", *args, 3, ws.data_ptr()))
        _lib.load().bd_attention_tc_select(1)
        t1 = bench(lambda: _lib.call("bd_attention_tc", *args, 1, ws.data_ptr()))
        t3 = bench(lambda: _lib.call("bd_attention_tc", *args, 3, ws.data_ptr()))
        fl = 4.0 * B * H * Lq * Lk * hd
        print(f"B={B:2d} {name:13s} Lq={Lq:4d} Lk={Lk:4d} | simt {t0:8.2f} us | gen1 bf16 {l1:8.2f} bf16x3 {l3:8.2f} us | ws bf16 {t1:8.2f} us bf16x3 {t3:8.2f} us | {fl/1e9:.3f} GF -> {fl/t3/1e6:.1f} TF/s useful (x3 MMAs)", flush=True)
