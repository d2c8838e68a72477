"""Attention kernel at the model's shapes (128 scenes, fp16 operands): two query tiles per CTA vs one query
tile per CTA with two CTAs per SM (bd_attention_tc_set_small_nk).  python tools/microbench_attention_tiles.py [B]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H, hd, E = 8, 36, 288


def bench(fn, n=10, warm=3):
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


print("| Lq | Lk | two tiles / CTA ms | one tile / CTA, 2 CTAs / SM ms | TFLOP/s (useful) best |")
print("|---|---|---|---|---|")
for Lq, Lk in ((1024, 1024), (1024, 80), (1024, 132), (80, 1024), (80, 80), (256, 256), (256, 80), (256, 132), (256, 1024)):
    q = torch.randn(B, Lq, E, device="cuda")
    kv = torch.randn(B, Lk, 2 * E, device="cuda")
    out = torch.empty(B, Lq, E, device="cuda")
    ws = torch.empty(lib.bd_attention_tc_workspace_bytes(B, H, Lq, Lk, 1), dtype=torch.uint8, device="cuda")
    ts = []
    for small in (0, 8):
        lib.bd_attention_tc_set_small_nk(small)
        ts.append(bench(lambda: _lib.call("bd_attention_tc", q.data_ptr(), E, Lq * E, kv.data_ptr(), 2 * E, Lk * 2 * E,
                                          kv[..., E:].data_ptr(), 2 * E, Lk * 2 * E, None, out.data_ptr(), E, Lq * E, B, H, Lq, Lk,
                                          hd, 1.0 / math.sqrt(hd), 1, ws.data_ptr())))
    lib.bd_attention_tc_set_small_nk(1 << 30)
    fl = 4.0 * B * H * Lq * Lk * hd
    print(f"| {Lq} | {Lk} | {ts[0]:.3f} | {ts[1]:.3f} | {fl / min(ts) / 1e9:.1f} |", flush=True)
