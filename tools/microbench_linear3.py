"""bd_linear_tc: narrow tiles with one or two CTAs per SM vs wide tiles, fp16 and bf16x3 operands."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
lib = _lib.load()
def bench(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for split in (1, 3):
    for (M, N, K) in [(16384, 288, 288), (16384, 576, 288), (65536, 288, 288), (65536, 576, 288), (65536, 864, 288), (65536, 256, 288)]:
        A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
        Y = torch.empty(M, N, device="cuda")
        out = []
        for wide, occ in ((False, 0), (False, 1), (True, 0)):
            lib.bd_linear_tc_set_occupancy(occ)
            Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split, wide=wide)
            t = bench(lambda: _lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, split))
            out.append(f"{'wide' if wide else 'narrow'}{'(2/SM)' if occ else ''}: {t:7.2f} us")
        print(f"split={split} M={M:6d} N={N:4d} K={K:4d} | " + " | ".join(out), flush=True)
lib.bd_linear_tc_set_occupancy(1)
