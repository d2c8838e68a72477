"""Steady-state timing of the linear kernels (CUDA events around N back-to-back launches, warm L2)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
_lib.load()
dev = "cuda"
def bench(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(256, 288, 48), (256, 288, 288), (1024, 288, 288), (256, 576, 288), (2048, 288, 288), (8192, 288, 288),
                  (256, 256, 288), (131072, 64, 64), (131072, 128, 64), (32768, 128, 136)]:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / math.sqrt(K); b = torch.randn(N, device=dev)
    Y = torch.empty(M, N, device=dev); R = torch.randn(M, N, device=dev); g_ = torch.ones(N, device=dev)
    out = []
    for split in (1, 3):
        Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split)
        t = bench(lambda: _lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, split))
        out.append(f"tc split{split}: {t:7.2f} us")
        if N <= 320:
            Wp2, (BN, KC, nch, nsub) = pack_weight_tc(W, split, full_rows=True)
            t = bench(lambda: _lib.call("bd_linear_ln_tc", A.data_ptr(), K, None, 0, Wp2.data_ptr(), b.data_ptr(), R.data_ptr(), N, g_.data_ptr(), b.data_ptr(), 1e-5, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, split))
            out.append(f"ln split{split}: {t:7.2f} us")
    t = bench(lambda: _lib.call("bd_linear_f32", A.data_ptr(), K, None, 0, W.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, 0))
    out.append(f"simt: {t:7.2f} us")
    fl = 2.0 * M * N * K
    print(f"M={M:7d} N={N:4d} K={K:4d} | " + " | ".join(out) + f" | {fl/1e9:.3f} GF")
