"""Linear kernels at the transformer's shapes, narrow vs wide tiling (CUDA graph of back-to-back launches)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
_lib.load()
dev = "cuda"
def bench(fn, n=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(8192, 288, 288), (8192, 576, 288), (8192, 864, 288), (8192, 256, 288), (8192, 288, 256), (32768, 288, 288), (32768, 576, 288), (32768, 864, 288), (2560, 576, 288), (4224, 576, 288)]:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / math.sqrt(K); b = torch.randn(N, device=dev)
    Y = torch.empty(M, N, device=dev); R = torch.randn(M, N, device=dev); g_ = torch.ones(N, device=dev)
    out = []
    for wide in (False, True):
        Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 3, wide=wide)
        t = bench(lambda: _lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, 3))
        out.append(f"{'wide' if wide else 'narrow'} BN={BN} nsub={nsub}: {t:7.2f} us")
    if N <= 320:
        Wp2, (BN, KC, nch, nsub) = pack_weight_tc(W, 3, full_rows=True)
        t = bench(lambda: _lib.call("bd_linear_ln_tc", A.data_ptr(), K, None, 0, Wp2.data_ptr(), b.data_ptr(), R.data_ptr(), N, g_.data_ptr(), b.data_ptr(), 1e-5, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 3))
        out.append(f"ln: {t:7.2f} us")
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:4d} K={K:4d} | " + " | ".join(out) + f" | {fl/1e9:.2f} GF, floor(x3 @1.6PF) {3*fl/1.6e9:.1f} us", flush=True)
