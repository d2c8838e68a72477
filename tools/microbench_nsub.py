import math, os, sys
sys.path.insert(0, "/root/repo")
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
_lib.load()
def bench(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(32768, 576, 288), (32768, 288, 288), (8192, 864, 288), (8192, 576, 288), (32768, 256, 288)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda"); Y = torch.empty(M, N, device="cuda")
    for nsub in ("1", "2"):
        wide = nsub != "1"
        Wp, (BN, KC, nch, ns) = pack_weight_tc(W, 3, wide=wide)
        t = bench(lambda: _lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, KC, nch, BN, ns, 0, 3))
        print(f"M={M} N={N} K={K} n_sub={ns} BN={BN}: {t:.1f} us  {2.0*M*N*K/t/1e6:.1f} TF/s")
