"""Does programmatic dependent launch shorten a chain of dependent small GEMMs? (graph replay and eager)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
lib = _lib.load()
for M in (256, 8192):
    N = K = 288
    xs = [torch.randn(M, K, device="cuda") for _ in range(2)]
    W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.zeros(N, device="cuda")
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, 3)
    def chain(n=40):
        for i in range(n):
            a, y = xs[i & 1], xs[(i + 1) & 1]
            _lib.call("bd_linear_tc", a.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, 3)
    for pdl in (0, 1):
        lib.bd_set_pdl(pdl)
        chain(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); chain(); e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 40 * 1e3
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain()
        g.replay(); torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        print(f"M={M} pdl={pdl}: eager {eager:.2f} us/kernel, graph {e0.elapsed_time(e1) / 40 * 1e3:.2f} us/kernel", flush=True)
lib.bd_set_pdl(1)
