import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
lib = _lib.load()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.bd_linear_tc_set_debug(dbg.data_ptr())
for (M, N, K, ln, split) in [(256, 288, 288, 0, 1), (256, 288, 288, 0, 3), (256, 288, 288, 1, 3)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
    Y = torch.empty(M, N, device="cuda"); R = torch.randn(M, N, device="cuda"); g_ = torch.ones(N, device="cuda")
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split, full_rows=bool(ln))
    for rep in range(3):
        dbg.zero_()
        if ln:
            _lib.call("bd_linear_ln_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), R.data_ptr(), N, g_.data_ptr(), b.data_ptr(), 1e-5, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, split)
        else:
            _lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, split)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist(); t0 = d[0]
    names = {50: "  c1 w waited", 51: "  c1 fenced", 52: "  c1 s0", 53: "  c1 s1", 54: "  c1 s2", 55: "  c1 s3", 56: "  c1 issued", 57: "  c1 committed", 0: "start", 1: "prologue", 2: "loads issued", 40: "mma done", 41: "epi1+dealloc", 43: "residual", 42: "end"}
    for c in range(nch):
        names.update({4 + 4 * c: f"c{c} stored", 5 + 4 * c: f"c{c} refills", 6 + 4 * c: f"c{c} synced", 7 + 4 * c: f"c{c} mma issued"})
    ev = sorted((v - t0, names.get(i, str(i))) for i, v in enumerate(d) if v)
    print(f"M={M} N={N} K={K} ln={ln} split={split} KC={KC} chunks={nch} nsub={nsub}")
    prev = 0
    for t, n in ev:
        print(f"   {t:8d} cyc (+{t - prev:6d})  {n}")
        prev = t
