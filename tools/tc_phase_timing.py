import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from butd_detr_b200 import _lib
from butd_detr_b200.engine import pack_weight_tc
lib = _lib.load()
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.bd_linear_tc_set_debug(dbg.data_ptr())
for (M, N, K, ln, split) in [(65536, 288, 288, 1, 1), (65536, 288, 288, 1, 3), (65536, 576, 288, 0, 1)]:
    A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
    Y = torch.empty(M, N, device="cuda"); R = torch.randn(M, N, device="cuda"); g_ = torch.ones(N, device="cuda")
    Wp, (BN, KC, nch, nsub) = pack_weight_tc(W, split, full_rows=(ln == 1), wide=(ln == 0))
    for rep in range(3):
        dbg.zero_()
        if ln == 1:
            _lib.call("bd_linear_ln_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), R.data_ptr(), N, g_.data_ptr(), b.data_ptr(), 1e-5, Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, split)
        else:
            _lib.call("bd_linear_tc", A.data_ptr(), K, None, 0, Wp.data_ptr(), b.data_ptr(), Y.data_ptr(), N, M, N, K, KC, nch, BN, nsub, 0, split)
        torch.cuda.synchronize()
    d = dbg.cpu().tolist(); t0 = d[0]
    names = {0: "start", 1: "prologue", 2: "first loads issued", 40: "mma done (epilogue starts)", 41: "tile written", 42: "end"}
    for c in range(nch):
        names[4 + c] = f"chunk {c} staged"
        names[10 + 3 * c] = f"      mma warp: W({c}) landed"
        names[11 + 3 * c] = f"      mma warp: A({c}) ready"
        names[12 + 3 * c] = f"      mma warp: MMA({c}) issued"
        names[30 + c] = f"   stage free for chunk {c}"
    ev = sorted((v - t0, names.get(i, str(i))) for i, v in enumerate(d) if v)
    print(f"M={M} N={N} K={K} ln={ln} split={split} KC={KC} chunks={nch} nsub={nsub}")
    prev = 0
    for t, n in ev:
        print(f"   {t:8d} cyc (+{t - prev:6d})  {n}")
        prev = t
