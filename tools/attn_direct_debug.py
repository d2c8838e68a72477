"""Debug aid: one bd_attention_tc_h call on fp16 K / V (direct tensor-copy path) against a CPU fp64 reference."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from butd_detr_b200 import _lib  # noqa: E402

B, Lq, Lk = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 128, 128)
H, hd = 8, 36
E = H * hd
lib = _lib.load()
g = torch.Generator().manual_seed(1)
q32 = torch.randn(B, Lq, E, generator=g)
kv32 = torch.randn(B, Lk, 2 * E, generator=g)
q = q32.half().cuda()
kv = kv32.half().cuda()
k, v = kv[..., :E], kv[..., E:]
out = torch.full((B, Lq, E), float("nan"), device="cuda", dtype=torch.float16)
_lib.call("bd_attention_tc_h", q.data_ptr(), E, Lq * E, k.data_ptr(), 2 * E, Lk * 2 * E, v.data_ptr(), 2 * E, Lk * 2 * E,
          None, out.data_ptr(), E, Lq * E, 15, B, H, Lq, Lk, hd, 1.0 / math.sqrt(hd), 1, None)
torch.cuda.synchronize()
qh = q32.half().double().reshape(B, Lq, H, hd).transpose(1, 2)
kh = kv32[..., :E].half().double().reshape(B, Lk, H, hd).transpose(1, 2)
vh = kv32[..., E:].half().double().reshape(B, Lk, H, hd).transpose(1, 2)
want = ((qh @ kh.transpose(-1, -2) / math.sqrt(hd)).softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E)
err = (out.cpu().double() - want).abs()
print("max err", float(err.max()), "nan", int(torch.isnan(out).sum()))
print("per-head max err", err.reshape(B, Lq, H, hd).amax((0, 1, 3)).tolist())
print("per-dim max err (head 0)", err.reshape(B, Lq, H, hd)[:, :, 0].amax((0, 1)).tolist())
