"""A/B of two library builds on the same box: python tools/ab/run_ab.py <lib.so | -> [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes
import torch
from butd_detr_b200 import _lib, synth
from butd_detr_b200.model import BeaUTyDETR
if sys.argv[1] != "-":
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
    probe = ctypes.CDLL(_lib.LIB_PATH)
    for k in list(_lib._SIGNATURES):
        if not hasattr(probe, k):
            del _lib._SIGNATURES[k]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
if os.environ.get("BD_SHORT_OFF"):
    _lib.load().bd_attention_tc_set_short(0)
if os.environ.get("BD_STREAM_OFF"):
    _lib.load().bd_linear_stream_set(0)
model = BeaUTyDETR(text_encoder=None, cuda_graph=True, precision="fp16")
synth.fill_state_dict_(model.state_dict(), 0)
model = model.cuda().eval()
pool = {k: v.cuda() for k, v in synth.synth_batch(7, 2 * B, 50000, 80, 132).items()}
batches = [{k: v[i * B:(i + 1) * B] for k, v in pool.items()} for i in range(2)]
for i in range(4):
    model(batches[i % 2])
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        model(batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    print(sys.argv[1], "ms/step", e0.elapsed_time(e1) / 10, "scenes/s", B * 1e4 / e0.elapsed_time(e1))
