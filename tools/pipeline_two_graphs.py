"""Experiment: two forward passes in flight (two engines with their own CUDA graph and static buffers, replayed on two
streams) against one pass at a time — does the serial farthest-point-sampling chain of one batch hide under the
transformer of the other?   python tools/pipeline_two_graphs.py [--batch 148] [--steps 12]"""
import argparse
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from butd_detr_b200 import BeaUTyDETR, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=148)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--in-flight", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, K, F = a.batch, a.steps, a.in_flight
    models = []
    for f in range(F):
        m = BeaUTyDETR(text_encoder=None, cuda_graph=True, precision="fp16")
        synth.fill_state_dict_(m.state_dict(), 0)
        models.append(m.to(dev).eval())
    n_pool = 4 * B
    pool = {k: v.to(dev) for k, v in bench.make_pool(n_pool, 100000).items()}

    def batch(i):
        s = (i % 4) * B
        return {k: v[s:s + B] for k, v in pool.items()}

    streams = [torch.cuda.Stream() for _ in range(F)]
    # capture (and warm up) each engine on its own stream
    for f in range(F):
        with torch.cuda.stream(streams[f]):
            for i in range(3):
                models[f](batch(i))
    torch.cuda.synchronize()
    res = {}
    # one at a time: engine 0 only
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(streams[0]):
        e0.record()
        for i in range(K):
            models[0](batch(i))
        e1.record()
    torch.cuda.synchronize()
    res["one_in_flight_scenes_per_s"] = K * B / (e0.elapsed_time(e1) * 1e-3)
    # F in flight: step i goes to engine i % F on its stream
    cur = torch.cuda.current_stream()
    e0.record(cur)
    for s in streams:
        s.wait_event(e0)
    for i in range(K):
        f = i % F
        with torch.cuda.stream(streams[f]):
            ep = models[f](batch(i))
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        cur.wait_event(ev)
    e1.record(cur)
    torch.cuda.synchronize()
    res[f"{F}_in_flight_scenes_per_s"] = K * B / (e0.elapsed_time(e1) * 1e-3)
    # same outputs?
    with torch.cuda.stream(streams[0]):
        ref = {k: v.clone() for k, v in models[0](batch(K - 1)).items() if torch.is_tensor(v)}
    torch.cuda.synchronize()
    f = (K - 1) % F
    res["max_abs_diff_vs_serial"] = max(float((ep[k].float() - ref[k].float()).abs().max()) for k in ref if ref[k].is_floating_point())
    res["batch"], res["steps"] = B, K
    print(json.dumps(res))


if __name__ == "__main__":
    main()
