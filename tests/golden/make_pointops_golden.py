"""Runs the reference's OWN CUDA extension (oracle/_ref, built unmodified from
/root/reference/pointnet2/_ext_src by oracle/build_ref_ext.py) on the seeded cases of
tests/pointops_cases.py and stores its outputs.  Run on a B200 through gpurun:

    gpurun -- python tests/golden/make_pointops_golden.py gpurun_out/pointops_refcuda.npz

then copy the file to tests/golden/pointops_refcuda.npz.  These vectors pin the CPU oracle
(oracle/point_ops_ref.c) to the reference implementation itself.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import build_ref_ext  # noqa: E402
from pointops_cases import BALL_CASES, FPS_CASES, ball_inputs, case_seed, cloud  # noqa: E402

NN_CASES = [(512, 256, "room"), (1024, 512, "room"), (300, 2, "uniform"), (257, 1500, "lattice")]


def main(out_path):
    ext = build_ref_ext.load_ref_ext()
    assert ext is not None, "oracle/_ref/pointnet2/_ext.so missing"
    out = {}
    for name, kind, B, N, m in FPS_CASES:
        xyz = cloud(case_seed(name), N, kind, B).cuda()
        out["fps/" + name] = ext.furthest_point_sampling(xyz, m).cpu().numpy()
    for case in BALL_CASES:
        xyz, new_xyz, r, ns = ball_inputs(case)
        out["ball/" + case[0]] = ext.ball_query(new_xyz.cuda(), xyz.cuda(), r, ns).cpu().numpy().astype(np.int32)
    for n, m, kind in NN_CASES:
        unknown, known = cloud(31, n, kind, 2).cuda(), cloud(32, m, kind, 2).cuda()
        d, i = ext.three_nn(unknown, known)
        out[f"nn_dist2/{n}_{m}"] = d.cpu().numpy()
        out[f"nn_idx/{n}_{m}"] = i.cpu().numpy()
        if m >= 3:
            g = torch.Generator().manual_seed(2)
            feats = torch.randn(2, 33, m, generator=g)
            w = torch.rand(2, n, 3, generator=g)
            w = (w / w.sum(-1, keepdim=True)).contiguous()
            out[f"interp/{n}_{m}"] = ext.three_interpolate(feats.cuda(), i, w.cuda()).cpu().numpy()
    torch.cuda.synchronize()
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "pointops_refcuda.npz"))
