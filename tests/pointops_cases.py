"""Seeded inputs shared by the point-op parity tests, the golden generator and the oracle tests."""
import zlib

import numpy as np
import torch

from butd_detr_b200 import synth


def case_seed(name):
    return zlib.crc32(name.encode()) % 1000


def cloud(seed, n, kind="room", batch=1):
    """(batch, n, 3) float32 clouds. kinds: room (ScanNet-shaped, duplicates, near-origin points),
    uniform, lattice (many exact distance ties), dup (heavy duplication)."""
    out = []
    for b in range(batch):
        g = np.random.Generator(np.random.PCG64([seed, b, 77]))
        if kind == "room":
            p = synth.synth_scene(seed * 131 + b, n, 8)["point_clouds"][:, :3]
        elif kind == "uniform":
            p = g.uniform(-2, 2, (n, 3)).astype(np.float32)
        elif kind == "lattice":
            p = g.integers(-6, 7, (n, 3)).astype(np.float32) * 0.25
        elif kind == "dup":
            base = g.uniform(-1, 1, (max(n // 8, 1), 3)).astype(np.float32)
            p = base[g.integers(0, len(base), n)]
        else:
            raise ValueError(kind)
        out.append(p)
    return torch.from_numpy(np.stack(out)).contiguous()


FPS_CASES = [  # (name, kind, batch, N, m)
    ("room4096", "room", 2, 4096, 2048),
    ("room50k", "room", 1, 50000, 2048),
    ("uniform2048", "uniform", 2, 2048, 1024),
    ("uniform1024", "uniform", 1, 1024, 512),
    ("uniform512", "uniform", 3, 512, 256),
    ("lattice3000", "lattice", 2, 3000, 700),
    ("dup9000", "dup", 1, 9000, 1500),
    ("tiny37", "uniform", 2, 37, 20),
    ("lattice300", "lattice", 1, 300, 300),
    ("uniform20000", "uniform", 2, 20000, 600),
]

BALL_CASES = [  # (name, kind, batch, n, m, radius, nsample)
    ("sa1_4096", "room", 2, 4096, 2048, 0.2, 64),
    ("sa2", "room", 2, 2048, 1024, 0.4, 32),
    ("sa3", "room", 1, 1024, 512, 0.8, 16),
    ("sa4", "room", 1, 512, 256, 1.2, 16),
    ("lattice", "lattice", 2, 3000, 333, 0.5, 16),   # d2 == r2 exactly on many pairs
    ("empty", "uniform", 1, 700, 50, 0.01, 8),       # mostly single-hit / empty balls
]


def ball_inputs(case):
    name, kind, B, n, m, r, ns = case
    xyz = cloud(case_seed(name), n, kind, B)
    if name == "empty":
        new_xyz = cloud(5, m, "uniform", B) + 10.0 * (torch.arange(m) % 2).view(1, m, 1)
    else:
        new_xyz = xyz[:, :m].clone()
    return xyz, new_xyz.contiguous(), r, ns
