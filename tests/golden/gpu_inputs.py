"""Seeded inputs shared by tests/golden/make_gpu_golden.py (which runs the UNMODIFIED reference extensions on them, on the
B200 box) and by the tests that compare the oracle (CPU) and our kernels (GPU) with the committed outputs
(tests/golden/gpu_ref.npz).  numpy's default_rng streams are platform independent, so only the outputs are stored."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

AABB = np.array([-1, -1, -1, 1, 1, 1], np.float32)
N_RAYS = 512


def scene():
    from seal3d_b200 import synth
    bits, grid = synth.lego_like_occupancy()
    o, d = synth.rays_for_step(0, N_RAYS)
    noises = np.random.default_rng(5).uniform(0, 1, N_RAYS).astype(np.float32)
    return dict(bits=bits, grid=grid, o=o, d=d, noises=noises)


def field_values(M, N):
    rng = np.random.default_rng(1)
    return dict(sigmas=rng.uniform(0, 40, M).astype(np.float32), rgbs=rng.uniform(0, 1, (M, 3)).astype(np.float32),
                g_ws=rng.normal(size=N).astype(np.float32), g_img=rng.normal(size=(N, 3)).astype(np.float32))


def morton_inputs():
    rng = np.random.default_rng(11)
    coords = rng.integers(0, 128, (2000, 3)).astype(np.int32)
    grid = rng.uniform(0, 20, (1, 128 * 64)).astype(np.float32)      # packbits works on any multiple of 8 cells
    return coords, grid


def grid_inputs(B=1024, seed=8):
    """the L16 / F2 / T19 hash grid of config 1 with a seeded table, B random points (three special rows)"""
    from seal3d_b200 import synth
    offsets, pls = synth.grid_offsets()
    rng = np.random.default_rng(seed)
    emb = rng.uniform(-1, 1, (int(offsets[-1]), 2)).astype(np.float32)
    x = rng.uniform(0, 1, (B, 3)).astype(np.float32)
    x[0] = [0.0, 1.0, 0.5]
    x[1] = [1.0001, 0.5, 0.5]
    x[2] = [-1e-7, 0.5, 0.5]
    g = rng.normal(size=(16, 128, 2)).astype(np.float32)              # backward runs on the first 128 points
    return offsets, pls, emb, x, g


def sh_freq_inputs():
    rng = np.random.default_rng(12)
    d = rng.normal(size=(512, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x = rng.uniform(-1, 1, (512, 3)).astype(np.float32)
    g_sh = rng.normal(size=(512, 16)).astype(np.float32)
    g_fr = rng.normal(size=(512, 3 + 3 * 2 * 6)).astype(np.float32)
    return d, x, g_sh, g_fr


def ffmlp_inputs(B=256, din=32, dh=64, dout=16, nl=2, seed=3):
    import oracle
    rng = np.random.default_rng(seed)
    nW = dh * din + dh * dh * (nl - 1) + dout * dh
    s = np.sqrt(3 / dh)
    W = oracle.round_to_half(rng.uniform(-s, s, nW).astype(np.float32))
    x = oracle.round_to_half(rng.normal(size=(B, din)).astype(np.float32))
    g = oracle.round_to_half((rng.normal(size=(B, dout)) / B).astype(np.float32))
    return dict(W=W, x=x, g=g, B=B, din=din, dh=dh, dout=dout, nl=nl)
