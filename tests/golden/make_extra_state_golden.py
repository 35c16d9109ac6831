"""Generate tests/golden/cpu_extra_state.npz: the reference's own ``NeRFRenderer.update_extra_state``
(nerf/renderer.py:445-538) run on CPU torch with an INJECTED random stream.

The class is lifted from the reference source with ``ast`` and exec'd unmodified.  What it calls outside of torch is
replaced by pure-torch stand-ins (``raymarching.morton3D / morton3D_invert / packbits`` -- CUDA-only in the reference;
``custom_meshgrid`` is lifted from nerf/utils.py), ``self.density`` is an arithmetic-only analytic field, and the name
``torch`` inside the lifted code is a proxy whose ``rand_like`` / ``randint`` hand out the draws of
``oracle.density_draws`` (the counter-based stream of csrc/density.cu) instead of torch's unseeded global generator --
everything else is torch's own.  Stored: the inputs' seeds and, per case, the resulting bitfield, mean density,
mean_count, a SHA-256 of the updated density grid (over the cells drawn at most once: a cell drawn twice keeps an
arbitrary one of its values under index_put, on the GPU and on CPU torch alike) and a strided sample of the whole grid.

Cases: bound 1 / 2 (one / two cascades) x full sweep (iter_density < 16) / partial update (iter_density >= 16).

Run:  python tests/golden/make_extra_state_golden.py"""
import hashlib
import math
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_cpu_golden import lift, REF, OUT  # noqa: E402
import gpu_inputs  # noqa: E402,F401  (puts the repo root on sys.path)

H = 128


def _expand(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def _compact(x):
    x = x & 0x49249249
    x = (x | (x >> 2)) & 0xc30c30c3
    x = (x | (x >> 4)) & 0x0f00f00f
    x = (x | (x >> 8)) & 0xff0000ff
    x = (x | (x >> 16)) & 0x0000ffff
    return x


def _morton3D(coords):
    c = coords.long()
    return (_expand(c[:, 0]) | (_expand(c[:, 1]) << 1) | (_expand(c[:, 2]) << 2)).int()


def _morton3D_invert(indices):
    i = indices.long()
    return torch.stack([_compact(i), _compact(i >> 1), _compact(i >> 2)], -1).int()


def _packbits(grid, thresh, bitfield):
    bits = (grid.reshape(-1, 8) > thresh).to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8)
    bitfield.copy_((bits * w).sum(-1).to(torch.uint8))
    return bitfield


def analytic_sigma(x):
    """arithmetic only (no transcendental whose rounding could differ between torch, numpy and CUDA):
    30 * relu(0.55 - max|x_i|)^2 + 3 * relu(0.2 - |x_0 - 0.5|)"""
    a = x.abs()
    m = torch.maximum(torch.maximum(a[:, 0], a[:, 1]), a[:, 2])
    return 30.0 * torch.relu(0.55 - m) ** 2 + 3.0 * torch.relu(0.2 - (x[:, 0] - 0.5).abs())


class TorchProxy:
    """`torch` as the lifted method sees it: rand_like / randint pop injected draws, everything else is torch"""

    def __init__(self, queue):
        self._q = queue

    def __getattr__(self, name):
        return getattr(torch, name)

    def rand_like(self, t):
        kind, arr = self._q.pop(0)
        assert kind == "jitter" and tuple(arr.shape) == tuple(t.shape), (kind, arr.shape, t.shape)
        return torch.from_numpy(arr)

    def randint(self, low, high, size, **kw):
        kind, arr = self._q.pop(0)
        if kind == "coords":
            assert high == H and tuple(size) == tuple(arr.shape)
            return torch.from_numpy(arr.astype(np.int64))
        assert kind == "occ" and low == 0
        k = (arr.astype(np.uint64) * np.uint64(high)) >> np.uint64(32)       # floor(u * Nz), the device's mapping
        return torch.from_numpy(k.astype(np.int64))


def initial_grid(cascade, seed):
    """a sparse positive grid with some untrained (-1) cells, reproducible from the seed"""
    rng = np.random.default_rng(seed)
    g = np.where(rng.uniform(size=(cascade, H ** 3)) < 0.04, rng.uniform(0.0, 4.0, (cascade, H ** 3)), 0.0).astype(np.float32)
    g[:, rng.integers(0, H ** 3, 5000)] = -1.0
    return g


def main():
    import oracle
    queue = []
    ns = {"torch": TorchProxy(queue), "nn": torch.nn, "np": np, "math": math,
          "raymarching": types.SimpleNamespace(morton3D=_morton3D, morton3D_invert=_morton3D_invert, packbits=_packbits),
          "print": lambda *a, **k: None, "pver": __import__("packaging.version").version}
    tmp = {"torch": torch, "pver": ns["pver"]}
    lift(os.path.join(REF, "nerf", "utils.py"), {"custom_meshgrid"}, tmp)
    ns["custom_meshgrid"] = tmp["custom_meshgrid"]
    lift(os.path.join(REF, "nerf", "renderer.py"), {"NeRFRenderer"}, ns)
    # meshgrid order of the full sweep (:457-469): position p <-> coords (x, y, z), x slowest
    ar = torch.arange(H, dtype=torch.int32)
    xx, yy, zz = ns["custom_meshgrid"](ar, ar, ar)
    mesh_coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    mesh_morton = _morton3D(mesh_coords).long().numpy()

    out = {}
    for tag, bound, iter_density, seed in (("b1_full_", 1, 3, 4242), ("b1_part_", 1, 16, 777), ("b2_full_", 2, 0, 31337), ("b2_part_", 2, 40, 99)):
        r = ns["NeRFRenderer"](bound=bound, cuda_ray=True, density_thresh=0.3)
        C = r.cascade
        g0 = initial_grid(C, seed)
        r.density_grid.copy_(torch.from_numpy(g0))
        r.iter_density = iter_density
        r.density = lambda x: {"sigma": analytic_sigma(x)}
        r.density_scale = 1.5
        r.step_counter[:, 0] = torch.arange(16, dtype=torch.int32) * 977 + 50000
        r.local_step = 11
        del queue[:]
        dup_mask = np.zeros((C, H ** 3), bool)     # cells drawn more than once: index_put keeps an arbitrary one of their values
        for cas in range(C):
            cseed = (seed + 7919 * cas) & 0xFFFFFFFF
            if iter_density < 16:
                d = oracle.density_draws(cseed, 0, H ** 3, H)          # jitter indexed by morton cell (the device's counter)
                queue.append(("jitter", d["jitter"][mesh_morton]))
            else:
                d = oracle.density_draws(cseed, H ** 3 // 4, H ** 3 // 4, H)
                queue += [("coords", d["coords"]), ("occ", d["occ_u32"]), ("jitter", d["jitter"])]
                cells, _ = oracle.density_cells_and_xyz(g0[cas], H, min(2 ** cas, bound), d)
                u, cnt = np.unique(cells, return_counts=True)
                dup_mask[cas, u[cnt > 1]] = True
        # (full sweep: the reference draws per (block, cascade); with S = 128 there is one block, so the order is cascade 0, 1, ...)
        r.update_extra_state()
        assert not queue
        grid = r.density_grid.numpy()
        out[tag + "bound"], out[tag + "iter_density"], out[tag + "seed"] = np.int64(bound), np.int64(iter_density), np.int64(seed)
        out[tag + "density_scale"], out[tag + "density_thresh"], out[tag + "local_step"] = np.float64(1.5), np.float64(0.3), np.int64(11)
        out[tag + "bitfield"] = r.density_bitfield.numpy().copy()
        out[tag + "mean_density"] = np.float64(r.mean_density)
        out[tag + "mean_count"] = np.int64(r.mean_count)
        # SHA-256 over the cells whose value is determined (cells drawn twice zeroed out), + a strided sample of everything
        det = np.where(dup_mask, np.float32(0), grid)
        out[tag + "grid_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(det).tobytes()).digest(), dtype=np.uint8)
        out[tag + "grid_sample"] = grid.reshape(-1)[::61].copy()
        out[tag + "n_duplicate_cells"] = np.int64(dup_mask.sum())
        print(tag, "cascade", C, "mean_density %.6f" % r.mean_density, "occupied bits", int(np.unpackbits(out[tag + "bitfield"]).sum()),
              "mean_count", r.mean_count)
    np.savez_compressed(os.path.join(OUT, "cpu_extra_state.npz"), **out)
    print("wrote cpu_extra_state.npz", os.path.getsize(os.path.join(OUT, "cpu_extra_state.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
