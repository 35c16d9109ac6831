"""Generate tests/golden/cpu_untrained.npz: the reference's own ``NeRFRenderer.mark_untrained_grid``
(nerf/renderer.py:379-443) run on CPU torch.  The class is lifted from the reference source with ``ast`` (its module
imports trimesh / raymarching at the top, which are absent or CUDA-only) and exec'd unmodified against
``custom_meshgrid`` lifted from nerf/utils.py and a pure-torch ``raymarching.morton3D`` (bit interleave).  The stored
result is the bit-packed mask of cells the method marks -1, for one and for two cascades.

Run:  python tests/golden/make_untrained_golden.py"""
import math
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_cpu_golden import lift, REF, OUT  # noqa: E402
import gpu_inputs  # noqa: E402,F401  (puts the repo root on sys.path)


def _morton3D(coords):
    def expand(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    c = coords.long()
    return (expand(c[:, 0]) | (expand(c[:, 1]) << 1) | (expand(c[:, 2]) << 2)).int()


def main():
    from seal3d_b200 import synth
    ns = {"torch": torch, "nn": torch.nn, "np": np, "math": math, "raymarching": types.SimpleNamespace(morton3D=_morton3D),
          "print": lambda *a, **k: None, "pver": __import__("packaging.version").version}
    lift(os.path.join(REF, "nerf", "utils.py"), {"custom_meshgrid"}, ns)
    lift(os.path.join(REF, "nerf", "renderer.py"), {"NeRFRenderer"}, ns)
    out = {}
    for tag, bound, n_pose, intr in (("b1_", 1, 3, (1111.11, 1111.11, 200.0, 200.0)), ("b2_", 2, 5, (900.0, 1000.0, 260.0, 180.0))):
        r = ns["NeRFRenderer"](bound=bound, cuda_ray=True)
        r.density_grid.fill_(0.5)
        poses = synth.make_poses(n_pose, seed=7)
        r.mark_untrained_grid(poses, intr)
        mask = (r.density_grid == -1).numpy()
        out[tag + "poses"], out[tag + "intrinsic"] = poses, np.array(intr, np.float64)
        out[tag + "mask_bits"] = np.packbits(mask.reshape(-1))
        out[tag + "n_marked"] = np.int64(mask.sum())
        assert ((r.density_grid == -1) | (r.density_grid == 0.5)).all()
        print(tag, "cascade", r.cascade, "marked", int(mask.sum()), "of", mask.size)
    np.savez_compressed(os.path.join(OUT, "cpu_untrained.npz"), **out)
    print("wrote cpu_untrained.npz", os.path.getsize(os.path.join(OUT, "cpu_untrained.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
