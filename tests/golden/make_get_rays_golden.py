"""Generate tests/golden/cpu_get_rays.npz: the reference's own ``get_rays`` (nerf/utils.py:53-140), lifted with ``ast`` and
run on CPU torch (nerf/utils.py imports tensorboardX / mcubes / lpips ... at module level, so the function is lifted
together with ``custom_meshgrid`` and exec'd unmodified).

Run:  python tests/golden/make_get_rays_golden.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_cpu_golden import lift, REF, OUT  # noqa: E402
import gpu_inputs  # noqa: E402,F401


def main():
    from seal3d_b200 import synth
    ns = {"torch": torch, "np": np, "pver": __import__("packaging.version").version}
    lift(os.path.join(REF, "nerf", "utils.py"), {"custom_meshgrid", "get_rays"}, ns)
    poses = torch.from_numpy(synth.make_poses(3, seed=11))
    intr = np.array([1111.111, 1050.5, 400.0, 390.25])
    H, W = 37, 52                                              # a small non-square image for the all-pixels case
    full = ns["get_rays"](poses, intr, H, W, -1)
    torch.manual_seed(5)
    some = ns["get_rays"](poses, np.array([1111.111, 1111.111, 400.0, 400.0]), 800, 800, 2048)
    np.savez_compressed(os.path.join(OUT, "cpu_get_rays.npz"), poses=poses.numpy(), intr_full=intr, H=H, W=W,
                        full_o=full["rays_o"].numpy(), full_d=full["rays_d"].numpy(),
                        inds=some["inds"].numpy(), some_o=some["rays_o"].numpy(), some_d=some["rays_d"].numpy())
    print("wrote cpu_get_rays.npz", os.path.getsize(os.path.join(OUT, "cpu_get_rays.npz")) // 1024, "KB", full["rays_d"].shape, some["inds"].shape)


if __name__ == "__main__":
    main()
