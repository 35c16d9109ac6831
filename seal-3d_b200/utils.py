"""Ray generation for the cuda-ray path -- mirror of ``nerf/utils.py::get_rays`` (:53-140) for the sampling modes the
Seal-3D trainers use: all pixels (``N = -1``, evaluation / proxy_dataset) and ``N`` uniformly random pixels shared by the
views of the batch (training).  Patch sampling and error-map sampling are not built."""
import torch

from . import _lib


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1, generator=None, inds=None):
    """poses [B,4,4] cam2world (CUDA), intrinsics (fx, fy, cx, cy) -> {'rays_o', 'rays_d' [B,N,3], 'inds' [B,N] (if N > 0)}.
    `inds` (int64 [N] or [B,N]) overrides the random draw; `generator` seeds it."""
    if error_map is not None or patch_size > 1:
        raise NotImplementedError("error-map and patch sampling of get_rays are not built (nerf/utils.py:73-113)")
    poses = poses.contiguous().float()
    _lib.check_cuda(poses)
    dev, B = poses.device, poses.shape[0]
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    results = {}
    if N > 0 or inds is not None:
        if inds is None:
            N = min(N, H * W)
            inds = torch.randint(0, H * W, size=[N], device=dev, generator=generator)      # may duplicate, like the reference
        inds = inds.to(dev).long().contiguous()
        rows = 1 if inds.dim() == 1 else inds.shape[0]
        n = inds.shape[-1]
        results["inds"] = inds.view(rows, n).expand(B, n)
    else:
        inds, rows, n = None, 1, H * W
    rays_o = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    _lib.call("s3d_get_rays", poses, B, fx, fy, cx, cy, H, W, inds, rows, n, rays_o, rays_d)
    results["rays_o"], results["rays_d"] = rays_o, rays_d
    return results
