"""Ray generation for the cuda-ray path -- mirror of ``nerf/utils.py::get_rays`` (:53-140): all pixels (``N = -1``,
evaluation / proxy_dataset), ``N`` uniformly random pixels shared by the views of the batch (training) and patch sampling
(``patch_size > 1``).  Error-map sampling (``torch.multinomial`` on a 128 x 128 error grid) is not built."""
import torch

from . import _lib


def patch_indices(H, W, N, patch_size, device="cpu", generator=None):
    """nerf/utils.py:73-92: N // patch_size^2 random patch corners, each expanded to a patch_size x patch_size block of
    flat pixel ids (row * W + col) -> int64 [num_patch * patch_size^2]"""
    num_patch = N // (patch_size ** 2)
    inds_x = torch.randint(0, H - patch_size, size=[num_patch], device=device, generator=generator)
    inds_y = torch.randint(0, W - patch_size, size=[num_patch], device=device, generator=generator)
    inds = torch.stack([inds_x, inds_y], dim=-1)
    pi, pj = torch.meshgrid(torch.arange(patch_size, device=device), torch.arange(patch_size, device=device), indexing="ij")
    offsets = torch.stack([pi.reshape(-1), pj.reshape(-1)], dim=-1)
    inds = (inds.unsqueeze(1) + offsets.unsqueeze(0)).view(-1, 2)
    return inds[:, 0] * W + inds[:, 1]


def error_map_indices(error_map, H, W, N, generator=None):
    """nerf/utils.py:99-115: weighted sampling on the 128 x 128 error grid (multinomial without replacement per view), each pick
    mapped to a uniformly random pixel of its coarse cell.  error_map [B, 128*128] -> (inds [B,N], inds_coarse [B,N])"""
    B = error_map.shape[0]
    dev = error_map.device
    inds_coarse = torch.multinomial(error_map, N, replacement=False, generator=generator)            # [B, N] in [0, 128*128)
    inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
    sx, sy = H / 128, W / 128
    inds_x = (inds_x * sx + torch.rand(B, N, device=dev, generator=generator) * sx).long().clamp(max=H - 1)
    inds_y = (inds_y * sy + torch.rand(B, N, device=dev, generator=generator) * sy).long().clamp(max=W - 1)
    return inds_x * W + inds_y, inds_coarse


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1, generator=None, inds=None):
    """poses [B,4,4] cam2world (CUDA), intrinsics (fx, fy, cx, cy) -> {'rays_o', 'rays_d' [B,N,3], 'inds' [B,N] (if N > 0)}.
    `inds` (int64 [N] or [B,N]) overrides the random draw; `generator` seeds it; `error_map` [B, 128*128] selects the
    reference's error-weighted sampling (adds 'inds_coarse')."""
    poses = poses.contiguous().float()
    _lib.check_cuda(poses)
    dev, B = poses.device, poses.shape[0]
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    results = {}
    if N > 0 or inds is not None:
        if inds is None:
            N = min(N, H * W)
            if patch_size > 1:
                inds = patch_indices(H, W, N, patch_size, dev, generator)
            elif error_map is not None:
                inds, results["inds_coarse"] = error_map_indices(error_map.to(dev), H, W, N, generator)
            else:
                inds = torch.randint(0, H * W, size=[N], device=dev, generator=generator)  # may duplicate, like the reference
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
