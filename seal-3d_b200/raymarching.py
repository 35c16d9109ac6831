"""Host-side mirror of the reference's ``raymarching/raymarching.py`` (same function names, argument
order and return values), driving the sm_100a kernels of csrc/raymarching.cu through the C-ABI.

Differences a caller can observe (all documented in DESIGN.md):
  * sample slots are ray-major and deterministic (``rays[i] = (i, offset_i, count_i)``);
  * the exact-size mode (``force_all_rays`` or ``mean_count <= 0``) counts first and allocates the
    sample buffers once at the right size instead of zero-filling ``N * max_steps`` rows
    (raymarching.py:196-207) and trimming after a D2H read; the one host read of the total stays;
  * launches use the current torch stream.
"""
import torch
from torch.autograd import Function

from . import _lib

__all__ = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "composite_rays_train", "march_rays", "composite_rays"]


def _f32c(t):
    return t.contiguous().float()


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    """raymarching.py:19-49.  rays_o/d [N,3], aabb [6] -> nears, fars [N] float32."""
    rays_o, rays_d = _f32c(rays_o.cuda()).view(-1, 3), _f32c(rays_d.cuda()).view(-1, 3)
    aabb = _f32c(aabb.to(rays_o.device))
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty_like(nears)
    _lib.call("s3d_near_far_from_aabb", rays_o, rays_d, aabb, N, float(min_near), nears, fars)
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    """raymarching.py:52-80 -> coords [N,2] in [-1,1]."""
    rays_o, rays_d = _f32c(rays_o.cuda()).view(-1, 3), _f32c(rays_d.cuda()).view(-1, 3)
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
    _lib.call("s3d_sph_from_ray", rays_o, rays_d, float(radius), N, coords)
    return coords


def morton3D(coords):
    """raymarching.py:83-104: int [N,3] in [0,1024) -> int32 [N]."""
    coords = coords.cuda().int().contiguous()
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    _lib.call("s3d_morton3D", coords, N, indices)
    return indices


def morton3D_invert(indices):
    """raymarching.py:106-126: int [N] -> int32 [N,3]."""
    indices = indices.cuda().int().contiguous()
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    _lib.call("s3d_morton3D_invert", indices, N, coords)
    return coords


def packbits(grid, thresh, bitfield=None):
    """raymarching.py:129-155: grid float [C, H^3] -> uint8 [C*H^3/8], bit = grid > thresh."""
    grid = _f32c(grid.cuda())
    N = grid.shape[0] * grid.shape[1] // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    _lib.call("s3d_packbits", grid, N, float(thresh), bitfield)
    return bitfield


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, noises=None):
    """raymarching.py:161-235.  Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays int32 [N,3].
    `noises` (optional float [N]) overrides the perturbation draws, for reproducible tests."""
    rays_o, rays_d = _f32c(rays_o.cuda()).view(-1, 3), _f32c(rays_d.cuda()).view(-1, 3)
    density_bitfield = density_bitfield.cuda().contiguous()
    nears, fars = _f32c(nears), _f32c(fars)
    dev = rays_o.device
    N = rays_o.shape[0]
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    if noises is None:
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else None
    else:
        noises = _f32c(noises)
    budgeted = (not force_all_rays) and mean_count > 0
    if budgeted:
        M = int(mean_count)
        if align > 0:
            M += align - M % align
        xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
        _lib.call("s3d_march_rays_train", rays_o, rays_d, density_bitfield, float(bound), float(dt_gamma), int(max_steps), N,
                  int(C), int(H), M, nears, fars, xyzs, dirs, deltas, rays, step_counter, noises)
        return xyzs, dirs, deltas, rays
    _lib.call("s3d_march_rays_train_count", rays_o, rays_d, density_bitfield, float(bound), float(dt_gamma), int(max_steps), N,
              int(C), int(H), nears, fars, rays, step_counter, noises)
    m = int(step_counter[0].item())  # the same single D2H read the reference does (raymarching.py:224)
    M = m
    if align > 0:
        M += align - M % align
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    _lib.call("s3d_march_rays_train_write", rays_o, rays_d, density_bitfield, float(bound), float(dt_gamma), int(max_steps), N,
              int(C), int(H), M, nears, fars, xyzs, dirs, deltas, rays, noises)
    return xyzs, dirs, deltas, rays


class _composite_rays_train(Function):
    """raymarching.py:238-291 (grad_depth is not propagated, like the reference)."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, deltas = _f32c(sigmas), _f32c(rgbs), _f32c(deltas)
        rays = rays.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        _lib.call("s3d_composite_rays_train_forward", sigmas, rgbs, deltas, rays, M, N, float(T_thresh), weights_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, image)
        ctx.dims = (M, N, float(T_thresh))
        return weights_sum, depth, image

    @staticmethod
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, deltas, rays, weights_sum, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_weights_sum, grad_image = _f32c(grad_weights_sum), _f32c(grad_image)
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _lib.call("s3d_composite_rays_train_backward", grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image,
                  M, N, T_thresh, grad_sigmas, grad_rgbs)
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
               perturb=False, dt_gamma=0, max_steps=1024):
    """raymarching.py:297-348 (inference): xyzs, dirs [n_alive*n_step (aligned), 3], deltas [.., 2]; unwritten rows stay 0."""
    rays_o, rays_d = _f32c(rays_o.cuda()).view(-1, 3), _f32c(rays_d.cuda()).view(-1, 3)
    dev = rays_o.device
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb else None
    _lib.call("s3d_march_rays", int(n_alive), int(n_step), rays_alive, rays_t, rays_o, rays_d, float(bound), float(dt_gamma),
              int(max_steps), int(C), int(H), density_bitfield, near, far, xyzs, dirs, deltas, noises)
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """raymarching.py:351-373: accumulates in place, marks finished rays with -1."""
    _lib.call("s3d_composite_rays", int(n_alive), int(n_step), float(T_thresh), rays_alive, rays_t, _f32c(sigmas), _f32c(rgbs),
              deltas, weights_sum, depth, image)
    return tuple()
