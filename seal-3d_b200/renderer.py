"""Mirror of ``nerf/renderer.py::NeRFRenderer`` for the cuda-ray path: ``run_cuda`` (training and
inference), ``update_extra_state`` (density-grid EMA + bitfield + mean_count) and ``render``.
The non-cuda-ray path (``run`` / ``sample_pdf``) is out of scope.

Two overridable hooks carry Seal-3D's teacher-side proxy mapping without duplicating the loop
(SealNeRF/renderer.py:291-316, :381-399): ``_map_samples(xyzs, dirs)`` and ``_map_colors(...)``.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from . import raymarching


class NeRFRenderer(nn.Module):
    def __init__(self, bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1):
        super().__init__()
        self._density_stats = self._mean_count_dev = None
        self._mean_density_host = 0
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = 128
        self.density_scale, self.min_near, self.density_thresh, self.bg_radius = density_scale, min_near, density_thresh, bg_radius
        aabb = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer("aabb_train", aabb)
        self.register_buffer("aabb_infer", aabb.clone())
        self.cuda_ray = cuda_ray
        if cuda_ray:
            self.register_buffer("density_grid", torch.zeros([self.cascade, self.grid_size ** 3]))
            self.register_buffer("density_bitfield", torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self.mean_density = 0
            self.iter_density = 0
            self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32))
            self.mean_count = 0
            self.local_step = 0

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    # ---- hooks (identity here; the Seal teacher overrides them) ------------------------------
    def _map_samples(self, xyzs, dirs):
        return xyzs, dirs, None

    def _map_colors(self, xyzs, dirs, rgbs, mask):
        return rgbs

    def _field(self, xyzs, dirs):
        mx, md, mask = self._map_samples(xyzs, dirs)
        sigmas, rgbs = self(mx, md)
        sigmas = self.density_scale * sigmas
        if mask is not None:
            rgbs = self._map_colors(mx, md, rgbs, mask)
        return sigmas, rgbs

    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=1e-4, **kwargs):
        """nerf/renderer.py:256-377.  rays_o/d [B,N,3] (B==1) -> {'image' [B,N,3], 'depth' [B,N], 'weights_sum'}."""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer, self.min_near)
        if bg_color is None:
            bg_color = 1
        results = {}
        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield, self.cascade,
                                                                   self.grid_size, nears, fars, counter, self.mean_count, perturb, 128,
                                                                   force_all_rays, dt_gamma, max_steps)
            sigmas, rgbs = self._field(xyzs, dirs)
            weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
            results["weights_sum"] = weights_sum
        else:
            weights_sum = torch.zeros(N, dtype=torch.float32, device=device)
            depth = torch.zeros(N, dtype=torch.float32, device=device)
            image = torch.zeros(N, 3, dtype=torch.float32, device=device)
            n_alive = N
            rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
            rays_t = nears.clone()
            step = 0
            while step < max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                                                            self.density_bitfield, self.cascade, self.grid_size, nears, fars, 128,
                                                            perturb if step == 0 else False, dt_gamma, max_steps)
                sigmas, rgbs = self._field(xyzs, dirs)
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        results["depth"] = depth.view(*prefix)
        results["image"] = image.view(*prefix, 3)
        return results

    @torch.no_grad()
    def render_single_pass(self, rays_o, rays_d, dt_gamma=0, bg_color=1, max_steps=1024, T_thresh=1e-4, field=None, **kwargs):
        """Full-image / evaluation render WITHOUT the host loop of nerf/renderer.py:335-372: the reference marches
        <= 8 steps per alive ray per iteration and compacts the alive list on the host every iteration (a boolean-mask
        gather = a device sync), which exists to skip the field for samples behind an opaque surface.  Here every sample
        of every ray is generated once (count -> scan -> write), the field runs once over all of them and the training
        compositor applies the same early termination, so the image equals the loop's image (same samples, same
        T_thresh); depth is converted to the eval convention (absolute t: + near * weights_sum, see
        raymarching.cu:845,872).  `field(xyzs, dirs) -> (sigmas, rgbs)` defaults to this module's own field."""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_infer, self.min_near)
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size,
                                                               nears, fars, None, -1, False, 128, True, dt_gamma, max_steps)
        sigmas, rgbs = (field or self._field)(xyzs, dirs)
        weights_sum, depth, image = raymarching.composite_rays_train(sigmas.float(), rgbs.float(), deltas, rays, T_thresh)
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        depth = depth + torch.where(weights_sum > 0, nears, torch.zeros_like(nears)) * weights_sum
        return {"image": image.view(*prefix, 3), "depth": depth.view(*prefix), "weights_sum": weights_sum.view(*prefix)}

    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """nerf/renderer.py:379-443: poses [B,4,4] cam2world (array or tensor), intrinsic = (fx, fy, cx, cy).  Cells of the
        density grid that no training camera sees are set to -1 so they never become occupied.  Returns the per-cell camera
        count [cascade, H^3] (morton order); `S` (the reference's chunk size) is accepted and ignored."""
        if not self.cuda_ray:
            return None
        dev = self.density_bitfield.device
        poses = torch.as_tensor(poses, dtype=torch.float32).to(dev).contiguous().view(-1, 4, 4)
        fx, fy, cx, cy = [float(v) for v in intrinsic]
        count = torch.zeros(self.cascade, self.grid_size ** 3, dtype=torch.int32, device=dev)
        grid = self.density_grid if self.density_grid.is_contiguous() else self.density_grid.contiguous()
        if poses.shape[0] <= 4000:                     # one launch marks the grid directly
            _lib.call("s3d_mark_untrained_grid", grid, poses, poses.shape[0], cx / fx, cy / fy, self.cascade, self.grid_size, float(self.bound), count)
            if grid is not self.density_grid:
                self.density_grid.copy_(grid)
            return count
        for b0 in range(0, poses.shape[0], 4000):      # shared-memory staging limit of the kernel; counts accumulate
            part = torch.zeros_like(count)
            tmp = grid.clone()
            _lib.call("s3d_mark_untrained_grid", tmp, poses[b0:b0 + 4000].contiguous(), min(4000, poses.shape[0] - b0), cx / fx, cy / fy,
                      self.cascade, self.grid_size, float(self.bound), part)
            count += part
        grid[count == 0] = -1
        if grid is not self.density_grid:
            self.density_grid.copy_(grid)
        return count

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, seed=None):
        """nerf/renderer.py:445-538: full sweep for the first 16 calls, then H^3/4 uniform + H^3/4 occupied cells; EMA-max
        into density_grid, mean density, packbits with min(mean, density_thresh), mean_count from the ring.

        One chain of launches (csrc/density.cu) with a single 4-byte host read at the end (mean_count, which sizes the next
        march's sample buffers -- the reference reads it the same way, :536); the occupied-cell set is compacted on the
        device instead of torch.nonzero, the bitfield threshold min(mean_density, density_thresh) stays in device memory,
        and `mean_density` is read back only when somebody asks for it.  All draws (cells, jitter) are a counter-based hash
        of `seed`, so data-parallel replicas that pass the same seed stay bit-identical without a broadcast; seed=None
        draws one from torch's generator like the reference's unseeded calls.  `S` is accepted and ignored."""
        if not self.cuda_ray:
            return
        dev = self.density_bitfield.device
        H = self.grid_size
        n_cells = H ** 3
        tmp_grid = torch.full_like(self.density_grid, -1.0)
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if seed is None else int(seed)
        full = self.iter_density < 16
        n = n_cells if full else 2 * (n_cells // 4)
        cells = None if full else torch.empty(n, dtype=torch.int32, device=dev)
        xyz = torch.empty(n, 3, dtype=torch.float32, device=dev)
        for cas in range(self.cascade):
            cseed = (seed + 7919 * cas) & 0xFFFFFFFF
            if not full:
                _lib.call("s3d_density_pick_cells", self.density_grid[cas], H, n_cells // 4, n_cells // 4, cseed, cells, None)
            bound_cas = min(2 ** cas, self.bound)
            _lib.call("s3d_density_cells_to_xyz", cells, n, H, float(bound_cas), cseed, xyz)
            sig = self.density(xyz)["sigma"].reshape(-1).detach().float().contiguous()
            _lib.call("s3d_density_scatter", cells, sig, n, float(self.density_scale), tmp_grid[cas])
        if self._density_stats is None or self._density_stats.device != dev:
            self._density_stats = torch.zeros(2, dtype=torch.float32, device=dev)
            self._mean_count_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("s3d_density_grid_update", self.density_grid, tmp_grid, self.density_grid.numel(), float(decay), float(self.density_thresh),
                  self._density_stats)
        self._mean_density_host = None            # stale: the property reads stats[0] when asked
        self.iter_density += 1
        _lib.call("s3d_packbits_dev_thresh", self.density_grid, self.density_bitfield.numel(), self._density_stats[1:], self.density_bitfield)
        total_step = min(16, self.local_step)
        if total_step > 0:
            _lib.call("s3d_mean_count", self.step_counter, total_step, self._mean_count_dev)
            self.mean_count = int(self._mean_count_dev.item())     # the one host read: it sizes the next march's buffers
        self.local_step = 0

    # mean_density lives on the device after a refresh (stats[0]); reading the attribute is what costs the host read
    @property
    def mean_density(self):
        if self._mean_density_host is None and self._density_stats is not None:
            self._mean_density_host = float(self._density_stats[0].item())
        return self._mean_density_host if self._mean_density_host is not None else 0

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density_host = v

    def render(self, rays_o, rays_d, **kwargs):
        if not self.cuda_ray:
            raise NotImplementedError("only the cuda-ray path is built (nerf/renderer.py:541-573)")
        return self.run_cuda(rays_o, rays_d, **kwargs)
