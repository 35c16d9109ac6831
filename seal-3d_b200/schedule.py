"""The Seal-3D student schedule around the hot path (BASELINE config 3; mirror of the host logic of
``SealNeRF/trainer.py`` init_pretraining :88-262, pretrain_one_epoch / pretrain_part :369-451, train :265-360 and
``SealNeRF/provider.py`` proxy_dataset :19-70):

    1. init_pretraining   sample a point lattice in the edit region (``local``), around it (``surrounding``) and optionally
                          over the whole box (``global``); keep the points the proxy mapping moves / does not move; cache
                          the teacher's sigma and (colour-edited) rgb for them once.
    2. pretrain_one_epoch  one pass over the cached sets in batches: L1(sigma) + L1(rgb), MLPs frozen for the NGP backbone,
                          learning rate forced to ``pretraining_lr``.
    3. proxy_dataset      the teacher (proxy-mapped) renders every training view once: images + depths.
    4. train_one_epoch    photometric + depth steps of the student against those images.

Everything numeric goes through the trainers' kernels (``pretrain_step``, ``finetune_step``, the teacher's field and
``render``); this file is orchestration only, with the reference's defaults (main_SealNeRF.py:89-112).
Works with either engine (``fused.FusedDistillTrainer`` or ``trainer.DistillTrainer``) and either backbone.
"""
import time

import numpy as np
import torch


def sample_points(bounds, point_step=0.005, angle_step=45):
    """SealNeRF/trainer.py:609-635: an axis-aligned lattice inside bounds [2,3] or [B,2,3] and the directions obtained by
    rotating (1 - 1e-5, 0, 0) through every xyz-Euler triple on an `angle_step` grid -> (points [P,3], dirs [A,3]) float32."""
    bounds = np.asarray(bounds, dtype=np.float32)
    if bounds.ndim == 2:
        bounds = bounds[None]
    pts = []
    for lo, hi in bounds:
        ax = [torch.arange(float(lo[d]), float(hi[d]), step=point_step) for d in range(3)]
        X, Y, Z = torch.meshgrid(*ax, indexing="ij")
        pts.append(torch.stack([X, Y, Z], dim=-1).reshape(-1, 3))
    a = np.deg2rad(np.arange(0, 360, angle_step, dtype=np.float64))
    rx, ry, rz = np.meshgrid(a, a, a, indexing="ij")
    rx, ry, rz = rx.reshape(-1), ry.reshape(-1), rz.reshape(-1)
    # scipy's Rotation.from_euler('xyz', [x,y,z]) (extrinsic) = Rz(z) Ry(y) Rx(x); applied to e_x only its first column is needed
    v = (1 - 1e-5) * np.stack([np.cos(rz) * np.cos(ry), np.sin(rz) * np.cos(ry), -np.sin(ry)], -1)
    dirs = np.concatenate([v] * bounds.shape[0], 0)
    return torch.cat(pts).float(), torch.from_numpy(dirs.astype(np.float32))


def rays_from_camera(intrinsic, H, W, device="cuda"):
    """-> rays_of_view(pose) for proxy_dataset / train_one_epoch: all H*W rays of a view through the get_rays kernel
    (utils.get_rays, nerf/utils.py:53-140)"""
    from .utils import get_rays

    def rays_of_view(pose):
        out = get_rays(torch.as_tensor(np.asarray(pose, np.float32)).to(device).view(1, 4, 4), intrinsic, H, W, -1)
        return out["rays_o"][0], out["rays_d"][0]
    return rays_of_view


class SealStudentSchedule:
    def __init__(self, trainer, num_rays=4096, log=None, consistent_depth=False):
        """consistent_depth: the reference compares the student's TRAINING depth (distance from the ray's near point: the
        training compositor starts t at 0, raymarching.cu:536-547 with last_t = near at :424) with the teacher's EVAL depth
        (distance from the origin: K10 starts at rays_t = near, raymarching.cu:845), so its L1 depth term carries a constant
        offset of near * weights_sum (nerf/utils.py:484-487 with SealNeRF/provider.py:49-57).  False mirrors that; True
        converts the proxied depths to the training convention so the two depths are comparable.  Either way the term only
        shows up in the reported loss: the compositor's backward drops the depth gradient (raymarching.py:271-288)."""
        self.tr = trainer
        self.consistent_depth = consistent_depth
        self.student, self.teacher = trainer.student, trainer.teacher
        self.dev = next(self.student.parameters()).device
        self.num_rays = num_rays
        self.pretraining_data, self.pretraining_epochs, self.is_pretraining = {}, 0, False
        self.images = self.depths = self.poses = None
        self.epoch = 0
        self.timer = {"pretraining": [], "training": [], "proxy_dataset": 0.0, "init_pretraining": 0.0}
        self.log = log or (lambda *a: None)

    # -- learning rate (SealNeRF/trainer.py:491-503 set_lr) -------------------------------------------------------
    def set_lr(self, lr):
        tr = self.tr
        if lr < 0:
            if getattr(self, "_cached_lr", None) is None:
                return
            tr.lr, self._cached_lr = self._cached_lr, None
            tr.lr_forced = False
        else:
            if getattr(self, "_cached_lr", None) is None:
                self._cached_lr = tr.lr
            tr.lr = lr
            tr.lr_forced = True       # a forced lr is constant: the trainer's LambdaLR factor is bypassed while it is set

    # -- stage 1 ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _teacher_values(self, points, dirs, mapped=None, chunk=1 << 22):
        """teacher sigma / rgb for a point set (in chunks; through the fused field when the trainer has one); `mapped` =
        (points', dirs') already taken to the original space: their colours get the mapper's colour edit (:136-137)"""
        t = self.teacher
        x, d = (points, dirs) if mapped is None else mapped
        fused_teacher = getattr(self.tr, "T", None)
        sig, rgb = [], []
        for s0 in range(0, x.shape[0], chunk):
            xs, ds = x[s0:s0 + chunk].contiguous(), d[s0:s0 + chunk].contiguous()
            if fused_teacher is not None:
                sg, c, _ = fused_teacher.forward(xs, ds)
            else:
                sg, c = t(xs, ds)
                sg = t.density_scale * sg
            c = c.float().contiguous()
            if mapped is not None and t.seal_mapper.has_color_edit():
                c = t.seal_mapper.map_color(xs, ds, c)
            sig.append(sg.float())
            rgb.append(c)
        return torch.cat(sig).contiguous(), torch.cat(rgb).contiguous()

    @torch.no_grad()
    def init_pretraining(self, epochs=100, batch_size=6144000, lr=0.07, local_point_step=0.001, local_angle_step=45,
                         surrounding_point_step=0.01, surrounding_angle_step=45, surrounding_bounds_extend=0.1,
                         global_point_step=-1, global_angle_step=45, seed=0):
        """SealNeRF/trainer.py:88-262 with the defaults of main_SealNeRF.py:89-112"""
        t0 = time.perf_counter()
        self.pretraining_epochs, self.pretraining_batch_size, self.pretraining_lr = epochs, batch_size, lr
        if epochs <= 0:
            return
        mapper = self.teacher.seal_mapper
        gen = torch.Generator(device="cpu").manual_seed(seed)
        e_x = torch.tensor([1.0, 0.0, 0.0], device=self.dev)
        fill = mapper.map_data["force_fill_bound"].detach().cpu().numpy().astype(np.float32)
        aabb = self.student.aabb_train.detach().cpu().numpy()

        def steps_of(n):
            st = list(range(0, n, batch_size))
            if not st or st[-1] != n:
                st.append(n)
            return st

        def pick_dirs(dirs, n):
            return dirs[torch.randint(dirs.shape[0], (n,), generator=gen)].to(self.dev)

        def store(key, pts, dirs, sig, rgb):
            self.pretraining_data[key] = {"points": pts.contiguous(), "dirs": dirs.contiguous(), "sigma": sig, "color": rgb, "steps": steps_of(pts.shape[0])}

        if local_point_step > 0:
            pts, dirs = sample_points(fill, local_point_step, local_angle_step)
            pts = pts.to(self.dev)
            mx, md, mask = mapper.map_to_origin(pts, torch.zeros_like(pts) + e_x)
            if "map_source" in mapper.map_data:
                mask = torch.ones_like(mask)           # with map_source every point of the fill bound is kept (:114-115)
            pts, mx, md = pts[mask], mx[mask], md[mask]
            if pts.shape[0] > 0:
                sig, rgb = self._teacher_values(None, None, mapped=(mx.contiguous(), md.contiguous()))
                store("local", pts, pick_dirs(dirs, pts.shape[0]), sig, rgb)
                self.is_pretraining = True
        outside = []
        if surrounding_point_step > 0:
            b = fill.reshape(-1, 2, 3).copy()
            b[:, 0] = np.maximum(b[:, 0] - surrounding_bounds_extend, aabb[:3])
            b[:, 1] = np.minimum(b[:, 1] + surrounding_bounds_extend, aabb[3:])
            outside.append(("surrounding", b, surrounding_point_step, surrounding_angle_step))
        if global_point_step > 0:
            outside.append(("global", aabb.reshape(1, 2, 3), global_point_step, global_angle_step))
        for key, b, pstep, astep in outside:
            pts, dirs = sample_points(b, pstep, astep)
            pts = pts.to(self.dev)
            _, _, mask = mapper.map_to_origin(pts, torch.zeros_like(pts) + e_x)
            pts = pts[~mask]                           # keep the points the edit does not touch
            if pts.shape[0] == 0:
                continue
            d = pick_dirs(dirs, pts.shape[0])
            sig, rgb = self._teacher_values(pts.contiguous(), d.contiguous())
            store(key, pts, d, sig, rgb)
        torch.cuda.synchronize()
        self.timer["init_pretraining"] = time.perf_counter() - t0

    # -- stage 2 ----------------------------------------------------------------------------------------------------
    def pretrain_one_epoch(self):
        """SealNeRF/trainer.py:369-398 + pretrain_part :401-451; returns the mean loss of the epoch"""
        self.set_lr(self.pretraining_lr)
        if not self.student.density_bitfield_hacked:
            self.student.hack_bitfield()
        self.student.train()
        tot, n = None, 0
        for src in self.pretraining_data.values():
            st = src["steps"]
            for i in range(len(st) - 1):
                a, b = st[i], st[i + 1]
                loss = self.tr.pretrain_step(src["points"][a:b], src["dirs"][a:b], src["sigma"][a:b], src["color"][a:b])
                tot = loss[0].clone() if tot is None else tot + loss[0]
                n += 1
        if getattr(self.tr, "ema", None) is not None:
            self.tr.ema_update()
        return float(tot.item()) / max(n, 1) if tot is not None else 0.0

    # -- stage 3 ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def proxy_dataset(self, poses, rays_of_view, n_batch=1, intrinsic=None):
        """SealNeRF/provider.py:19-70: teacher images / depths for every pose.  rays_of_view(pose) -> (rays_o, rays_d) [HW,3]
        numpy or tensors.  Uses the fused field when the trainer has one.  With intrinsic = (fx, fy, cx, cy) the student's
        density grid is also marked for the region no training camera sees (SealNeRF/trainer.py:285-287)."""
        t0 = time.perf_counter()
        if intrinsic is not None:
            self.student.mark_untrained_grid(np.stack([np.asarray(p, np.float32) for p in poses]), intrinsic)
        if not self.teacher.density_bitfield_hacked:
            self.teacher.hack_bitfield()
        fused_teacher = getattr(self.tr, "T", None)
        images, depths = [], []
        for pose in poses:
            o, d = rays_of_view(pose)
            o = torch.as_tensor(o, dtype=torch.float32).to(self.dev).view(-1, 3)
            d = torch.as_tensor(d, dtype=torch.float32).to(self.dev).view(-1, 3)
            per = (o.shape[0] + n_batch - 1) // n_batch
            img, dep = [], []
            for s in range(0, o.shape[0], per):
                if fused_teacher is not None:
                    out = fused_teacher.render_image(o[s:s + per], d[s:s + per], bg_color=1)
                else:
                    out = self.teacher.render_single_pass(o[s:s + per], d[s:s + per], bg_color=1)
                img.append(torch.nan_to_num(out["image"], nan=0.0))
                dp = torch.nan_to_num(out["depth"], nan=0.0)
                if self.consistent_depth:
                    from . import raymarching
                    nears, _ = raymarching.near_far_from_aabb(o[s:s + per], d[s:s + per], self.teacher.aabb_infer, self.teacher.min_near)
                    dp = dp - nears * out["weights_sum"]
                dep.append(dp)
            images.append(torch.cat(img))
            depths.append(torch.cat(dep))
        self.images, self.depths, self.poses = torch.stack(images), torch.stack(depths), list(poses)
        self._rays_of_view = rays_of_view
        torch.cuda.synchronize()
        self.timer["proxy_dataset"] = time.perf_counter() - t0
        return self.images, self.depths

    # -- stage 4 ----------------------------------------------------------------------------------------------------
    def train_one_epoch(self, seed=None):
        """nerf/utils.py:823-905: one step per view; num_rays random pixels of the view (get_rays, nerf/utils.py:99-101),
        targets = the teacher's image / depth at those pixels"""
        self.set_lr(-1)
        self.student.train()
        gen = torch.Generator(device="cpu").manual_seed(self.epoch if seed is None else seed)
        order = torch.randperm(len(self.poses), generator=gen).tolist()
        tot = None
        for v in order:
            o, d = self._rays_of_view(self.poses[v])
            o = torch.as_tensor(o, dtype=torch.float32).view(-1, 3)
            d = torch.as_tensor(d, dtype=torch.float32).view(-1, 3)
            inds = torch.randint(0, o.shape[0], (self.num_rays,), generator=gen)
            ro, rd = o[inds].to(self.dev), d[inds].to(self.dev)
            inds = inds.to(self.dev)
            loss = self.tr.finetune_step(ro, rd, self.images[v][inds].contiguous(), self.depths[v][inds].contiguous(), perturb=True)
            tot = loss.clone() if tot is None else tot + loss
        return (tot / len(order)).tolist()

    # -- evaluation ---------------------------------------------------------------------------------------------
    @staticmethod
    def psnr(pred, truth):
        """nerf/utils.py:207-235 PSNRMeter.update for one image with values in [0, 1]: -10 log10(mean squared error)"""
        return float(-10.0 * torch.log10(torch.mean((pred.float() - truth.float()) ** 2)))

    @torch.no_grad()
    def evaluate(self, views=None):
        """nerf/utils.py:907-1013 evaluate_one_epoch reduced to its metric: render the proxied views with the student (EMA
        parameters if the trainer keeps an EMA, :919-921) and return the mean PSNR against the teacher's images"""
        views = range(len(self.poses)) if views is None else views
        tr = self.tr
        use_ema = getattr(tr, "ema", None) is not None and hasattr(tr, "ema_apply")
        if use_ema:
            tr.ema_apply()
        try:
            vals = []
            for v in views:
                o, d = self._rays_of_view(self.poses[v])
                o = torch.as_tensor(o, dtype=torch.float32).to(self.dev).view(-1, 3)
                d = torch.as_tensor(d, dtype=torch.float32).to(self.dev).view(-1, 3)
                fused = getattr(tr, "S", None)
                out = fused.render_image(o, d, bg_color=1) if fused is not None else self.student.render_single_pass(o, d, bg_color=1)
                vals.append(self.psnr(out["image"], self.images[v]))
        finally:
            if use_ema:
                tr.ema_restore()
        return sum(vals) / max(len(vals), 1)

    def train(self, max_epochs):
        """SealNeRF/trainer.py:323-349: the first `pretraining_epochs` epochs pretrain, the rest fine-tune"""
        first = self.epoch + 1
        history = []
        for epoch in range(first, max_epochs + 1):
            self.epoch = epoch
            if self.is_pretraining and epoch - first >= self.pretraining_epochs:
                self.is_pretraining = False
            t0 = time.perf_counter()
            if self.is_pretraining:
                history.append(("pretrain", self.pretrain_one_epoch()))
                torch.cuda.synchronize()
                self.timer["pretraining"].append(time.perf_counter() - t0)
            else:
                history.append(("train", self.train_one_epoch()))
                torch.cuda.synchronize()
                self.timer["training"].append(time.perf_counter() - t0)
            self.log(epoch, history[-1])
        return history
