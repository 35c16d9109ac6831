"""Student-teacher distillation steps (the hot path of ``SealNeRF/trainer.py`` + ``nerf/utils.py``
train_step) over a flat parameter / gradient arena with one fused Adam pass and, for data-parallel
runs, ONE NCCL all-reduce of the gradient arena per step.

  pretrain_step   SealNeRF/trainer.py:456-469  per-sample L1(sigma) + L1(rgb) against cached teacher values,
                  MLPs frozen (only the two hash tables receive gradients, trainer.py:472-488)
  finetune_step   nerf/utils.py:436-537        per-ray MSE(rgb) + L1(depth) against teacher-rendered targets
  distill_step    the fused schedule of the north star: march once on the student's occupancy, run the
                  teacher (proxy-mapped, no grad) and the student on the SAME sample buffer, composite both,
                  loss + backward in one sequence -- no teacher images are materialised.
"""
import torch
import torch.distributed as dist

from . import _lib
from . import raymarching


class ParamArena:
    """All trainable parameters as views of one contiguous fp32 buffer, in the optimizer group order of
    nerf/network.py:199-212 (encoder, sigma_net, encoder_color, encoder_dir, color_net); gradients as views of a
    second buffer, Adam moments in two more, and an fp16 shadow that the fused Adam kernel refreshes in the same pass."""

    def __init__(self, model, with_half_shadow=True, n_lr=1):
        # get_params(lr) for the NGP field, get_params(lr1, lr2) for TensoRF: pass the lr INDEX as the value to learn which
        # learning rate each optimizer group uses
        groups = [{"params": list(g["params"]), "lr": g["lr"]} for g in model.get_params(*[float(i) for i in range(n_lr)])]
        self.params = [p for g in groups for p in g["params"]]
        self.lr_index = {id(p): int(g["lr"]) for g in groups for p in g["params"]}
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        n = (self.numel + 7) // 8 * 8
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(n, dtype=torch.float16, device=dev) if with_half_shadow else None
        self.offsets = {}
        off = 0
        for p in self.params:
            k = p.numel()
            # keep the parameter's physical layout (TensoRF factors are channels_last): view the segment in stride order
            perm = sorted(range(p.dim()), key=lambda i: (p.shape[i] != 1, -p.stride(i), i))
            inv = [perm.index(i) for i in range(p.dim())]
            phys = p.data.permute(perm)
            if not phys.is_contiguous():
                raise ValueError("parameter with overlapping / non-dense strides cannot live in the arena")
            self.flat[off:off + k].view(phys.shape).copy_(phys)
            p.data = self.flat[off:off + k].view(phys.shape).permute(inv)
            p.grad = self.grad[off:off + k].view(phys.shape).permute(inv)
            self.offsets[id(p)] = (off, k)
            off += k
        if self.shadow is not None:
            _lib.call("s3d_cast_f32_to_f16", self.flat, self.shadow, n)
            for mod in model.modules():
                if hasattr(mod, "half_table") and hasattr(mod, "embeddings"):
                    o, k = self.offsets[id(mod.embeddings)]
                    mod.external_shadow = self.shadow[o:o + k].view(mod.embeddings.shape)
        self.steps = {}
        self._shadowed = [mod for mod in model.modules() if hasattr(mod, "half_table") and hasattr(mod, "embeddings")]

    def segment(self, p):
        return self.offsets[id(p)]

    def adam_step(self, lr, beta1=0.9, beta2=0.99, eps=1e-15, grad_scale=1.0, only=None):
        """torch.optim.Adam(betas=(0.9,0.99), eps=1e-15) semantics (main_SealNeRF.py:283-284); zeroes the gradients.
        `only` restricts the update to a set of parameters (pretraining freezes the MLPs, trainer.py:472-488); every
        parameter keeps its own step count for the bias correction, like torch does.  `lr` is one value or one per
        learning-rate slot of get_params (TensoRF: factors, MLPs)."""
        lrs = [float(v) for v in lr] if isinstance(lr, (tuple, list)) else None
        active = []
        for p in self.params:
            if only is not None and id(p) not in only:
                continue
            off, k = self.offsets[id(p)]
            self.steps[id(p)] = self.steps.get(id(p), 0) + 1
            plr = lrs[self.lr_index[id(p)]] if lrs is not None else float(lr)
            if active and active[-1][0] + active[-1][1] == off and active[-1][2] == self.steps[id(p)] and active[-1][3] == plr:
                active[-1] = (active[-1][0], active[-1][1] + k, active[-1][2], plr)
            else:
                active.append((off, k, self.steps[id(p)], plr))
        for off, k, st, plr in active:
            _lib.call("s3d_adam_step", self.flat[off:], self.grad[off:], self.exp_avg[off:], self.exp_avg_sq[off:],
                      self.shadow[off:] if self.shadow is not None else None, k, plr, beta1, beta2, eps,
                      st, float(grad_scale), 1, 0, None)
        if only is not None:
            self.grad.zero_()
        if self.shadow is None:
            # the kernels write the parameters through raw pointers, which does not bump torch's version counter: without an
            # arena-kept shadow, tell the encoders their cached fp16 table (GridEncoder.half_table) is stale
            for mod in self._shadowed:
                mod._shadow_version = -1


class DistillTrainer:
    def __init__(self, student, teacher=None, lr=1e-2, precision="fp32", loss_scale=1.0, bg_color=1.0, T_thresh=1e-4,
                 max_steps=1024, dt_gamma=0.0, world_size=1, update_interval=16, lr_decay_iters=None):
        self.student, self.teacher = student, teacher
        self.lr, self.bg_color, self.T_thresh, self.max_steps, self.dt_gamma = lr, float(bg_color), T_thresh, max_steps, dt_gamma
        self.precision, self.loss_scale = precision, float(loss_scale)
        self.world_size = world_size
        self.update_interval = update_interval
        self.arena = ParamArena(student, with_half_shadow=(precision == "fp16"), n_lr=len(lr) if isinstance(lr, (tuple, list)) else 1)
        dev = self.arena.flat.device
        self.loss_buf = torch.zeros(2, dtype=torch.float32, device=dev)
        self.global_step = 0
        # LambdaLR position (main_SealNeRF.py:287-288): only train steps advance it -- the reference's pretraining leaves the
        # scheduler alone (SealNeRF/trainer.py:431-432, commented out) and runs at a forced, constant lr (:491-503)
        self.sched_step, self.lr_decay_iters, self.lr_forced = 0, lr_decay_iters, False
        # pretraining freezes the NGP MLPs (SealNeRF/trainer.py:472-488 freeze_mlp); for TensoRF nothing is frozen (:476-483)
        self._tables_only = ({id(student.encoder.embeddings), id(student.encoder_color.embeddings)}
                             if hasattr(student, "encoder_color") else None)

    # -- helpers ------------------------------------------------------------------------------
    def _autocast(self):
        return torch.autocast(device_type="cuda", dtype=torch.float16, enabled=(self.precision == "fp16"))

    def current_lr(self):
        """lr * 0.1 ** min(sched_step / iters, 1); a forced lr (pretraining) is not decayed"""
        if not self.lr_decay_iters or self.lr_forced:
            return self.lr
        k = 0.1 ** min(self.sched_step / float(self.lr_decay_iters), 1.0)
        return [v * k for v in self.lr] if isinstance(self.lr, (tuple, list)) else self.lr * k

    def _reduce_and_step(self, only=None, advance_schedule=True):
        if self.world_size > 1:
            dist.all_reduce(self.arena.grad)  # the single collective of the step (NCCL over NVLink / NVSwitch)
        self.arena.adam_step(self.current_lr(), grad_scale=1.0 / (self.world_size * self.loss_scale), only=only)
        self.global_step += 1
        if advance_schedule:
            self.sched_step += 1

    def _student_render(self, rays_o, rays_d, perturb, force_all_rays):
        s = self.student
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, s.aabb_train, s.min_near)
        counter = s.step_counter[s.local_step % 16]
        counter.zero_()
        s.local_step += 1
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, s.bound, s.density_bitfield, s.cascade, s.grid_size,
                                                               nears, fars, counter, s.mean_count, perturb, 128, force_all_rays,
                                                               self.dt_gamma, self.max_steps)
        return xyzs, dirs, deltas, rays

    def _ray_loss_backward(self, sig_s, rgb_s, deltas, rays, image_t, depth_t):
        ws, depth, comp = raymarching.composite_rays_train(sig_s, rgb_s, deltas, rays, self.T_thresh)
        N = rays.shape[0]
        g_img = torch.empty(N, 3, dtype=torch.float32, device=comp.device)
        g_ws = torch.empty(N, dtype=torch.float32, device=comp.device)
        self.loss_buf.zero_()
        _lib.call("s3d_finetune_loss", comp.detach(), ws.detach(), depth.detach(), image_t, depth_t, N, self.bg_color, self.loss_buf, g_img, g_ws)
        if self.loss_scale != 1.0:
            g_img.mul_(self.loss_scale)
            g_ws.mul_(self.loss_scale)
        torch.autograd.backward([comp, ws], [g_img, g_ws])
        return self.loss_buf

    # -- steps --------------------------------------------------------------------------------
    def pretrain_step(self, points, dirs, sigma_t, rgb_t):
        """one pretraining step on a cached point batch; returns the device loss buffer ([0] = L1 sigma + L1 rgb)"""
        self.student.train()
        with self._autocast():
            sig_s, rgb_s = self.student(points, dirs)
        sig_s, rgb_s = sig_s.float().contiguous(), rgb_s.float().contiguous()
        M = points.shape[0]
        g_s = torch.empty_like(sig_s)
        g_c = torch.empty_like(rgb_s)
        self.loss_buf.zero_()
        _lib.call("s3d_pretrain_loss", sig_s.detach(), rgb_s.detach(), sigma_t, rgb_t, M, self.loss_buf, g_s, g_c)
        if self.loss_scale != 1.0:
            g_s.mul_(self.loss_scale)
            g_c.mul_(self.loss_scale)
        torch.autograd.backward([sig_s, rgb_s], [g_s, g_c])
        self._reduce_and_step(only=self._tables_only, advance_schedule=False)
        return self.loss_buf

    def finetune_step(self, rays_o, rays_d, image_t, depth_t=None, perturb=True, force_all_rays=False):
        """photometric step against teacher-rendered targets (image_t [N,3], depth_t [N] or None)"""
        self.student.train()
        self._maybe_update_grid()
        xyzs, dirs, deltas, rays = self._student_render(rays_o.view(-1, 3), rays_d.view(-1, 3), perturb, force_all_rays)
        with self._autocast():
            sig_s, rgb_s = self.student(xyzs, dirs)
        loss = self._ray_loss_backward(sig_s.float(), rgb_s.float(), deltas, rays, image_t, depth_t)
        self._reduce_and_step()
        return loss

    def distill_step(self, rays_o, rays_d, perturb=True, force_all_rays=False):
        """fused schedule: teacher (proxy-mapped, no grad) and student on the same samples"""
        self.student.train()
        self._maybe_update_grid()
        rays_o, rays_d = rays_o.view(-1, 3), rays_d.view(-1, 3)
        xyzs, dirs, deltas, rays = self._student_render(rays_o, rays_d, perturb, force_all_rays)
        with torch.no_grad(), self._autocast():
            t = self.teacher
            mx, md, mask = t._map_samples(xyzs, dirs)
            sig_t, rgb_t = t(mx, md)
            sig_t = (t.density_scale * sig_t).float().contiguous()
            rgb_t = rgb_t.float().contiguous()
            if mask is not None:
                rgb_t = t._map_colors(mx, md, rgb_t, mask)
            ws_t, depth_t, img_t = raymarching.composite_rays_train(sig_t, rgb_t, deltas, rays, self.T_thresh)
            img_t = img_t + (1 - ws_t).unsqueeze(-1) * self.bg_color
        with self._autocast():
            sig_s, rgb_s = self.student(xyzs, dirs)
        loss = self._ray_loss_backward(sig_s.float(), rgb_s.float(), deltas, rays, img_t, depth_t)
        self._reduce_and_step()
        return loss

    def _maybe_update_grid(self):
        # nerf/utils.py:845-847: `global_step % update_extra_interval == 0`, step 0 included
        if self.update_interval and self.global_step % self.update_interval == 0:
            self.refresh_occupancy()

    @torch.no_grad()
    def refresh_occupancy(self, seed=None):
        """update_extra_state every 16 steps (nerf/utils.py:845-847).  All ranks use the same seed, so replicas stay
        identical without a broadcast."""
        seed = 1234 + self.global_step if seed is None else seed
        with self._autocast():
            self.student.update_extra_state(seed=seed)
