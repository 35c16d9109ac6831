"""Fused execution of the NGP field and of the distillation step (no autograd, no intermediate tensors
beyond one 128-byte feature row per sample), on the kernels of csrc/field.cu.

``FusedNGP`` wraps a ``NeRFNetwork`` (same parameters -- reference checkpoints load into the module as
usual) and keeps the device-side forms the kernels want: both hash tables interleaved into one fp16
table, the five MLP matrices in fp16, and -- for a trainable field -- an interleaved fp32 gradient table,
Adam moments and a flat fp32 arena for the MLP.  ``FusedDistillTrainer`` runs the schedule of
``DistillTrainer`` (teacher and student on the same sample buffer) on top of it, with the single
gradient all-reduce per step for data-parallel runs.
"""
import numpy as np
import os

import torch
import torch.distributed as dist

from . import _lib
from . import raymarching

_MLP = ("sigma_net.0", "sigma_net.1", "color_net.0", "color_net.1", "color_net.2")


class FusedNGP:
    def __init__(self, model, trainable=False, deterministic=None):
        """deterministic (default: S3D_DETERMINISTIC=1): accumulate the table / weight gradients as 64-bit fixed point with integer
        reductions (s3d_ngp_scatter_fixed, s3d_ngp_mlp_backward_fixed) instead of float atomics: two runs of the same steps then
        end with bit-identical parameters (the float scatter differs run to run in the last bits); about 2x the scatter time and
        a second, 8-byte-per-value arena."""
        self.model = model
        self.deterministic = (os.environ.get("S3D_DETERMINISTIC", "0") == "1") if deterministic is None else bool(deterministic)
        enc, encc = model.encoder, model.encoder_color
        assert enc.level_dim == 2 and encc.level_dim == 2 and enc.input_dim == 3 and enc.gridtype_id == 0 and not enc.align_corners
        assert torch.equal(enc.offsets, encc.offsets), "the fused field needs both grids to share their level geometry"
        self.dev = enc.embeddings.device
        self.N = enc.embeddings.shape[0]
        self.L = enc.offsets.shape[0] - 1
        self.S = float(np.log2(enc.per_level_scale))
        self.H = int(enc.base_resolution)
        self.offsets = enc.offsets
        self.bound = float(model.bound)
        self.density_scale = float(model.density_scale)
        self.weights = [model.sigma_net[0].weight, model.sigma_net[1].weight, model.color_net[0].weight, model.color_net[1].weight,
                        model.color_net[2].weight]
        assert [tuple(w.shape) for w in self.weights] == [(64, 32), (16, 64), (64, 63), (64, 64), (3, 64)], "NGP field shapes (nerf/network.py)"
        # flat fp32 arena for the MLP (master copy; the module's parameters become views of it)
        n_mlp = sum(w.numel() for w in self.weights)
        self.n_mlp = (n_mlp + 7) // 8 * 8
        self.mlp32 = torch.zeros(self.n_mlp, dtype=torch.float32, device=self.dev)
        self.mlp16 = torch.zeros(self.n_mlp, dtype=torch.float16, device=self.dev)
        off = 0
        self._w_off = []
        for w in self.weights:
            k = w.numel()
            self.mlp32[off:off + k].copy_(w.data.reshape(-1))
            w.data = self.mlp32[off:off + k].view_as(w.data)
            self._w_off.append((off, k))
            off += k
        self.table4 = torch.empty(self.N, 4, dtype=torch.float16, device=self.dev)
        # where the kernels read / the Adam kernel writes this model's fp16 entries: (tensor, byte offset, byte stride)
        self._tbl, self._tbl_off, self._tbl_stride = self.table4, 0, 8
        self.sync_from_module()
        self.trainable = trainable
        if trainable:
            for p in (enc.embeddings, encc.embeddings):
                assert p.dtype == torch.float32 and p.is_contiguous()
            # one contiguous gradient arena: [interleaved table gradient N*4 | MLP gradient] -> one all-reduce
            self.grad = torch.zeros(self.N * 4 + self.n_mlp, dtype=torch.float32, device=self.dev)
            self.grad4 = self.grad[:self.N * 4]
            self.gmlp = self.grad[self.N * 4:]
            self.m4 = torch.zeros(self.N * 4, dtype=torch.float32, device=self.dev)
            self.v4 = torch.zeros(self.N * 4, dtype=torch.float32, device=self.dev)
            self.m_mlp = torch.zeros(self.n_mlp, dtype=torch.float32, device=self.dev)
            self.v_mlp = torch.zeros(self.n_mlp, dtype=torch.float32, device=self.dev)
            self.step_tables = 0
            self.step_mlp = 0
            if self.deterministic:
                self.fixed = torch.zeros(self.N * 4 + self.n_mlp, dtype=torch.int64, device=self.dev)
                self.nonfinite = torch.zeros(1, dtype=torch.int32, device=self.dev)

    # -- state --------------------------------------------------------------------------------
    def sync_from_module(self):
        """refresh the fp16 forms after the module's parameters were changed from outside (load_state_dict, ...)"""
        enc, encc = self.model.encoder, self.model.encoder_color
        _lib.call("s3d_ngp_interleave_tables", enc.embeddings.detach(), encc.embeddings.detach(), self.table4, self.N)
        _lib.call("s3d_cast_f32_to_f16", self.mlp32, self.mlp16, self.n_mlp)

    def _table_ptr(self):
        return self._tbl.data_ptr() + self._tbl_off

    def use_paired_table(self, table8, half):
        """read / refresh this model's entries inside a paired table [N, 8] fp16 (half 0 = teacher, 1 = student)"""
        self._tbl, self._tbl_off, self._tbl_stride = table8, 8 * half, 16

    def _w16(self):
        return [self.mlp16[o:] for o, _ in self._w_off]

    def _gw(self):
        return [self.gmlp[o:] for o, _ in self._w_off]

    # -- forward ------------------------------------------------------------------------------
    def encode(self, xyz, sigma_only=False):
        M = xyz.shape[0]
        feats = torch.empty(M, 64, dtype=torch.float16, device=self.dev)
        _lib.call("s3d_ngp_encode", xyz, M, self.bound, self._table_ptr(), self._tbl_stride, self.offsets, self.L, self.S, self.H, feats, int(sigma_only))
        return feats

    def mlp_forward(self, feats, dirs, sigma_only=False, want_geo=False):
        M = feats.shape[0]
        sigma = torch.empty(M, dtype=torch.float32, device=self.dev)
        rgb = None if sigma_only else torch.empty(M, 3, dtype=torch.float32, device=self.dev)
        geo = torch.empty(M, 15, dtype=torch.float32, device=self.dev) if want_geo else None
        w = self._w16()
        _lib.call("s3d_ngp_mlp_forward", feats, dirs, M, w[0], w[1], w[2], w[3], w[4], self.density_scale, sigma, rgb, geo, int(sigma_only))
        return sigma, rgb, geo

    def forward(self, xyz, dirs):
        """(sigma * density_scale, rgb) like NeRFRenderer._field; also returns the feature rows for the backward"""
        xyz = xyz.contiguous().float()
        dirs = dirs.contiguous().float()
        feats = self.encode(xyz)
        sigma, rgb, _ = self.mlp_forward(feats, dirs)
        return sigma, rgb, feats

    def field(self, xyz, dirs):
        """drop-in for NeRFRenderer._field (includes the Seal proxy hooks of the wrapped module)"""
        m = self.model
        mx, md, mask = m._map_samples(xyz, dirs)
        sigma, rgb, _ = self.forward(mx, md)
        if mask is not None:
            rgb = m._map_colors(mx, md, rgb, mask)
        return sigma, rgb

    @torch.no_grad()
    def render_image(self, rays_o, rays_d, **kwargs):
        """single-pass full-image render through the fused field (what proxy_dataset / evaluation need)"""
        return self.model.render_single_pass(rays_o, rays_d, field=self.field, **kwargs)

    def density(self, xyz):
        """NeRFNetwork.density (nerf/network.py:130-147): {'sigma', 'geo_feat'} (sigma WITHOUT density_scale, like the module)"""
        xyz = xyz.contiguous().float()
        feats = self.encode(xyz, sigma_only=True)
        sigma, _, geo = self.mlp_forward(feats, None, sigma_only=True, want_geo=True)
        return {"sigma": sigma / self.density_scale, "geo_feat": geo}

    # -- backward / optimizer ---------------------------------------------------------------------
    def grad_chunks(self, n_chunks):
        """split the levels into n_chunks groups of whole 4-level blocks with about equal table bytes ->
        [(level_begin, level_end, arena_begin, arena_end)]; the last chunk's arena slice also carries the MLP gradients"""
        off = self.offsets.detach().cpu().tolist()
        bounds, blocks = [0], list(range(4, self.L, 4)) + [self.L]
        for k in range(1, n_chunks):
            target = off[self.L] * k / n_chunks
            cand = min((b for b in blocks if b > bounds[-1] and b < self.L), key=lambda b: abs(off[b] - target), default=None)
            if cand is None:
                break
            bounds.append(cand)
        bounds.append(self.L)
        out = []
        for a, b in zip(bounds[:-1], bounds[1:]):
            out.append((a, b, off[a] * 4, off[b] * 4 if b < self.L else self.grad.numel()))
        return out

    def _hi_stream(self):
        if getattr(self, "_hi", None) is None:
            # CUDA priorities: lower number = higher priority; torch's current stream has the lowest (0)
            self._hi = torch.cuda.Stream(device=self.dev, priority=-1)
        return self._hi

    def backward(self, xyz, dirs, feats, g_sigma, g_rgb, loss_scale=1.0, train_mlp=True, before_scatter=None, chunks=None, after_chunk=None,
                 sample_chunks=1):
        """accumulates loss_scale * dL/dparams into the gradient arena; `before_scatter()` is called between the two launches.
        chunks (from grad_chunks) + after_chunk(i, arena_begin, arena_end): scatter the levels chunk by chunk and report each
        finished arena slice (the data-parallel trainer starts its all-reduce there).

        sample_chunks > 1: the samples are cut into that many slices; the MLP backward of slice k+1 (persistent, one CTA per
        SM, tensor-core / latency bound, almost no LSU traffic) runs on a high-priority stream while the gradient scatter of
        slice k (bound by the SMs' REDG issue rate) runs on the current stream: one scatter CTA fits next to the MLP CTA on
        every SM (registers 544 x <= 80 + 256 x 64, shared memory 222 KB + 0.3 KB), so the two limiters overlap."""
        M = feats.shape[0]
        w, gw = self._w16(), self._gw()
        dfeats = torch.empty(M, 64, dtype=torch.float16, device=self.dev)
        if sample_chunks > 1 and not chunks and not self.deterministic and M >= 128 * sample_chunks:
            per = (M + sample_chunks - 1) // sample_chunks
            per = (per + 127) // 128 * 128
            main, hi = torch.cuda.current_stream(), self._hi_stream()
            hi.wait_stream(main)
            evs, spans = [], []
            with torch.cuda.stream(hi):
                for a in range(0, M, per):
                    b = min(M, a + per)
                    _lib.call("s3d_ngp_mlp_backward", feats[a:b], dirs[a:b], b - a, w[0], w[1], w[2], w[3], w[4], self.density_scale, g_sigma[a:b],
                              g_rgb[a:b], dfeats[a:b], 1.0, gw[0], gw[1], gw[2], gw[3], gw[4], int(train_mlp))
                    ev = torch.cuda.Event()
                    ev.record(hi)
                    evs.append(ev)
                    spans.append((a, b))
            if before_scatter is not None:
                before_scatter()
            for ev, (a, b) in zip(evs, spans):
                main.wait_event(ev)
                _lib.call("s3d_ngp_scatter", xyz[a:b], dfeats[a:b], b - a, self.bound, self.grad4, self.offsets, self.L, self.S, self.H, 1.0)
            return
        if self.deterministic:
            fx = self.fixed
            gf = [fx[self.N * 4 + o:] for o, _ in self._w_off]
            _lib.call("s3d_ngp_mlp_backward_fixed", feats, dirs, M, w[0], w[1], w[2], w[3], w[4], self.density_scale, g_sigma, g_rgb, dfeats,
                      1.0, gf[0], gf[1], gf[2], gf[3], gf[4], int(train_mlp), self.nonfinite)
            if before_scatter is not None:
                before_scatter()
            _lib.call("s3d_ngp_scatter_fixed", xyz, dfeats, M, self.bound, fx, self.offsets, self.L, self.S, self.H, 1.0, self.nonfinite)
            _lib.call("s3d_fixed_to_float", fx, self.grad, fx.numel(), self.nonfinite)
            if chunks and after_chunk is not None:   # the data-parallel trainer all-reduces the finished arena slices
                for i, (l0, l1, a0, a1) in enumerate(chunks):
                    after_chunk(i, a0, a1)
            return
        _lib.call("s3d_ngp_mlp_backward", feats, dirs, M, w[0], w[1], w[2], w[3], w[4], self.density_scale, g_sigma, g_rgb, dfeats,
                  1.0, gw[0], gw[1], gw[2], gw[3], gw[4], int(train_mlp))
        if before_scatter is not None:
            before_scatter()
        if chunks:
            for i, (l0, l1, a0, a1) in enumerate(chunks):
                _lib.call("s3d_ngp_scatter_levels", xyz, dfeats, M, self.bound, self.grad4, self.offsets, self.L, self.S, self.H, 1.0, l0, l1)
                if after_chunk is not None:
                    after_chunk(i, a0, a1)
            return
        _lib.call("s3d_ngp_scatter", xyz, dfeats, M, self.bound, self.grad4, self.offsets, self.L, self.S, self.H, 1.0)

    def adam_step(self, lr, grad_scale=1.0, beta1=0.9, beta2=0.99, eps=1e-15, train_mlp=True, scaler_state=None, lr_mlp=None,
                  scaler_state_mlp=None):
        """fused Adam over the tables (+ MLP arena).  `scaler_state` / `scaler_state_mlp` = the two blocks of the device
        GradScaler state (see GradScalerState): the kernels then take scale, skip decision and bias corrections from them
        instead of the host-side step counts (which, with a dynamic scaler, also count skipped steps)."""
        enc, encc = self.model.encoder, self.model.encoder_color
        self.step_tables += 1
        _lib.call("s3d_ngp_adam_tables", enc.embeddings.data, encc.embeddings.data, self.grad4, self.m4, self.v4, self._table_ptr(), self._tbl_stride,
                  self.N, float(lr), beta1, beta2, eps, self.step_tables, float(grad_scale), scaler_state)
        if train_mlp:
            self.step_mlp += 1
            _lib.call("s3d_adam_step", self.mlp32, self.gmlp, self.m_mlp, self.v_mlp, self.mlp16, self.n_mlp, float(lr if lr_mlp is None else lr_mlp),
                      beta1, beta2, eps, self.step_mlp, float(grad_scale), 1, 0, scaler_state_mlp if scaler_state is not None else None)
        else:
            self.gmlp.zero_()
        # raw-pointer parameter updates do not bump torch's version counter: mark the modules' cached fp16 tables stale
        enc._shadow_version = encc._shadow_version = -1

    # ---- parameter views for EMA / checkpoints: (tensor, ...) in the order tables, MLP arena
    def param_tensors(self):
        return [self.model.encoder.embeddings.data, self.model.encoder_color.embeddings.data, self.mlp32]


class GradScalerState:
    """torch.cuda.amp.GradScaler (nerf/utils.py:361, 857-859) with its state on the device: float[16] = two blocks of
    [scale, growth_tracker, found_inf, optimizer_steps, 1/(1-b1^t), 1/sqrt(1-b2^t), -, -] -- block 0 for the hash tables (or a
    flat arena), block 1 (`state_mlp`) for the MLP arena, whose step count advances only when the MLP is trained (torch's
    Adam keeps a step per parameter; pretraining freezes the MLP).  Nothing here syncs with the host."""

    def __init__(self, device, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self._all = torch.zeros(16, dtype=torch.float32, device=device)
        self.state, self.state_mlp = self._all[:8], self._all[8:]
        self.state[0] = float(init_scale)
        self.state_mlp[0] = float(init_scale)
        self.growth_factor, self.backoff_factor, self.growth_interval = float(growth_factor), float(backoff_factor), int(growth_interval)

    def set_steps(self, step_tables, step_mlp, beta1=0.9, beta2=0.99):
        """resume: applied-step counts of the two parameter families + the bias corrections the Adam kernels read"""
        for blk, t in ((self.state, int(step_tables)), (self.state_mlp, int(step_mlp))):
            blk[3] = float(t)
            blk[4] = 1.0 / (1.0 - beta1 ** t) if t > 0 else 0.0
            blk[5] = 1.0 / (1.0 - beta2 ** t) ** 0.5 if t > 0 else 0.0

    def steps(self):
        """(tables, MLP) optimizer steps actually applied (skipped steps are not counted) -- one host read"""
        st = self._all.detach().cpu()
        return int(st[3]), int(st[11])

    @property
    def scale_tensor(self):
        return self.state[0]      # 0-dim device tensor: usable as a multiplier without a host read

    def check(self, grad_arena, beta1=0.9, beta2=0.99, advance_mlp=True):
        _lib.call("s3d_grad_scaler_check", grad_arena, grad_arena.numel(), self.state, beta1, beta2, int(advance_mlp))

    def update(self):
        _lib.call("s3d_grad_scaler_update", self.state, self.growth_factor, self.backoff_factor, self.growth_interval)

    def get_scale(self):
        return float(self.state[0].item())


class ParamEMA:
    """torch_ema.ExponentialMovingAverage as the reference uses it (nerf/utils.py:356-357 create, :882-883 update once per
    epoch, :919-921 / :1010-1011 store + copy_to / restore around evaluation).  torch_ema is a third-party package that is
    not in the reference tree (unpinned in requirements.txt); its published update rule with use_num_updates=True:
        decay_t = min(decay, (1 + n) / (10 + n)),  n = number of updates so far (after increment)
        shadow -= (1 - decay_t) * (shadow - param)"""

    def __init__(self, tensors, decay=0.95):
        self.tensors, self.decay, self.num_updates = list(tensors), float(decay), 0
        self.shadow = [t.detach().clone() for t in self.tensors]
        self.stored = None

    def update(self):
        self.num_updates += 1
        d = min(self.decay, (1.0 + self.num_updates) / (10.0 + self.num_updates))
        for sh, t in zip(self.shadow, self.tensors):
            _lib.call("s3d_ema_update", sh, t, t.numel(), float(d))

    def store(self):
        self.stored = [t.detach().clone() for t in self.tensors]

    def copy_to(self):
        for sh, t in zip(self.shadow, self.tensors):
            t.copy_(sh)

    def restore(self):
        for st, t in zip(self.stored, self.tensors):
            t.copy_(st)
        self.stored = None


class FusedDistillTrainer:
    """Same public steps as trainer.DistillTrainer (pretrain_step / finetune_step / distill_step), fused kernels inside."""

    def __init__(self, student, teacher=None, lr=1e-2, loss_scale=None, bg_color=1.0, T_thresh=1e-4, max_steps=1024, dt_gamma=0.0,
                 world_size=1, update_interval=16, lr_decay_iters=None, ema_decay=None, scaler_kwargs=None, deterministic=None):
        """deterministic: see FusedNGP (fixed-point gradient accumulation, bit-identical runs).  loss_scale: None = static 32 x batch units (default), a float = static, "dynamic" = GradScaler semantics on the
        device (scaler_kwargs: init_scale / growth_factor / backoff_factor / growth_interval).  lr_decay_iters = the LambdaLR
        of main_SealNeRF.py:287-288, lr * 0.1 ** min(step / iters, 1), stepped every step.  ema_decay = torch_ema decay
        (0.95 in main_SealNeRF.py:292-302); call ema_update() once per epoch like nerf/utils.py:882-883."""
        self.student, self.teacher = student, teacher
        self.S = FusedNGP(student, trainable=True, deterministic=deterministic)
        self.T = FusedNGP(teacher, trainable=False) if teacher is not None else None
        # gradient tiles are fp16 (fp32 accumulation): the loss is scaled so that they sit in fp16's normal range and
        # the scale is divided out inside the Adam kernels.  The mean reductions make gradients ~ 1/units, so the
        # default scale is 32 * units (rays for the photometric loss, samples for the pretraining loss).
        self.lr, self.loss_scale, self.bg_color, self.T_thresh = lr, loss_scale, float(bg_color), T_thresh
        self.max_steps, self.dt_gamma, self.world_size, self.update_interval = max_steps, dt_gamma, world_size, update_interval
        self.loss_buf = torch.zeros(2, dtype=torch.float32, device=self.S.dev)
        self.global_step = 0
        # LambdaLR position: advanced by finetune / distill steps only; pretraining runs at a forced constant lr and leaves
        # the scheduler alone (SealNeRF/trainer.py:431-432, :491-503), so fine-tuning starts the decay from step 0
        self.sched_step, self.lr_forced = 0, False
        self.table8 = None
        # data parallel: optionally (S3D_GRAD_CHUNKS = 2..4) the table gradient is scattered in level chunks and each finished
        # slice is all-reduced (async, NCCL's own stream) under the next chunk.  Measured on 8 B200s: 2 chunks 8.26 ms/step,
        # 3 chunks 8.74, one all-reduce after a single scatter launch 8.36; on 2 GPUs 3 chunks cost 3 % (the extra launches and
        # NCCL's CTAs slow the scatter by about what the overlap hides) -- so the default is the single collective.
        n_chunks = int(os.environ.get("S3D_GRAD_CHUNKS", 1))
        self.grad_chunks = self.S.grad_chunks(n_chunks) if (world_size > 1 and n_chunks > 1) else None
        self._pending = []
        self.fused_forward = os.environ.get("S3D_PAIR_FORWARD", "0") != "0"   # gather + both MLPs in one kernel (k_ngp_pair_fwd)
        self.bwd_chunks = int(os.environ.get("S3D_BWD_CHUNKS", 1))   # MLP backward / scatter overlap (FusedNGP.backward)
        self._side, self._pref = None, None     # side stream + the pre-marched next batch (see _prefetch)
        self._cur = self._old = None            # pre-marched tensors in use by this / the previous step (kept alive, see _prefetch)
        self.lr_decay_iters = lr_decay_iters
        self.scaler = GradScalerState(self.S.dev, **(scaler_kwargs or {})) if loss_scale == "dynamic" else None
        self.ema = ParamEMA(self.S.param_tensors(), ema_decay) if ema_decay is not None else None
        try:   # per-step buffers are sized by the sample budget, which moves a little at every occupancy refresh: let the
            # caching allocator round large requests up (1/16 of a power of two) so refreshed sizes reuse cached blocks
            if "roundup_power2_divisions" not in os.environ.get("PYTORCH_CUDA_ALLOC_CONF", ""):
                setter = getattr(torch._C, "_accelerator_setAllocatorSettings", None) or torch.cuda.memory._set_allocator_settings
                setter("roundup_power2_divisions:16")
        except Exception:
            pass
        if self.T is not None and self.T.N == self.S.N and torch.equal(self.T.offsets, self.S.offsets):
            # teacher and student share the level geometry: pair their fp16 tables so one 128-bit load serves both
            self.table8 = torch.empty(self.S.N, 8, dtype=torch.float16, device=self.S.dev)
            _lib.call("s3d_ngp_pair_tables", self.T.table4, self.S.table4, self.table8, self.S.N)
            self.S.use_paired_table(self.table8, 1)
        self.peer, self._moments_gathered_at = None, None
        if world_size > 1 and self.scaler is None and os.environ.get("S3D_PEER_ADAM", "1") != "0" and self.S.dev.type == "cuda":
            self._setup_peer_step()

    # -- data parallel over NVLink peer memory -----------------------------------------------------------------------------
    # Instead of all_reduce(gradient arena) + a full Adam pass on every rank, every rank owns 1/world of the table entries:
    # one kernel (s3d_ngp_peer_adam_tables) reads the entry's gradient from every rank's arena through peer pointers, sums in
    # rank order, steps Adam with its shard of the moments and writes the new entry + fp16 shadow into every rank's tables.
    # The gradient arena, the two fp32 tables and the fp16 table live in symmetric memory for that (parallel.PeerBuffer); the
    # MLP's 12 K gradients are summed by every rank for itself (s3d_peer_sum) and stepped locally.  Two device-side barriers per
    # step order the peers (after the scatters, after the updates); no NCCL call is left in the step.
    # Static loss scale only (a dynamic scaler needs a cross-rank found-inf decision: that path keeps the all-reduce);
    # S3D_PEER_ADAM=0 or a failing rendezvous also keep the all-reduce.
    def _setup_peer_step(self):
        from .parallel import PeerBuffer, shard_bounds
        S = self.S
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() != self.world_size:
            return
        err = None
        try:
            g = PeerBuffer(S.grad.numel(), torch.float32, S.dev)
            par = PeerBuffer(S.N * 4, torch.float32, S.dev)
            tbl = PeerBuffer(S._tbl.numel(), torch.float16, S.dev)
        except Exception as e:      # no symmetric memory on this system / group: the NCCL path stays
            err = e
        # every rank must take the same path: one that could not set up its buffers sends everybody to the all-reduce
        ok = torch.tensor([0.0 if err is not None else 1.0], device=S.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            if dist.get_rank() == 0:
                print("seal3d_b200: peer-memory optimizer step unavailable (%s); using all_reduce" %
                      ("%s: %s" % (type(err).__name__, err) if err is not None else "another rank could not set it up"), flush=True)
            return
        enc, encc = S.model.encoder, S.model.encoder_color
        g.local.zero_()
        par.local[:S.N * 2].copy_(enc.embeddings.data.reshape(-1))
        par.local[S.N * 2:].copy_(encc.embeddings.data.reshape(-1))
        tbl.local.copy_(S._tbl.reshape(-1))
        # re-home the arena, the parameters and the fp16 table (same values, peer-visible storage)
        S.grad = g.local
        S.grad4, S.gmlp = S.grad[:S.N * 4], S.grad[S.N * 4:]
        enc.embeddings.data = par.local[:S.N * 2].view(S.N, 2)
        encc.embeddings.data = par.local[S.N * 2:].view(S.N, 2)
        new_tbl = tbl.local.view_as(S._tbl)
        if self.table8 is not None:
            self.table8 = new_tbl
        else:
            S.table4 = new_tbl
        S._tbl = new_tbl
        if self.ema is not None:
            self.ema = ParamEMA(S.param_tensors(), self.ema.decay)
        W, rank = g.world, g.rank
        e0, e1 = shard_bounds(S.N, rank, W)
        hp = _lib.host_ptrs
        self.peer = dict(
            g=g, par=par, tbl=tbl, world=W, rank=rank, shard=(e0, e1),
            grad=hp(g.ptrs), sigma=hp(par.ptrs), color=hp([p + S.N * 2 * 4 for p in par.ptrs]), shadow=hp([p + S._tbl_off for p in tbl.ptrs]),
            gmlp=hp([p + S.N * 4 * 4 for p in g.ptrs]), gmlp_sum=torch.zeros(S.n_mlp, dtype=torch.float32, device=S.dev))
        g.barrier(0)

    def _peer_step(self, scale, train_mlp):
        S, P = self.S, self.peer
        lr, gs = self.current_lr(), 1.0 / (self.world_size * scale)
        P["g"].barrier(0)                       # every rank's scatter / weight-gradient flush has finished
        S.step_tables += 1
        _lib.call("s3d_ngp_peer_adam_tables", P["grad"][1], P["sigma"][1], P["color"][1], P["shadow"][1], P["world"], P["rank"], S.m4, S.v4,
                  S._tbl_stride, P["shard"][0], P["shard"][1], float(lr), 0.9, 0.99, 1e-15, S.step_tables, float(gs))
        if train_mlp:
            S.step_mlp += 1
            _lib.call("s3d_peer_sum", P["gmlp"][1], P["world"], P["gmlp_sum"], S.n_mlp)
            _lib.call("s3d_adam_step", S.mlp32, P["gmlp_sum"], S.m_mlp, S.v_mlp, S.mlp16, S.n_mlp, float(lr), 0.9, 0.99, 1e-15, S.step_mlp, float(gs), 1, 0, None)
        else:
            pass                                # (the MLP part of the arena is cleared with the rest below)
        P["g"].barrier(1)                       # every shard is written everywhere, every arena has been read
        S.grad.zero_()
        # raw-pointer parameter updates do not bump torch's version counter: mark the modules' cached fp16 tables stale
        S.model.encoder._shadow_version = S.model.encoder_color._shadow_version = -1
        self._moments_gathered_at = None

    def gather_optimizer_state(self):
        """peer mode keeps each entry's Adam moments on the rank that owns it: bring every rank's copy up to date (checkpoints)"""
        if self.peer is None:
            return
        from .parallel import shard_bounds
        S = self.S
        for r in range(self.peer["world"]):
            a, b = shard_bounds(S.N, r, self.peer["world"])
            dist.broadcast(S.m4[a * 4:b * 4], src=r)
            dist.broadcast(S.v4[a * 4:b * 4], src=r)
        self._moments_gathered_at = self.global_step

    def _march_under_scatter(self):
        """where the next batch is marched (side stream): under the gradient scatter + optimizer kernels on one GPU and in peer
        mode (whose reduce + Adam kernel fills the SMs), under the NCCL all-reduce (a few SMs) on the all-reduce path"""
        return self.world_size == 1 or self.peer is not None

    def _scale(self, units):
        if self.scaler is not None:
            return self.scaler.scale_tensor          # device scalar; divided out inside the Adam kernels
        return float(self.loss_scale) if self.loss_scale is not None else 32.0 * float(units)

    def current_lr(self):
        if not self.lr_decay_iters or self.lr_forced:
            return self.lr
        return self.lr * 0.1 ** min(self.sched_step / float(self.lr_decay_iters), 1.0)

    def _start_reduce(self, i, a0, a1):
        self._pending.append(dist.all_reduce(self.S.grad[a0:a1], async_op=True))

    def _reduce_and_step(self, scale, train_mlp=True, advance_schedule=True):
        if self.peer is not None:
            self._peer_step(scale, train_mlp)
            self.global_step += 1
            if advance_schedule:
                self.sched_step += 1
            return
        if self.world_size > 1:
            if self._pending:              # the chunks of the gradient arena, each started right after its scatter launch
                for w in self._pending:
                    w.wait()
                self._pending = []
            else:
                dist.all_reduce(self.S.grad)   # one collective over the whole arena
        if self.scaler is not None:
            # the check runs on the all-reduced arena, so every rank takes the same skip / backoff decision
            self.scaler.check(self.S.grad, advance_mlp=train_mlp)
            self.S.adam_step(self.current_lr(), grad_scale=1.0 / self.world_size, train_mlp=train_mlp, scaler_state=self.scaler.state,
                             scaler_state_mlp=self.scaler.state_mlp)
            self.scaler.update()
        else:
            self.S.adam_step(self.current_lr(), grad_scale=1.0 / (self.world_size * scale), train_mlp=train_mlp)
        self.global_step += 1
        if advance_schedule:
            self.sched_step += 1

    def ema_update(self):
        if self.ema is not None:
            self.ema.update()

    def ema_apply(self):
        """evaluation with the averaged parameters (nerf/utils.py:919-921): store, copy_to, rebuild the fp16 shadows"""
        self.ema.store()
        self.ema.copy_to()
        self.S.sync_from_module()
        self._refresh_pairing()

    def ema_restore(self):
        self.ema.restore()
        self.S.sync_from_module()
        self._refresh_pairing()

    def _refresh_pairing(self):
        if self.table8 is not None:
            _lib.call("s3d_ngp_pair_tables", self.T.table4, self.S.table4, self.table8, self.S.N)

    def _march(self, rays_o, rays_d, perturb, force_all_rays):
        s = self.student
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, s.aabb_train, s.min_near)
        counter = s.step_counter[s.local_step % 16]
        counter.zero_()
        s.local_step += 1
        return raymarching.march_rays_train(rays_o, rays_d, s.bound, s.density_bitfield, s.cascade, s.grid_size, nears, fars, counter,
                                            s.mean_count, perturb, 128, force_all_rays, self.dt_gamma, self.max_steps)

    # -- software pipelining of the marcher ---------------------------------------------------------------------
    # Marching depends on the rays and the occupancy bitfield only, not on the parameters, so the samples of step n+1 can be
    # generated while step n's gradient scatter and Adam run (one GPU) or while its gradient all-reduce runs (data parallel:
    # NCCL occupies a few SMs, the marcher gets the rest): the marcher is latency-bound (divergent per-ray DDA), the scatter
    # is bound by L2 atomics, and they share the SMs well.  (Under the persistent MLP kernels it does not pay: their static
    # tile schedule turns any SM the marcher delays into the kernel's tail -- measured, profiles/r1d_experiments.md.)  `prefetch=(rays_o, rays_d)` on a step hands the NEXT
    # batch (device tensors or pinned host tensors; host tensors are copied on the side stream too) to a second stream.
    # The pre-marched batch is used when the next call passes the same tensor objects.  Nothing is pre-marched across an
    # occupancy refresh, so every march sees exactly the bitfield it would have seen without pipelining.
    def _prefetch(self, rays, perturb, force_all_rays):
        if rays is None or (self.update_interval and (self.global_step + 1) % self.update_interval == 0):
            return
        ro, rd = rays
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.S.dev)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)        # ordered after everything issued so far (bitfield updates, the previous march)
        # Tensors marched on the side stream live in its allocator pool.  They are kept referenced until the side stream has
        # waited for the whole step that consumed them, then dropped -- so the pool can only hand their memory to a march
        # that is ordered after their last reader, without record_stream (whose deferred frees make a host that runs ahead
        # of the GPU allocate fresh memory every step).
        self._old, self._cur = self._cur, None
        with torch.cuda.stream(self._side):
            o = ro.to(self.S.dev, non_blocking=True).view(-1, 3)
            d = rd.to(self.S.dev, non_blocking=True).view(-1, 3)
            res = self._march(o, d, perturb, force_all_rays)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._pref = (ro, rd, o, d, res, ev)

    def _march_or_take(self, rays_o, rays_d, perturb, force_all_rays):
        """-> device rays_o, rays_d and the march result, from the pre-marched batch if this is the batch that was handed in"""
        pref, self._pref = self._pref, None
        if pref is not None and pref[0] is rays_o and pref[1] is rays_d:
            _, _, o, d, res, ev = pref
            torch.cuda.current_stream().wait_event(ev)
            self._cur = (o, d, res)
            return o, d, res
        if pref is not None:       # a different batch arrived: the pre-marched one is dropped, its ring slot is simply overwritten later
            torch.cuda.current_stream().wait_event(pref[5])
        o = rays_o.to(self.S.dev, non_blocking=True).view(-1, 3)
        d = rays_d.to(self.S.dev, non_blocking=True).view(-1, 3)
        return o, d, self._march(o, d, perturb, force_all_rays)

    def _composite(self, sigmas, rgbs, deltas, rays):
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        ws = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        _lib.call("s3d_composite_rays_train_forward", sigmas, rgbs, deltas, rays, M, N, float(self.T_thresh), ws, depth, image)
        return ws, depth, image

    def _student_backward(self, xyzs, dirs, feats, sig_s, rgb_s, deltas, rays, image_t, depth_t, before_scatter=None, teacher=None):
        """per-ray part (k_distill_rays: student composite, loss, compositor backward -- with teacher=(sig_t, rgb_t) also the
        teacher's composite on the same samples) in one launch, then the field backward.  -> (loss buffer, loss scale)"""
        M, N = sig_s.shape[0], rays.shape[0]
        dev = sig_s.device
        self.loss_buf.zero_()
        scale = self._scale(N)
        g_sig = torch.zeros(M, dtype=torch.float32, device=dev)
        g_rgb = torch.zeros(M, 3, dtype=torch.float32, device=dev)
        sig_t, rgb_t = teacher if teacher is not None else (None, None)
        dyn = isinstance(scale, torch.Tensor)
        _lib.call("s3d_distill_rays", sig_t, rgb_t, image_t, depth_t, sig_s, rgb_s, deltas, rays, M, N, float(self.T_thresh), self.bg_color,
                  1.0 if dyn else float(scale), scale if dyn else None, self.loss_buf, g_sig, g_rgb)
        self.S.backward(xyzs, dirs, feats, g_sig, g_rgb, before_scatter=before_scatter, chunks=self.grad_chunks,
                        after_chunk=self._start_reduce if self.grad_chunks else None, sample_chunks=self.bwd_chunks)
        return self.loss_buf, scale

    def _teacher_field(self, mx, md, mask, feats_t):
        """teacher sigma / rgb on the (proxy-mapped) samples, colour edit applied"""
        t = self.teacher
        sig_t, rgb_t, _ = self.T.mlp_forward(feats_t, md)
        if mask is not None and t.seal_mapper is not None and t.seal_mapper.has_color_edit():
            t.seal_mapper.map_color_(rgb_t, mask, mx)
        return sig_t, rgb_t

    def _teacher_composite(self, mx, md, mask, feats_t, deltas, rays):
        sig_t, rgb_t = self._teacher_field(mx, md, mask, feats_t)
        ws_t, depth_t, img_t = self._composite(sig_t, rgb_t, deltas, rays)
        img_t.add_((1 - ws_t).unsqueeze(-1) * self.bg_color)
        return img_t, depth_t

    def teacher_targets(self, xyzs, dirs, deltas, rays):
        """teacher on the student's samples: proxy map -> field -> colour edit -> composite (+ background)"""
        mx, md, mask = self.teacher._map_samples(xyzs, dirs)
        feats_t = self.T.encode(mx.contiguous().float())
        return self._teacher_composite(mx, md.contiguous().float(), mask, feats_t, deltas, rays)

    @torch.no_grad()
    def distill_step(self, rays_o, rays_d, perturb=True, force_all_rays=False, prefetch=None):
        """prefetch = (rays_o, rays_d) of the NEXT step (optional): marched on a side stream under this step's kernels"""
        self._maybe_update_grid()
        rays_o, rays_d, (xyzs, dirs, deltas, rays) = self._march_or_take(rays_o, rays_d, perturb, force_all_rays)
        ahead = (lambda: self._prefetch(prefetch, perturb, force_all_rays)) if prefetch is not None else None
        if self.table8 is not None:
            # one gather pass for both models (moved samples take a second gather for the teacher)
            M = xyzs.shape[0]
            mx, md, mask = self.teacher._map_samples(xyzs, dirs)
            feats = torch.empty(M, 64, dtype=torch.float16, device=xyzs.device)
            m8 = mask.view(torch.uint8) if mask is not None else None
            if self.fused_forward:
                # gather + both MLPs in one kernel: features go from the gather warps straight into tensor memory
                dev = xyzs.device
                sig_t, sig_s = torch.empty(M, dtype=torch.float32, device=dev), torch.empty(M, dtype=torch.float32, device=dev)
                rgb_t, rgb_s = torch.empty(M, 3, dtype=torch.float32, device=dev), torch.empty(M, 3, dtype=torch.float32, device=dev)
                wt, ws = self.T._w16(), self.S._w16()
                _lib.call("s3d_ngp_pair_forward", xyzs, mx if mask is not None else None, m8, dirs, md if mask is not None else None, M, self.S.bound,
                          self.table8, self.S.offsets, self.S.L, self.S.S, self.S.H, wt[0], wt[1], wt[2], wt[3], wt[4], ws[0], ws[1], ws[2], ws[3], ws[4],
                          self.T.density_scale, self.S.density_scale, sig_t, rgb_t, sig_s, rgb_s, feats)
                t = self.teacher
                if mask is not None and t.seal_mapper is not None and t.seal_mapper.has_color_edit():
                    t.seal_mapper.map_color_(rgb_t, mask, mx)
                teacher_out = (sig_t, rgb_t)
            else:
                feats_t = torch.empty(M, 64, dtype=torch.float16, device=xyzs.device)
                _lib.call("s3d_ngp_encode_pair", xyzs, mx if mask is not None else None, m8, M, self.S.bound, self.table8, self.S.offsets, self.S.L,
                          self.S.S, self.S.H, feats_t, feats)
                teacher_out = self._teacher_field(mx, md, mask, feats_t)
                sig_s, rgb_s, _ = self.S.mlp_forward(feats, dirs)
        else:
            mx, md, mask = self.teacher._map_samples(xyzs, dirs)
            teacher_out = self._teacher_field(mx, md.contiguous().float(), mask, self.T.encode(mx.contiguous().float()))
            sig_s, rgb_s, feats = self.S.forward(xyzs, dirs)
        loss, scale = self._student_backward(xyzs, dirs, feats, sig_s, rgb_s, deltas, rays, None, None,
                                             before_scatter=ahead if self._march_under_scatter() else None, teacher=teacher_out)
        if ahead is not None and not self._march_under_scatter():
            ahead()      # data parallel: the next batch is marched under the gradient all-reduce (NCCL needs only a few SMs)
        self._reduce_and_step(scale)
        return loss

    @torch.no_grad()
    def finetune_step(self, rays_o, rays_d, image_t, depth_t=None, perturb=True, force_all_rays=False, prefetch=None):
        self._maybe_update_grid()
        rays_o, rays_d, (xyzs, dirs, deltas, rays) = self._march_or_take(rays_o, rays_d, perturb, force_all_rays)
        ahead = (lambda: self._prefetch(prefetch, perturb, force_all_rays)) if prefetch is not None else None
        sig_s, rgb_s, feats = self.S.forward(xyzs, dirs)
        loss, scale = self._student_backward(xyzs, dirs, feats, sig_s, rgb_s, deltas, rays, image_t, depth_t,
                                             before_scatter=ahead if self._march_under_scatter() else None)
        if ahead is not None and not self._march_under_scatter():
            ahead()
        self._reduce_and_step(scale)
        return loss

    @torch.no_grad()
    def pretrain_step(self, points, dirs, sigma_t, rgb_t):
        points, dirs = points.contiguous().float(), dirs.contiguous().float()
        sig_s, rgb_s, feats = self.S.forward(points, dirs)
        M = points.shape[0]
        g_s = torch.empty(M, dtype=torch.float32, device=points.device)
        g_c = torch.empty(M, 3, dtype=torch.float32, device=points.device)
        self.loss_buf.zero_()
        _lib.call("s3d_pretrain_loss", sig_s, rgb_s, sigma_t, rgb_t, M, self.loss_buf, g_s, g_c)
        scale = self._scale(M)
        g_s.mul_(scale)
        g_c.mul_(scale)
        self.S.backward(points, dirs, feats, g_s, g_c, train_mlp=False)
        self._reduce_and_step(scale, train_mlp=False, advance_schedule=False)
        return self.loss_buf

    def _maybe_update_grid(self):
        # nerf/utils.py:845-847: `global_step % update_extra_interval == 0`, step 0 included
        if self.update_interval and self.global_step % self.update_interval == 0:
            self.refresh_occupancy()

    @torch.no_grad()
    def refresh_occupancy(self, seed=None):
        seed = 1234 + self.global_step if seed is None else seed
        if self._pref is not None:      # a batch pre-marched on the old bitfield is dropped
            torch.cuda.current_stream().wait_event(self._pref[5])
            self._pref = None
        orig = self.student.density
        self.student.density = self.S.density
        try:
            self.student.update_extra_state(seed=seed)
        finally:
            self.student.density = orig
