"""Checkpoint I/O in the reference's ``.pth`` format (``nerf/utils.py:1015-1136`` save_checkpoint / load_checkpoint;
SURVEY.md 8f-4 and appendix B), so that Seal-3D checkpoints load here and checkpoints written here load in Seal-3D.

A checkpoint is ``torch.save`` of a dict::

    epoch, global_step, stats                       always
    mean_count, mean_density                        cuda-ray models (nerf/utils.py:1026-1028)
    model                                           model.state_dict() with the reference's keys / shapes
    optimizer, lr_scheduler, scaler, ema            only with full=True (:1030-1035)

``optimizer`` is a ``torch.optim.Adam.state_dict()`` (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` in the
parameter order of ``get_params``), built from whichever arena the trainer uses: the flat arena of
``trainer.DistillTrainer`` or the interleaved table moments of ``fused.FusedDistillTrainer`` (de-interleaved here).
``scaler`` follows ``torch.cuda.amp.GradScaler.state_dict()``, ``ema`` follows ``torch_ema``'s, ``lr_scheduler`` follows
``LambdaLR``'s.  Tensors are written contiguous in their logical shape, so the channels_last storage of the TensoRF
factors is invisible in the file.  This is host-side I/O: plain torch copies, no kernels.
"""
import glob
import os

import torch

_ADAM_DEFAULTS = dict(betas=(0.9, 0.99), eps=1e-15, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                      capturable=False, differentiable=False, fused=None)


def model_state_dict(model):
    return {k: v.detach().contiguous().clone() for k, v in model.state_dict().items()}


# ---- per-parameter views of the Adam state ------------------------------------------------------------------------

def _groups(model, lr):
    lrs = lr if isinstance(lr, (tuple, list)) else (lr,)
    return [{"params": list(g["params"]), "lr": g["lr"]} for g in model.get_params(*lrs)]


def _moment_views(trainer):
    """-> {id(param): (step, exp_avg view, exp_avg_sq view)} for every trainable parameter of trainer.student"""
    out = {}
    if hasattr(trainer, "arena"):                     # trainer.DistillTrainer: one flat arena in the parameters' own layout
        a = trainer.arena
        for p in a.params:
            off, k = a.offsets[id(p)]
            perm = sorted(range(p.dim()), key=lambda i: (p.shape[i] != 1, -p.stride(i), i))
            inv = [perm.index(i) for i in range(p.dim())]
            shape = [p.shape[i] for i in perm]
            out[id(p)] = (a.steps.get(id(p), 0), a.exp_avg[off:off + k].view(shape).permute(inv), a.exp_avg_sq[off:off + k].view(shape).permute(inv))
        return out
    S = trainer.S                                     # fused.FusedDistillTrainer: [N,4] interleaved table moments + MLP arena
    m4, v4 = S.m4.view(S.N, 4), S.v4.view(S.N, 4)
    st_t, st_m = S.step_tables, S.step_mlp
    if getattr(trainer, "scaler", None) is not None:  # dynamic loss scale: the host counters also count skipped steps
        st_t, st_m = trainer.scaler.steps()
    out[id(S.model.encoder.embeddings)] = (st_t, m4[:, 0:2], v4[:, 0:2])
    out[id(S.model.encoder_color.embeddings)] = (st_t, m4[:, 2:4], v4[:, 2:4])
    for w, (off, k) in zip(S.weights, S._w_off):
        out[id(w)] = (st_m, S.m_mlp[off:off + k].view_as(w), S.v_mlp[off:off + k].view_as(w))
    return out


def _set_steps(trainer, steps):
    if hasattr(trainer, "arena"):
        trainer.arena.steps = dict(steps)
        return
    S = trainer.S
    S.step_tables = steps.get(id(S.model.encoder.embeddings), 0)
    S.step_mlp = steps.get(id(S.weights[0]), 0)


def optimizer_state_dict(trainer):
    groups = _groups(trainer.student, trainer.lr)
    views = _moment_views(trainer)
    lr_now = trainer.current_lr() if hasattr(trainer, "current_lr") else trainer.lr
    state, pgs, idx = {}, [], 0
    for g in groups:
        ids = []
        for p in g["params"]:
            step, m, v = views[id(p)]
            if step > 0:                              # torch creates a parameter's state at its first step
                state[idx] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().contiguous().clone(),
                              "exp_avg_sq": v.detach().contiguous().clone()}
            ids.append(idx)
            idx += 1
        base = float(g["lr"])
        scale = (float(lr_now) / float(trainer.lr)) if not isinstance(trainer.lr, (tuple, list)) and trainer.lr else 1.0
        pgs.append(dict(lr=base * scale, initial_lr=base, params=ids, **_ADAM_DEFAULTS))
    return {"state": state, "param_groups": pgs}


def load_optimizer_state_dict(trainer, sd):
    groups = _groups(trainer.student, trainer.lr)
    params = [p for g in groups for p in g["params"]]
    views = _moment_views(trainer)
    steps = {}
    for idx, p in enumerate(params):
        st = sd["state"].get(idx, sd["state"].get(str(idx)))
        _, m, v = views[id(p)]
        if st is None:
            m.zero_()
            v.zero_()
            continue
        m.copy_(st["exp_avg"].to(m.device))
        v.copy_(st["exp_avg_sq"].to(v.device))
        steps[id(p)] = int(float(st["step"]))
    _set_steps(trainer, steps)


def _scheduler_state_dict(trainer):
    step = int(getattr(trainer, "sched_step", trainer.global_step))     # LambdaLR counts train steps only, not pretraining
    base = list(trainer.lr) if isinstance(trainer.lr, (tuple, list)) else [trainer.lr]
    now = trainer.current_lr() if hasattr(trainer, "current_lr") else trainer.lr
    last = [now] if not isinstance(now, (tuple, list)) else list(now)
    return {"base_lrs": base, "last_epoch": step, "_step_count": step + 1, "_get_lr_called_within_step": False,
            "_last_lr": last, "lr_lambdas": [None] * len(base)}


def _scaler_state_dict(scaler):
    st = scaler.state.detach().cpu()
    return {"scale": float(st[0]), "growth_factor": scaler.growth_factor, "backoff_factor": scaler.backoff_factor,
            "growth_interval": scaler.growth_interval, "_growth_tracker": int(st[1])}


def _ema_param_lists(trainer):
    """the EMA tensors re-expressed per model parameter, in model.parameters() order (what torch_ema stores)"""
    S = trainer.S
    sh_s, sh_c, sh_mlp = trainer.ema.shadow
    by_id = {id(S.model.encoder.embeddings): sh_s, id(S.model.encoder_color.embeddings): sh_c}
    for w, (off, k) in zip(S.weights, S._w_off):
        by_id[id(w)] = sh_mlp[off:off + k].view_as(w)
    return [by_id[id(p)] for p in S.model.parameters() if id(p) in by_id]


# ---- the two reference entry points -------------------------------------------------------------------------------

def save_checkpoint(path, trainer=None, model=None, epoch=0, stats=None, full=False):
    """nerf/utils.py:1015-1052 (the non-'best' branch).  Pass a trainer (model = trainer.student) or just a model."""
    model = model if model is not None else trainer.student
    state = {"epoch": int(epoch), "global_step": int(getattr(trainer, "global_step", 0)),
             "stats": stats if stats is not None else {"loss": [], "valid_loss": [], "results": [], "checkpoints": [], "best_result": None}}
    if getattr(model, "cuda_ray", False):
        state["mean_count"] = model.mean_count
        state["mean_density"] = model.mean_density
    if full and trainer is not None:
        if getattr(trainer, "peer", None) is not None and getattr(trainer, "_moments_gathered_at", None) != trainer.global_step:
            # peer-memory data parallel keeps each entry's Adam moments on its owner rank only
            raise RuntimeError("call trainer.gather_optimizer_state() on EVERY rank before saving a full checkpoint of a peer-mode trainer")
        state["optimizer"] = optimizer_state_dict(trainer)
        state["lr_scheduler"] = _scheduler_state_dict(trainer)
        if getattr(trainer, "scaler", None) is not None:
            state["scaler"] = _scaler_state_dict(trainer.scaler)
        if getattr(trainer, "ema", None) is not None:
            state["ema"] = {"decay": trainer.ema.decay, "num_updates": trainer.ema.num_updates,
                            "shadow_params": [t.detach().contiguous().clone() for t in _ema_param_lists(trainer)], "collected_params": None}
    state["model"] = model_state_dict(model)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(state, path)
    return state


def latest_checkpoint(ckpt_dir, name):
    """nerf/utils.py:1071-1073"""
    found = sorted(glob.glob(os.path.join(ckpt_dir, "%s_ep*.pth" % name)))
    return found[-1] if found else None


def load_checkpoint(path, trainer=None, model=None, model_only=False, map_location=None):
    """nerf/utils.py:1068-1136.  Returns (missing_keys, unexpected_keys, checkpoint dict)."""
    model = model if model is not None else trainer.student
    dev = next(model.parameters()).device
    ck = torch.load(path, map_location=map_location or dev, weights_only=False)
    if "model" not in ck:                                        # a bare state dict (:1082-1085)
        res = model.load_state_dict(ck)
        _after_model_load(trainer)
        return list(res.missing_keys), list(res.unexpected_keys), ck
    res = model.load_state_dict(ck["model"], strict=False)
    if getattr(model, "cuda_ray", False):
        model.mean_count = ck.get("mean_count", model.mean_count)
        model.mean_density = ck.get("mean_density", model.mean_density)
    _after_model_load(trainer)
    if trainer is not None and getattr(trainer, "ema", None) is not None and "ema" in ck:
        trainer.ema.decay, trainer.ema.num_updates = float(ck["ema"]["decay"]), int(ck["ema"]["num_updates"])
        for dst, src in zip(_ema_param_lists(trainer), ck["ema"]["shadow_params"]):
            dst.copy_(src.to(dst.device))
    if model_only or trainer is None:
        return list(res.missing_keys), list(res.unexpected_keys), ck
    trainer.global_step = int(ck.get("global_step", 0))
    if hasattr(trainer, "sched_step"):
        trainer.sched_step = int(ck.get("lr_scheduler", {}).get("last_epoch", trainer.global_step))
    if "optimizer" in ck:
        load_optimizer_state_dict(trainer, ck["optimizer"])
    if getattr(trainer, "scaler", None) is not None and "scaler" in ck:
        trainer.scaler.state[0] = float(ck["scaler"]["scale"])
        trainer.scaler.state_mlp[0] = float(ck["scaler"]["scale"])
        trainer.scaler.state[1] = float(ck["scaler"].get("_growth_tracker", 0))
    if getattr(trainer, "scaler", None) is not None and hasattr(trainer, "S"):
        # with a device-side scaler the Adam kernels take their bias corrections from its applied-step counts
        # (skipped steps excluded), not from the host counters: restore them from the optimizer entry
        trainer.scaler.set_steps(trainer.S.step_tables, trainer.S.step_mlp)
    return list(res.missing_keys), list(res.unexpected_keys), ck


def _after_model_load(trainer):
    """the fused trainer keeps fp16 forms of the parameters: rebuild them from the freshly loaded masters"""
    if trainer is None or not hasattr(trainer, "S"):
        return
    trainer.S.sync_from_module()
    trainer._refresh_pairing()
