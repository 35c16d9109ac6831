"""Test / measurement infrastructure (never imported by the product): the reference's own Python wrappers (staged unmodified
under the git-ignored oracle/_ref/py/ by oracle/build_ref.py) loaded over a chosen backend -- the reference's compiled
extensions (oracle/_ref/_ref_*.so) or this repo's drop-in modules -- and the reference's `-O` (fp16 autocast) NGP field and
distillation step composed from them.  Used by tests/test_gpu_reference_wrappers.py and scripts/ref_ab.py."""
import importlib
import importlib.util
import os
import sys
import types

import torch

import refext

WRAP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "py")


def _ours(name):
    import seal3d_b200
    d = os.path.dirname(seal3d_b200.__file__)
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module("_" + name)


def wrapper(pkg, backend, alias):
    """import the staged reference package `pkg` with `import _<pkg> as _backend` resolving to `backend`"""
    init = os.path.join(WRAP, pkg, "__init__.py")
    if not os.path.exists(init):
        import pytest
        pytest.skip("reference wrappers not staged (oracle/build_ref.py wrappers)")
    if "turtle" not in sys.modules:      # ffmlp/ffmlp.py:2 imports two unused names from turtle (needs tkinter)
        sys.modules["turtle"] = types.SimpleNamespace(backward=None, forward=None)
    saved = sys.modules.get("_" + pkg)
    sys.modules["_" + pkg] = backend
    try:
        spec = importlib.util.spec_from_file_location(alias, init, submodule_search_locations=[os.path.join(WRAP, pkg)])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["_" + pkg] = saved
        else:
            sys.modules.pop("_" + pkg, None)
    return mod


def both(pkg):
    return wrapper(pkg, refext.load(pkg), "refwrap_" + pkg), wrapper(pkg, _ours(pkg), "ourwrap_" + pkg)


class _RefTruncExp(torch.autograd.Function):
    """activation.py:5-17 restated (exp forward in fp32, backward g * exp(clamp(x, -15, 15)))"""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class RefAmpField(torch.nn.Module):
    """nerf/network.py:99-128 composed from the reference's own pieces: its GridEncoder / SHEncoder wrappers over its own
    compiled extensions (oracle/_ref) and torch nn.Linear, to be run under torch.autocast like the reference's `-O`"""

    def __init__(self, grid_pkg, sh_pkg, fp):
        super().__init__()
        self.encoder = grid_pkg.GridEncoder(desired_resolution=2048)
        self.encoder_color = grid_pkg.GridEncoder(desired_resolution=2048)
        self.encoder_dir = sh_pkg.SHEncoder(degree=4)
        L = lambda i, o: torch.nn.Linear(i, o, bias=False)
        self.sigma_net = torch.nn.ModuleList([L(32, 64), L(64, 16)])
        self.color_net = torch.nn.ModuleList([L(63, 64), L(64, 64), L(64, 3)])
        self.encoder.embeddings.data.copy_(torch.from_numpy(fp["emb_sigma"]))
        self.encoder_color.embeddings.data.copy_(torch.from_numpy(fp["emb_color"]))
        for lin, k in ((self.sigma_net[0], "w_s0"), (self.sigma_net[1], "w_s1"), (self.color_net[0], "w_c0"), (self.color_net[1], "w_c1"), (self.color_net[2], "w_c2")):
            lin.weight.data.copy_(torch.from_numpy(fp[k]))

    def forward(self, x, d):
        h = self.encoder(x, bound=1)
        h = torch.relu(self.sigma_net[0](h))
        h = self.sigma_net[1](h)
        sigma = _RefTruncExp.apply(h[..., 0])
        geo = h[..., 1:]
        d = self.encoder_dir(d)
        h = torch.cat([d, geo, self.encoder_color(x, bound=1)], dim=-1)
        h = torch.relu(self.color_net[0](h))
        h = torch.relu(self.color_net[1](h))
        return sigma, torch.sigmoid(self.color_net[2](h))




class RefDistillStep:
    """The benchmark's teacher->student distillation step the way the REFERENCE's code would run it on this GPU: its
    extensions (oracle/_ref) through its own wrappers, torch autocast nn.Linear (cuBLAS), torch autograd, GradScaler and
    torch.optim.Adam -- nerf/renderer.py:256-377 run_cuda (training branch), nerf/network.py:99-128, nerf/utils.py:484-489, 857-862.
    The teacher's proxy mapping (SealNeRF/seal_utils.py, pure torch in the reference, not importable here) is delegated to
    `map_samples(xyzs, dirs) -> (xyzs', dirs', mask)` -- this repo's kernel, i.e. the reference arm is credited with it."""

    def __init__(self, teacher_fp, student_fp, bitfield, device, lr=1e-2, map_samples=None, T_thresh=1e-4):
        RG, _ = both("gridencoder")
        RS, _ = both("shencoder")
        self.R, _ = both("raymarching")
        self.dev = device
        self.teacher = RefAmpField(RG, RS, teacher_fp).to(device).eval()
        for p in self.teacher.parameters():
            p.requires_grad_(False)
        self.student = RefAmpField(RG, RS, student_fp).to(device)
        self.bits = bitfield
        self.aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=device)
        self.opt = torch.optim.Adam(self.student.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)
        self.scaler = torch.amp.GradScaler("cuda")
        self.counter = torch.zeros(2, dtype=torch.int32, device=device)
        self.map_samples, self.T = map_samples, T_thresh

    def step(self, rays_o, rays_d, mean_count, perturb=True):
        R = self.R
        nears, fars = R.near_far_from_aabb(rays_o, rays_d, self.aabb, 0.2)
        self.counter.zero_()
        xyzs, dirs, deltas, rays = R.march_rays_train(rays_o, rays_d, 1.0, self.bits, 1, 128, nears, fars, self.counter, mean_count, perturb, 128, False)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            mx, md = xyzs, dirs
            if self.map_samples is not None:
                mx, md, _ = self.map_samples(xyzs, dirs)
            sig_t, rgb_t = self.teacher(mx, md)
            ws_t, dep_t, img_t = R.composite_rays_train(sig_t, rgb_t, deltas, rays, self.T)
            img_t = img_t + (1 - ws_t).unsqueeze(-1)
        with torch.autocast("cuda", dtype=torch.float16):
            sig, rgb = self.student(xyzs, dirs)
            ws, dep, img = R.composite_rays_train(sig, rgb, deltas, rays, self.T)
            img = img + (1 - ws).unsqueeze(-1)
            loss = ((img - img_t) ** 2).mean(-1).mean() + (dep - dep_t).abs().mean()      # nerf/utils.py:484-489, 530
        self.opt.zero_grad(set_to_none=True)
        self.scaler.scale(loss).backward()
        self.scaler.step(self.opt)
        self.scaler.update()
        return loss.detach()
