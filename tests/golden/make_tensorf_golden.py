"""Generate tests/golden/cpu_tensorf.npz from the reference's OWN TensoRF network code run on CPU torch
(needs /root/reference; the output is committed because the reference tree does not exist on the GPU box).

``tensoRF/network.py`` cannot be imported as a module here (its imports pull the CUDA-only ``raymarching`` /
``freqencoder`` extensions and ``nerf.renderer``'s trimesh), so the class ``NeRFNetwork`` is lifted from the
reference source with ``ast`` and exec'd unmodified against three stand-ins:
  * ``NeRFRenderer``  -- a stub base that only stores ``bound`` / ``bg_radius`` / ``aabb_train`` / density-grid state
                         (what ``nerf/renderer.py:59-101`` registers and the VM methods read);
  * ``get_encoder('frequency', ...)`` -- the reference's own pure-torch ``encoding.FreqEncoder`` (encoding.py:5-43) with
                         ``max_freq_log2 = multires - 1, N_freqs = multires``, the equivalence the reference states itself at
                         encoding.py:55; same column layout as the CUDA ``freqencoder`` ([x | sin, cos per octave]);
  * ``raymarching.morton3D_invert`` -- a bit de-interleave in plain torch (only ``shrink_model`` uses it).
Nothing from the reference is copied into this repository.

Run:  python tests/golden/make_tensorf_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_cpu_golden import lift, REF, OUT  # noqa: E402


class _RendererStub(torch.nn.Module):
    def __init__(self, bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1, **kw):
        super().__init__()
        self.bound, self.cascade, self.grid_size = bound, 1, 128
        self.density_scale, self.min_near, self.density_thresh, self.bg_radius, self.cuda_ray = density_scale, min_near, density_thresh, bg_radius, cuda_ray
        aabb = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer("aabb_train", aabb.clone())
        self.register_buffer("aabb_infer", aabb.clone())
        self.register_buffer("density_grid", torch.zeros(1, 128 ** 3))
        self.mean_density = 0


def _morton3D_invert(idx):
    def compact(v):
        v = v & 0x49249249
        v = (v ^ (v >> 2)) & 0xC30C30C3
        v = (v ^ (v >> 4)) & 0x0F00F00F
        v = (v ^ (v >> 8)) & 0xFF0000FF
        v = (v ^ (v >> 16)) & 0x0000FFFF
        return v
    idx = idx.reshape(-1).long()
    return torch.stack([compact(idx), compact(idx >> 1), compact(idx >> 2)], -1).int()


def main():
    g = torch.Generator().manual_seed(4321)
    enc = types.ModuleType("encoding_ref")
    src = open(os.path.join(REF, "encoding.py")).read()
    exec(compile(src, "encoding.py", "exec"), enc.__dict__)

    def get_encoder(encoding, input_dim=3, multires=6, **kw):
        assert encoding == "frequency"
        e = enc.FreqEncoder(input_dim=input_dim, max_freq_log2=multires - 1, N_freqs=multires, log_sampling=True)
        return e, e.output_dim

    act = types.ModuleType("activation")
    exec(compile(open(os.path.join(REF, "activation.py")).read(), "activation.py", "exec"), act.__dict__)
    rm = types.SimpleNamespace(morton3D_invert=_morton3D_invert)
    ns = {"torch": torch, "nn": torch.nn, "F": torch.nn.functional, "np": np, "get_encoder": get_encoder,
          "trunc_exp": act.trunc_exp, "NeRFRenderer": _RendererStub, "raymarching": rm, "print": lambda *a, **k: None}
    lift(os.path.join(REF, "tensoRF", "network.py"), {"NeRFNetwork"}, ns)
    Net = ns["NeRFNetwork"]
    Net._self = Net

    torch.manual_seed(99)
    res = [12, 14, 16]
    net = Net(resolution=res, bound=1)
    out = {"resolution": np.array(res, np.int32)}

    def params(prefix, m):
        for i in range(3):
            out["%ssigma_mat%d" % (prefix, i)] = m.sigma_mat[i].detach().numpy().copy()
            out["%ssigma_vec%d" % (prefix, i)] = m.sigma_vec[i].detach().numpy().copy()
            out["%scolor_mat%d" % (prefix, i)] = m.color_mat[i].detach().numpy().copy()
            out["%scolor_vec%d" % (prefix, i)] = m.color_vec[i].detach().numpy().copy()

    params("", net)
    out["basis_mat"] = net.basis_mat.weight.detach().numpy().copy()
    for l in range(3):
        out["color_net%d" % l] = net.color_net[l].weight.detach().numpy().copy()

    M = 512
    x = torch.rand(M, 3, generator=g) * 2.1 - 1.05          # a few points outside the box: zero padding of grid_sample
    x[:4] = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [1.0, -1.0, 0.5]])   # exact borders / centre
    d = torch.randn(M, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    for tag, aabb in (("", None), ("shrunk_", torch.tensor([-0.8, -0.7, -0.9, 0.9, 0.85, 0.6]))):
        if aabb is not None:
            net.aabb_train.copy_(aabb)
        net.zero_grad()
        xn = 2 * (x - net.aabb_train[:3]) / (net.aabb_train[3:] - net.aabb_train[:3]) - 1
        out[tag + "sigma_feat"] = net.get_sigma_feat(xn).detach().numpy()
        out[tag + "color_feat"] = net.get_color_feat(xn).detach().numpy()
        sigma, rgb = net(x, d)
        out[tag + "sigma"], out[tag + "rgb"] = sigma.detach().numpy(), rgb.detach().numpy()
        assert torch.allclose(net.density(x)["sigma"], sigma)
        msk = torch.zeros(M, dtype=torch.bool)
        msk[::3] = True
        out[tag + "color_masked"] = net.color(x, d, mask=msk).detach().numpy()
        ws, wc = torch.randn(M, generator=g) * 0.1, torch.randn(M, 3, generator=g)
        out[tag + "g_sigma"], out[tag + "g_rgb"] = ws.numpy(), wc.numpy()
        ((sigma * ws).sum() + (rgb * wc).sum()).backward()
        for i in range(3):
            out["%sgrad_sigma_mat%d" % (tag, i)] = net.sigma_mat[i].grad.numpy().copy()
            out["%sgrad_sigma_vec%d" % (tag, i)] = net.sigma_vec[i].grad.numpy().copy()
            out["%sgrad_color_mat%d" % (tag, i)] = net.color_mat[i].grad.numpy().copy()
            out["%sgrad_color_vec%d" % (tag, i)] = net.color_vec[i].grad.numpy().copy()
        out[tag + "grad_basis_mat"] = net.basis_mat.weight.grad.numpy().copy()
        for l in range(3):
            out["%sgrad_color_net%d" % (tag, l)] = net.color_net[l].weight.grad.numpy().copy()
    out["aabb_shrunk"] = net.aabb_train.numpy().copy()
    out["x"], out["d"] = x.numpy(), d.numpy()
    out["density_loss"] = np.float32(net.density_loss().item())

    # upsample_model (network.py:263-277): F.interpolate(bilinear, align_corners=True) of every plane / line
    up = [18, 15, 20]
    net.upsample_model(up)
    out["up_resolution"] = np.array(up, np.int32)
    params("up_", net)

    # shrink_model (network.py:279-317) on a synthetic occupancy: a box of occupied cells in the 128^3 density grid
    torch.manual_seed(7)
    net2 = Net(resolution=res, bound=1)
    params("pre_shrink_", net2)
    occ = torch.zeros(128, 128, 128)
    occ[30:90, 40:100, 20:110] = 50.0       # x, y, z cell ranges
    coords = torch.nonzero(occ > 0)
    xs, ys, zs = coords[:, 0].long(), coords[:, 1].long(), coords[:, 2].long()

    def expand(v):
        v = (v * 0x00010001) & 0xFF0000FF
        v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3
        v = (v * 0x00000005) & 0x49249249
        return v
    midx = expand(xs) | (expand(ys) << 1) | (expand(zs) << 2)
    net2.density_grid.zero_()
    net2.density_grid[0, midx] = 50.0
    net2.density_thresh, net2.mean_density = 10.0, 20.0
    out["shrink_density_grid_cells"] = np.array([[30, 90], [40, 100], [20, 110]], np.int32)
    net2.shrink_model()
    params("shrink_", net2)
    out["shrink_aabb"] = net2.aabb_train.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "cpu_tensorf.npz"), **out)
    print("wrote cpu_tensorf.npz:", len(out), "arrays,", os.path.getsize(os.path.join(OUT, "cpu_tensorf.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
