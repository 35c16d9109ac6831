"""TensoRF VM lookup kernels (csrc/tensorf.cu) on the B200: time per launch, algorithmic GB/s, and the reference's own
formulation (tensoRF/network.py:115-158: twelve F.grid_sample calls on channel-major [1,R,H,W] images) on the same GPU
and the same inputs.  Development / measurement tool; prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def timed(fn, n=8, flush=None):
    ts = []
    for i in range(n + 3):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))


def ref_lookup(x, mats, vecs, reduce):
    """the reference's get_sigma_feat / get_color_feat body on NCHW-contiguous factor images"""
    N = x.shape[0]
    mat_ids, vec_ids = [[0, 1], [0, 2], [1, 2]], [2, 1, 0]
    mat_coord = torch.stack((x[..., mat_ids[0]], x[..., mat_ids[1]], x[..., mat_ids[2]])).view(3, -1, 1, 2)
    vec_coord = torch.stack((x[..., vec_ids[0]], x[..., vec_ids[1]], x[..., vec_ids[2]]))
    vec_coord = torch.stack((torch.zeros_like(vec_coord), vec_coord), dim=-1).view(3, -1, 1, 2)
    if reduce:
        out = torch.zeros([N], device=x.device)
        for i in range(3):
            mf = F.grid_sample(mats[i], mat_coord[[i]], align_corners=True).view(-1, N)
            vf = F.grid_sample(vecs[i], vec_coord[[i]], align_corners=True).view(-1, N)
            out = out + torch.sum(mf * vf, dim=0)
        return out
    mf, vf = [], []
    for i in range(3):
        mf.append(F.grid_sample(mats[i], mat_coord[[i]], align_corners=True).view(-1, N))
        vf.append(F.grid_sample(vecs[i], vec_coord[[i]], align_corners=True).view(-1, N))
    return (torch.cat(mf, 0) * torch.cat(vf, 0)).T


def main():
    from seal3d_b200 import synth, raymarching as rm
    from seal3d_b200.tensorf import TensoRFNetwork
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    res = int(os.environ.get("VM_RES", 300))
    net = TensoRFNetwork(resolution=[res] * 3, bound=1).to(dev)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # ray-ordered samples of the bench scene (what training feeds the field) and uniformly random points
    bits, _ = synth.lego_like_occupancy()
    o, d = synth.rays_for_step(0, 1 << 18)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    nears, fars = rm.near_far_from_aabb(to(o), to(d), net.aabb_train, 0.2)
    xyzs, _, _, _ = rm.march_rays_train(to(o), to(d), 1.0, to(bits), 1, 128, nears, fars, None, -1, True, 128, True)
    M = 1 << 22
    sets = {"ray-ordered": xyzs[:M].contiguous(), "random": (torch.rand(M, 3, device=dev) * 2 - 1)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, x in sets.items():
        M = x.shape[0]
        for which, reduce, R in (("sigma", True, 16), ("color", False, 48)):
            mats, vecs = list(getattr(net, which + "_mat")), list(getattr(net, which + "_vec"))
            nchw = [m.detach().contiguous() for m in mats], [v.detach().contiguous() for v in vecs]
            with torch.no_grad():
                ours = net._lookup(x, mats, vecs, reduce)
                ref = ref_lookup(x, nchw[0], nchw[1], reduce)
                err = float((ours - ref).abs().max() / ref.abs().max())
                t_f = timed(lambda: net._lookup(x, mats, vecs, reduce), flush=flush)
                t_rf = timed(lambda: ref_lookup(x, nchw[0], nchw[1], reduce), flush=flush)
            g = torch.randn_like(ours)

            def bwd_ours():
                out = net._lookup(x, mats, vecs, reduce)
                out.backward(g)
            pr = [p.clone().requires_grad_() for p in nchw[0] + nchw[1]]

            def bwd_ref():
                out = ref_lookup(x, pr[:3], pr[3:], reduce)
                out.backward(g)
            t_fb = timed(bwd_ours, flush=flush)
            t_rfb = timed(bwd_ref, flush=flush)
            fwd_bytes = 12 + 3 * 6 * R * 4 + (4 if reduce else 3 * R * 4)
            bwd_bytes = 12 + 2 * 3 * 6 * R * 4 + (4 if reduce else 3 * R * 4)     # re-read both factors + one RED per tap
            print(json.dumps({"points": name, "field": which, "M": M, "res": res, "rel_err_vs_grid_sample": err,
                              "fwd_ms": round(t_f, 4), "fwd_GBps": round(M * fwd_bytes / t_f / 1e6, 1), "fwd_frac_hbm": round(M * fwd_bytes / t_f / 1e6 / peak, 3),
                              "fwd+bwd_ms": round(t_fb, 4), "bwd_ms": round(t_fb - t_f, 4), "bwd_GBps": round(M * bwd_bytes / max(t_fb - t_f, 1e-6) / 1e6, 1),
                              "grid_sample_fwd_ms": round(t_rf, 4), "grid_sample_fwd+bwd_ms": round(t_rfb, 4),
                              "speedup_fwd": round(t_rf / t_f, 2), "speedup_fwd+bwd": round(t_rfb / t_fb, 2)}), flush=True)


if __name__ == "__main__":
    main()
