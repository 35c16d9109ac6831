"""Generate tests/golden/gpu_ref.npz: outputs of the UNMODIFIED reference extensions (oracle/_ref/_ref_*.so, built from
/root/reference/*/src by oracle/build_ref.py) on the seeded inputs of gpu_inputs.py.  Needs a GPU:

    gpurun -- 'python tests/golden/make_gpu_golden.py gpurun_out/gpu_ref.npz'      then copy the file to tests/golden/

These vectors pin the CPU oracle against the reference's own kernels inside the `-m "not gpu"` suite
(tests/test_oracle_golden.py) and are a second, box-independent witness for the `-m gpu` parity tests."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gpu_inputs as gi  # noqa: E402

REF = os.path.join(gi.ROOT, "oracle", "_ref")


def load(name):
    spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, "_ref_%s.so" % name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(out_path):
    dev = torch.device("cuda:0")
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    npy = lambda t: t.detach().cpu().numpy()
    out = {}
    rm = load("raymarching")
    sc = gi.scene()
    N = gi.N_RAYS
    o, d, bits, noises = sc["o"], sc["d"], sc["bits"], sc["noises"]
    # ---- near / far, march (train), canonical per-ray view ---------------------------------------------------------
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    rm.near_far_from_aabb(to(o), to(d), to(gi.AABB), N, 0.2, nears, fars)
    out["nears"], out["fars"] = npy(nears), npy(fars)
    for tag, nz in (("", np.zeros(N, np.float32)), ("perturb_", noises)):
        M = N * 64
        xyzs, dirs, deltas = (torch.zeros(M, k, device=dev) for k in (3, 3, 2))
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        rm.march_rays_train(to(o), to(d), to(bits), 1.0, 0.0, 1024, N, 1, 128, M, nears, fars, xyzs, dirs, deltas, rays, counter, to(nz))
        r = npy(rays)
        order = np.argsort(r[:, 0], kind="stable")
        r = r[order]
        assert np.array_equal(r[:, 0], np.arange(N)) and int(npy(counter)[0]) <= M
        xs, ds, ls = npy(xyzs), npy(dirs), npy(deltas)
        # ray-major concatenation (the reference's slot order depends on atomic order; per-ray content does not)
        out[tag + "march_counts"] = r[:, 2].copy()
        out[tag + "march_xyzs"] = np.concatenate([xs[a:a + k] for _, a, k in r])
        out[tag + "march_deltas"] = np.concatenate([ls[a:a + k] for _, a, k in r])
        out[tag + "march_total"] = npy(counter).copy()
        if tag == "":
            xyz_rm, dir_rm, del_rm = out["march_xyzs"], np.concatenate([ds[a:a + k] for _, a, k in r]), out["march_deltas"]
            counts = r[:, 2].copy()
    # ---- composite (train) forward / backward on the ray-major samples ----------------------------------------------
    Mtot = int(counts.sum())
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    rays_rm = np.stack([np.arange(N, dtype=np.int32), offs, counts.astype(np.int32)], 1)
    fv = gi.field_values(Mtot, N)
    for T in (1e-4, 0.0):
        ws, dp, im = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
        rm.composite_rays_train_forward(to(fv["sigmas"]), to(fv["rgbs"]), to(del_rm), to(rays_rm), Mtot, N, T, ws, dp, im)
        gs, gc = torch.zeros(Mtot, device=dev), torch.zeros(Mtot, 3, device=dev)
        rm.composite_rays_train_backward(to(fv["g_ws"]), to(fv["g_img"]), to(fv["sigmas"]), to(fv["rgbs"]), to(del_rm), to(rays_rm), ws, im, Mtot, N, T, gs, gc)
        k = "T%g_" % T
        out[k + "ws"], out[k + "depth"], out[k + "image"], out[k + "g_sigmas"], out[k + "g_rgbs"] = npy(ws), npy(dp), npy(im), npy(gs), npy(gc)
    # ---- inference march + composite, first iteration of the eval loop (n_step = 8) ----------------------------------
    n_step = 8
    alive = torch.arange(N, dtype=torch.int32, device=dev)
    rays_t = nears.clone()
    Mi = N * n_step
    Mi += 128 - Mi % 128
    xi, di, li = (torch.zeros(Mi, k, device=dev) for k in (3, 3, 2))
    rm.march_rays(N, n_step, alive, rays_t, to(o), to(d), 1.0, 0.0, 1024, 1, 128, to(bits), nears, fars, xi, di, li, to(np.zeros(N, np.float32)))
    out["infer_xyzs"], out["infer_deltas"] = npy(xi), npy(li)
    rng = np.random.default_rng(2)
    si, ci = rng.uniform(0, 60, Mi).astype(np.float32), rng.uniform(0, 1, (Mi, 3)).astype(np.float32)
    ws, dp, im = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
    rm.composite_rays(N, n_step, 1e-2, alive, rays_t, to(si), to(ci), li, ws, dp, im)
    out["infer_alive"], out["infer_rays_t"], out["infer_ws"], out["infer_depth"], out["infer_image"] = npy(alive), npy(rays_t), npy(ws), npy(dp), npy(im)
    # ---- morton / packbits ---------------------------------------------------------------------------------------------
    coords, grid = gi.morton_inputs()
    idx = torch.empty(coords.shape[0], dtype=torch.int32, device=dev)
    rm.morton3D(to(coords), coords.shape[0], idx)
    back = torch.empty(coords.shape[0], 3, dtype=torch.int32, device=dev)
    rm.morton3D_invert(idx, coords.shape[0], back)
    bf = torch.empty(grid.size // 8, dtype=torch.uint8, device=dev)
    rm.packbits(to(grid), grid.size // 8, 10.0, bf)
    out["morton"], out["morton_invert"], out["packbits"] = npy(idx), npy(back), npy(bf)
    # ---- grid encoder ------------------------------------------------------------------------------------------------------
    ge = load("gridencoder")
    offsets, pls, emb, x, g = gi.grid_inputs()
    B, S = x.shape[0], float(np.log2(pls))
    o32 = torch.empty(16, B, 2, device=dev)
    dy = torch.empty(B, 16 * 3 * 2, device=dev)
    ge.grid_encode_forward(to(x), to(emb), to(offsets), o32, B, 3, 2, 16, S, 16, dy, 0, False, 0)
    o16 = torch.empty(16, B, 2, device=dev, dtype=torch.float16)
    ge.grid_encode_forward(to(x), to(emb).half(), to(offsets), o16, B, 3, 2, 16, S, 16, None, 0, False, 0)
    out["grid_fwd_f32"], out["grid_dy_dx"], out["grid_fwd_f16"] = npy(o32), npy(dy), npy(o16.float())
    gemb = torch.zeros(int(offsets[-1]), 2, device=dev)
    ge.grid_encode_backward(to(g), to(x[:128]), to(emb), to(offsets), gemb, 128, 3, 2, 16, S, 16, None, None, 0, False, 0)
    gnp = npy(gemb)
    nz = np.nonzero(np.abs(gnp).sum(1))[0].astype(np.int32)
    out["grid_bwd_rows"], out["grid_bwd_vals"] = nz, gnp[nz]
    # level scales as the device evaluates them (the one libm-dependent constant of the path)
    out["grid_level_scales"] = npy(torch.exp2(torch.arange(16, device=dev, dtype=torch.float32) * S) * 16 - 1)
    # ---- SH / frequency ---------------------------------------------------------------------------------------------------
    she, fre = load("shencoder"), load("freqencoder")
    dd, xx, g_sh, g_fr = gi.sh_freq_inputs()
    y = torch.empty(512, 16, device=dev)
    dyd = torch.empty(512, 48, device=dev)
    she.sh_encode_forward(to(dd), y, 512, 3, 4, dyd)
    gin = torch.zeros(512, 3, device=dev)
    she.sh_encode_backward(to(g_sh), to(dd), 512, 3, 4, dyd, gin)
    out["sh_fwd"], out["sh_bwd"] = npy(y), npy(gin)
    yf = torch.empty(512, 39, device=dev)
    fre.freq_encode_forward(to(xx), 512, 3, 6, 39, yf)
    gf = torch.zeros(512, 3, device=dev)
    fre.freq_encode_backward(to(g_fr), yf, 512, 3, 6, 39, gf)
    out["freq_fwd"], out["freq_bwd"] = npy(yf), npy(gf)
    # ---- FFMLP ------------------------------------------------------------------------------------------------------------
    ff = load("ffmlp")
    c = gi.ffmlp_inputs()
    ff.allocate_splitk(c["nl"] + 1)
    fb = torch.zeros(c["nl"], c["B"], c["dh"], device=dev, dtype=torch.float16)
    yo = torch.zeros(c["B"], c["dout"], device=dev, dtype=torch.float16)
    ff.ffmlp_forward(to(c["x"]).half(), to(c["W"]).half(), c["B"], c["din"], c["dout"], c["dh"], c["nl"], 0, 6, fb, yo)
    bb = torch.zeros_like(fb)
    gx = torch.zeros(c["B"], c["din"], device=dev, dtype=torch.float16)
    gw = torch.zeros(c["W"].shape[0], device=dev, dtype=torch.float16)
    ff.ffmlp_backward(to(c["g"]).half(), to(c["x"]).half(), to(c["W"]).half(), fb, c["B"], c["din"], c["dout"], c["dh"], c["nl"], 0, 6, True, bb, gx, gw)
    torch.cuda.synchronize()
    out["ffmlp_fwd"], out["ffmlp_buffer"], out["ffmlp_gx"], out["ffmlp_gw"] = npy(yo.float()), npy(fb.float()), npy(gx.float()), npy(gw.float())
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, len(out), "arrays", os.path.getsize(out_path) // 1024, "KB; samples", Mtot)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "gpu_ref.npz"))
