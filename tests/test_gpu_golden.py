"""GPU: our kernels (through the C-ABI host mirror) against the COMMITTED outputs of the unmodified reference kernels
(tests/golden/gpu_ref.npz, made by tests/golden/make_gpu_golden.py on a B200) -- a witness that does not depend on
oracle/_ref being present on the box.  Bit-exact for ray tables / sample positions / morton / packbits; float tolerances
stated per assert."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, G)
import gpu_inputs as gi  # noqa: E402

REF = np.load(os.path.join(G, "gpu_ref.npz"))


def dev():
    return torch.device("cuda:0")


def to(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def npy(t):
    return t.detach().cpu().numpy()


def test_marching_bit_exact_with_reference_outputs():
    from seal3d_b200 import raymarching as rm
    g, sc, N = REF, gi.scene(), gi.N_RAYS
    nears, fars = rm.near_far_from_aabb(to(sc["o"]), to(sc["d"]), to(gi.AABB), 0.2)
    assert np.array_equal(npy(nears), g["nears"]) and np.array_equal(npy(fars), g["fars"])
    for tag, nz in (("", np.zeros(N, np.float32)), ("perturb_", sc["noises"])):
        counter = torch.zeros(2, dtype=torch.int32, device=dev())
        x, d, l, r = rm.march_rays_train(to(sc["o"]), to(sc["d"]), 1.0, to(sc["bits"]), 1, 128, nears, fars, counter, -1, True, 128, True,
                                         0.0, 1024, noises=to(nz))
        M = int(g[tag + "march_total"][0])
        assert np.array_equal(npy(counter), g[tag + "march_total"])
        assert np.array_equal(npy(r)[:, 2], g[tag + "march_counts"]) and np.array_equal(npy(r)[:, 0], np.arange(N))
        assert np.array_equal(npy(x)[:M], g[tag + "march_xyzs"]) and np.array_equal(npy(l)[:M], g[tag + "march_deltas"])
    xi, di, li = rm.march_rays(N, 8, torch.arange(N, dtype=torch.int32, device=dev()), nears.clone(), to(sc["o"]), to(sc["d"]), 1.0, to(sc["bits"]),
                               1, 128, nears, fars, 128, False, 0.0, 1024)
    assert np.array_equal(npy(xi), g["infer_xyzs"]) and np.array_equal(npy(li), g["infer_deltas"])
    coords, grid = gi.morton_inputs()
    assert np.array_equal(npy(rm.morton3D(to(coords))), g["morton"]) and np.array_equal(npy(rm.morton3D_invert(to(g["morton"]))), g["morton_invert"])
    assert np.array_equal(npy(rm.packbits(to(grid), 10.0)), g["packbits"])


def test_compositing_matches_reference_outputs():
    from seal3d_b200 import raymarching as rm
    g = REF
    counts = g["march_counts"].astype(np.int32)
    N, M = counts.shape[0], int(counts.sum())
    rays = np.stack([np.arange(N, dtype=np.int32), np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32), counts], 1)
    fv = gi.field_values(M, N)
    for T in (1e-4, 0.0):
        k = "T%g_" % T
        s, c = to(fv["sigmas"]).requires_grad_(True), to(fv["rgbs"]).requires_grad_(True)
        ws, dp, im = rm.composite_rays_train(s, c, to(g["march_deltas"]), to(rays), T)
        torch.autograd.backward([ws, im], [to(fv["g_ws"]), to(fv["g_img"])])
        np.testing.assert_allclose(npy(ws), g[k + "ws"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(dp), g[k + "depth"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(im), g[k + "image"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(c.grad), g[k + "g_rgbs"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(s.grad), g[k + "g_sigmas"], rtol=1e-4, atol=1e-5)
    rng = np.random.default_rng(2)
    Mi = g["infer_deltas"].shape[0]
    si, ci = rng.uniform(0, 60, Mi).astype(np.float32), rng.uniform(0, 1, (Mi, 3)).astype(np.float32)
    alive, rays_t = torch.arange(N, dtype=torch.int32, device=dev()), to(g["nears"]).clone()
    ws, dp, im = torch.zeros(N, device=dev()), torch.zeros(N, device=dev()), torch.zeros(N, 3, device=dev())
    rm.composite_rays(N, 8, alive, rays_t, to(si), to(ci), to(g["infer_deltas"]), ws, dp, im, 1e-2)
    assert np.array_equal(npy(alive), g["infer_alive"])
    np.testing.assert_allclose(npy(rays_t), g["infer_rays_t"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(npy(ws), g["infer_ws"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(dp), g["infer_depth"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(npy(im), g["infer_image"], rtol=1e-5, atol=1e-6)


def test_grid_encoder_matches_reference_outputs():
    from seal3d_b200 import _lib
    g = REF
    offsets, pls, emb, x, gr = gi.grid_inputs()
    B, S = x.shape[0], float(np.log2(pls))
    out = torch.empty(16, B, 2, device=dev())
    dy = torch.empty(B, 96, device=dev())
    _lib.call("s3d_grid_encode_forward", to(x), to(emb), to(offsets), out, B, 3, 2, 16, S, 16, dy, 0, 0, 0, 0)
    assert np.array_equal(npy(out), g["grid_fwd_f32"]), "float32 forward is bit-identical to the reference kernel"
    # dy_dx: the reference initialises pos_deriv = {1.0f} (element 0 only), so only d/dx0 is non-zero with linear interpolation
    np.testing.assert_allclose(npy(dy), g["grid_dy_dx"], rtol=1e-5, atol=1e-4)
    assert not g["grid_dy_dx"].reshape(B, 16, 3, 2)[:, :, 1:].any()
    oh = torch.empty(16, B, 2, device=dev(), dtype=torch.float16)
    _lib.call("s3d_grid_encode_forward", to(x), to(emb).half(), to(offsets), oh, B, 3, 2, 16, S, 16, None, 0, 0, 0, 1)
    np.testing.assert_allclose(npy(oh.float()), g["grid_fwd_f16"], rtol=0, atol=4e-3)      # the reference rounds to half after every corner
    ge = torch.zeros(int(offsets[-1]), 2, device=dev())
    _lib.call("s3d_grid_encode_backward", to(gr), to(x[:128]), to(emb), to(offsets), ge, 128, 3, 2, 16, S, 16, None, None, 0, 0, 0, 0)
    got = npy(ge)
    rows = g["grid_bwd_rows"]
    assert np.array_equal(np.nonzero(np.abs(got).sum(1))[0], rows)
    np.testing.assert_allclose(got[rows], g["grid_bwd_vals"], rtol=1e-5, atol=1e-6)


def test_sh_freq_ffmlp_match_reference_outputs():
    from seal3d_b200 import _lib
    from seal3d_b200.shencoder import sh_encode
    from seal3d_b200.freqencoder import FreqEncoder
    g = REF
    d, x, g_sh, g_fr = gi.sh_freq_inputs()
    dt = to(d).requires_grad_(True)
    y = sh_encode(dt, 4, True)
    y.backward(to(g_sh))
    np.testing.assert_allclose(npy(y), g["sh_fwd"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(npy(dt.grad), g["sh_bwd"], rtol=2e-4, atol=2e-5)
    xt = to(x).requires_grad_(True)
    yf = FreqEncoder(3, 6)(xt)
    yf.backward(to(g_fr))
    np.testing.assert_allclose(npy(yf), g["freq_fwd"], atol=1e-6)
    np.testing.assert_allclose(npy(xt.grad), g["freq_bwd"], rtol=1e-4, atol=1e-4)
    c = gi.ffmlp_inputs()
    fb = torch.zeros(c["nl"], c["B"], c["dh"], device=dev(), dtype=torch.float16)
    out = torch.zeros(c["B"], c["dout"], device=dev(), dtype=torch.float16)
    _lib.call("s3d_ffmlp_forward", to(c["x"]).half(), to(c["W"]).half(), c["B"], c["din"], c["dout"], c["dh"], c["nl"], 0, 6, fb, out)
    # the reference accumulates in fp16: its own rounding (2e-2 of the magnitude) bounds the comparison
    assert np.abs(npy(out.float()) - g["ffmlp_fwd"]).max() <= 2e-2 * max(1.0, np.abs(g["ffmlp_fwd"]).max())
    assert np.abs(npy(fb.float()) - g["ffmlp_buffer"]).max() <= 2e-2 * max(1.0, np.abs(g["ffmlp_buffer"]).max())
    bb = torch.zeros_like(fb)
    gx = torch.zeros(c["B"], c["din"], device=dev(), dtype=torch.float16)
    gw = torch.zeros(c["W"].shape[0], device=dev(), dtype=torch.float16)
    fb_ref = to(g["ffmlp_buffer"]).half()      # the reference's own activations: same ReLU masks on both sides
    _lib.call("s3d_ffmlp_backward", to(c["g"]).half(), to(c["x"]).half(), to(c["W"]).half(), fb_ref, c["B"], c["din"], c["dout"], c["dh"], c["nl"], 0, 6, 1, bb, gx, gw)
    assert np.abs(npy(gx.float()) - g["ffmlp_gx"]).max() <= 3e-2 * np.abs(g["ffmlp_gx"]).max() + 1e-6
    assert np.abs(npy(gw.float()) - g["ffmlp_gw"]).max() <= 3e-2 * np.abs(g["ffmlp_gw"]).max() + 1e-6
