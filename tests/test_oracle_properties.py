"""CPU: size-independent properties of the oracle (oracle/seal_oracle.c) that the domain offers -- round trips, partition of
unity, linearity, finite differences, an independent numpy restatement -- on top of the golden-vector pins of
test_oracle_golden.py.  These are the same invariants the `-m gpu` suite checks on the kernels at BASELINE sizes."""
import numpy as np
import pytest

import oracle

AABB = np.array([-1, -1, -1, 1, 1, 1], np.float32)


@pytest.fixture(scope="module")
def scene():
    from seal3d_b200 import synth
    bits, grid = synth.lego_like_occupancy()
    o, d = synth.rays_for_step(3, 256)
    return dict(bits=bits, grid=grid, o=o, d=d)


def test_morton_roundtrip_and_packbits_against_numpy():
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (5000, 3)).astype(np.int32)
    m = oracle.morton3D(c)
    assert np.array_equal(oracle.morton3D_invert(m), c)
    # bit i of x lands on bit 3i, of y on 3i+1, of z on 3i+2
    for axis, shift in ((0, 0), (1, 1), (2, 2)):
        one = np.zeros((10, 3), np.int32)
        one[:, axis] = 1 << np.arange(10)
        assert np.array_equal(oracle.morton3D(one), (1 << (3 * np.arange(10) + shift)).astype(np.int32))
    grid = rng.uniform(0, 20, 4096).astype(np.float32)
    grid[::7] = 10.0                                   # exactly at the threshold: strict '>' (raymarching.cu:283)
    assert np.array_equal(oracle.packbits(grid, 10.0), np.packbits(grid > 10.0, bitorder="little"))


def test_marching_invariants(scene):
    o, d, bits = scene["o"], scene["d"], scene["bits"]
    nears, fars = oracle.near_far_from_aabb(o, d, AABB, 0.2)
    x, dd, l, r, c = oracle.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars)
    N, M = o.shape[0], int(c[0])
    assert c[1] == N and M == r[:, 2].sum() and np.array_equal(r[:, 0], np.arange(N))
    assert np.array_equal(r[:, 1], np.concatenate([[0], np.cumsum(r[:, 2])[:-1]]))          # ray-major, gap-free offsets
    dt_min = np.float32(2 * np.float32(1.7320508075688772) / 1024)
    assert np.all(l[:M, 0] == dt_min)                                                        # dt_gamma = 0: every step is dt_min
    assert np.all(l[:M, 1] >= dt_min * np.float32(0.999)) and np.all(np.abs(x[:M]) <= 1.0)
    assert not x[M:].any() and not l[M:].any()
    # every sample lies in an occupied cell of the bitfield it was marched on
    cell = np.clip(((x[:M] + 1) * 64).astype(np.int32), 0, 127)
    idx = oracle.morton3D(cell).astype(np.int64)
    assert np.all((bits[idx // 8] >> (idx % 8)) & 1)
    # directions are copied through, samples of a ray advance along it
    k = int(np.argmax(r[:, 2]))
    a, n = r[k, 1], r[k, 2]
    assert n > 3 and np.all(dd[a:a + n] == d[k])
    t = ((x[a:a + n] - o[k]) * d[k]).sum(1)
    assert np.all(np.diff(t) > 0)


def test_compositing_is_linear_in_colour_and_matches_finite_differences(scene):
    o, d, bits = scene["o"][:64], scene["d"][:64], scene["bits"]
    nears, fars = oracle.near_far_from_aabb(o, d, AABB, 0.2)
    _, _, l, r, c = oracle.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars)
    M, N = int(c[0]), 64
    rng = np.random.default_rng(1)
    sig = rng.uniform(0, 30, M).astype(np.float32)
    c1, c2 = rng.uniform(0, 1, (M, 3)).astype(np.float32), rng.uniform(0, 1, (M, 3)).astype(np.float32)
    ws1, dp1, im1 = oracle.composite_rays_train_forward(sig, c1, l[:M], r, 0.0)
    ws2, dp2, im2 = oracle.composite_rays_train_forward(sig, c2, l[:M], r, 0.0)
    ws3, dp3, im3 = oracle.composite_rays_train_forward(sig, (c1 + 2 * c2).astype(np.float32), l[:M], r, 0.0)
    assert np.array_equal(ws1, ws2) and np.array_equal(dp1, dp2)                              # colour does not touch the weights
    np.testing.assert_allclose(im3, im1 + 2 * im2, rtol=1e-5, atol=1e-6)
    assert np.all(ws1 <= 1.0 + 1e-6) and np.all(ws1 >= 0)
    # backward vs central differences of L = <g_ws, ws> + <g_img, img> on one ray's densities
    g_ws, g_im = rng.normal(size=N).astype(np.float32), rng.normal(size=(N, 3)).astype(np.float32)
    gs, gc = oracle.composite_rays_train_backward(g_ws, g_im, sig, c1, l[:M], r, ws1, im1, 0.0)

    def loss(s):
        w, _, im = oracle.composite_rays_train_forward(s.astype(np.float32), c1, l[:M], r, 0.0)
        return float((w.astype(np.float64) * g_ws).sum() + (im.astype(np.float64) * g_im).sum())

    k = int(np.argmax(r[:, 2]))
    for j in (r[k, 1], r[k, 1] + r[k, 2] // 2):
        e = np.zeros(M, np.float64)
        e[j] = 0.05
        fd = (loss(sig + e) - loss(sig - e)) / 0.1
        assert abs(fd - gs[j]) <= 2e-2 * max(1.0, abs(fd)), (j, fd, gs[j])


def test_grid_encoder_partition_of_unity_and_adjoint():
    offsets, pls = oracle.grid_offsets(desired_resolution=2048)
    rng = np.random.default_rng(2)
    x = rng.uniform(0, 1, (512, 3)).astype(np.float32)
    x[0] = [1.5, 0.5, 0.5]
    const = np.tile(np.array([[0.75, -2.0]], np.float32), (int(offsets[-1]), 1))
    out, _ = oracle.grid_encode_forward(x, const, offsets, pls, 16)
    np.testing.assert_allclose(out[:, 1:], np.broadcast_to(const[0], out[:, 1:].shape), rtol=2e-6)    # trilinear weights sum to 1
    assert not out[:, 0].any()                                                                        # out of range -> 0
    # backward is the adjoint of forward: <forward(E), G> == <E, backward(G)>
    emb = rng.normal(size=const.shape).astype(np.float32)
    g = rng.normal(size=(16, 512, 2)).astype(np.float32)
    f, _ = oracle.grid_encode_forward(x, emb, offsets, pls, 16)
    ge = oracle.grid_encode_backward(g, x, emb.shape, offsets, pls, 16)
    lhs, rhs = float((f.astype(np.float64) * g).sum()), float((emb.astype(np.float64) * ge).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


def test_vm_lookup_against_an_independent_numpy_bilinear():
    rng = np.random.default_rng(3)
    R, res = 8, (9, 7, 11)                                           # resolution (x, y, z)
    mat_ids, vec_ids = [(0, 1), (0, 2), (1, 2)], [2, 1, 0]
    mats = [rng.normal(size=(R, res[b], res[a])).astype(np.float32) for a, b in mat_ids]
    vecs = [rng.normal(size=(R, res[v])).astype(np.float32) for v in vec_ids]
    x = rng.uniform(-1, 1, (200, 3)).astype(np.float32)

    def lin(img, axis_len, u):                                       # 1-D linear interpolation along the last axis, align_corners
        p = (u.astype(np.float64) + 1) / 2 * (axis_len - 1)
        i0 = np.clip(np.floor(p).astype(int), 0, axis_len - 2)
        f = p - i0
        return img[..., i0] * (1 - f) + img[..., i0 + 1] * f

    want = np.zeros((200, 3 * R))
    for i, ((a, b), v) in enumerate(zip(mat_ids, vec_ids)):
        rows = lin(mats[i].astype(np.float64), res[a], x[:, a])                  # [R, H, N] -> interpolate W
        py = (x[:, b].astype(np.float64) + 1) / 2 * (res[b] - 1)
        j0 = np.clip(np.floor(py).astype(int), 0, res[b] - 2)
        fy = py - j0
        n = np.arange(200)
        plane = rows[:, j0, n] * (1 - fy) + rows[:, j0 + 1, n] * fy               # [R, N]
        line = lin(vecs[i].astype(np.float64), res[v], x[:, v])                  # [R, N]
        want[:, i * R:(i + 1) * R] = (plane * line).T
    got = oracle.vm_forward(x, mats, vecs, False)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(oracle.vm_forward(x, mats, vecs, True), want.sum(1), rtol=1e-4, atol=1e-5)
    # adjoint: <forward, G> == <mats, g_mats> == <vecs, g_vecs> (the lookup is linear in each factor)
    g = rng.normal(size=(200, 3 * R)).astype(np.float32)
    gm, gv = oracle.vm_backward(x, mats, vecs, False, g)
    lhs = float((got.astype(np.float64) * g).sum())
    assert abs(sum(float((a.astype(np.float64) * b).sum()) for a, b in zip(mats, gm)) - lhs) <= 1e-4 * abs(lhs)
    assert abs(sum(float((a.astype(np.float64) * b).sum()) for a, b in zip(vecs, gv)) - lhs) <= 1e-4 * abs(lhs)


def test_sh_is_orthonormal_on_the_sphere_and_freq_matches_numpy():
    rng = np.random.default_rng(4)
    d = rng.normal(size=(200000, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    y, _ = oracle.sh_encode_forward(d, 4)
    gram = (y.astype(np.float64).T @ y) / d.shape[0] * 4 * np.pi                 # Monte-Carlo integral of Y_i Y_j over the sphere
    np.testing.assert_allclose(gram, np.eye(16), atol=0.03)
    x = rng.uniform(-1, 1, (300, 5)).astype(np.float32)
    f = oracle.freq_encode_forward(x, 3)
    want = np.concatenate([x] + [fn(x.astype(np.float64) * 2 ** k) for k in range(3) for fn in (np.sin, np.cos)], 1)
    np.testing.assert_allclose(f, want, atol=2e-6)
