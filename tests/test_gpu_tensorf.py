"""GPU parity of the TensoRF VM path (BASELINE config 5): csrc/tensorf.cu through the C-ABI and the
``tensorf.TensoRFNetwork`` mirror, against
  (1) the CPU oracle (oracle.vm_forward / vm_backward / TensoRFField) on the same inputs,
  (2) tests/golden/cpu_tensorf.npz = the reference's own tensoRF/network.py::NeRFNetwork run on CPU torch.
Tolerance: 1e-4 relative (fp32), stated per assert; full-size runs use size-independent properties."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cpu_tensorf.npz"))


def dev():
    return torch.device("cuda:0")


def to(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def npy(t):
    return t.detach().cpu().numpy()


def golden_net(cls=None, prefix="", res=None):
    from seal3d_b200.tensorf import TensoRFNetwork
    cls = cls or TensoRFNetwork
    net = cls(resolution=list(res if res is not None else G["resolution"]), bound=1).to(dev())
    sd = {}
    for i in range(3):
        for name in ("sigma_mat", "sigma_vec", "color_mat", "color_vec"):
            sd["%s.%d" % (name, i)] = torch.from_numpy(G["%s%s%d" % (prefix, name, i)])
    sd["basis_mat.weight"] = torch.from_numpy(G["basis_mat"])
    for l in range(3):
        sd["color_net.%d.weight" % l] = torch.from_numpy(G["color_net%d" % l])
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("aabb", "density", "step_counter")) for k in missing), (missing, unexpected)
    return net


def factor_lists(net, which):
    mats = [npy(p)[0] for p in getattr(net, which + "_mat")]            # reference layout [R,H,W]
    vecs = [npy(p)[0, :, :, 0] for p in getattr(net, which + "_vec")]   # [R,D]
    return mats, vecs


def test_state_dict_keeps_reference_shapes_in_channel_last_storage():
    net = golden_net()
    sd = net.state_dict()
    R, (rx, ry, rz) = 16, G["resolution"]
    assert tuple(sd["sigma_mat.0"].shape) == (1, R, ry, rx) and tuple(sd["sigma_vec.0"].shape) == (1, R, rz, 1)
    assert tuple(sd["color_mat.2"].shape) == (1, 48, rz, ry) and tuple(sd["basis_mat.weight"].shape) == (27, 144)
    assert net.sigma_mat[0].stride() == (ry * rx * R, 1, rx * R, R)     # physically [H,W,R]
    np.testing.assert_array_equal(npy(sd["color_vec.1"]), G["color_vec1"])


@pytest.mark.parametrize("tag", ["", "shrunk_"])
def test_vm_lookups_match_oracle_and_reference(tag):
    net = golden_net()
    aabb = G["aabb_shrunk"] if tag else np.array([-1, -1, -1, 1, 1, 1], np.float32)
    net.aabb_train.copy_(to(aabb))
    x = to(G["x"])
    sf = net.get_sigma_feat(x, net.aabb_train)
    sm, sv = factor_lists(net, "sigma")
    np.testing.assert_allclose(npy(sf), oracle.vm_forward(G["x"], sm, sv, True, aabb), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(sf), G[tag + "sigma_feat"], rtol=1e-4, atol=1e-6)
    cm, cv = factor_lists(net, "color")
    prod = net._lookup(x, net.color_mat, net.color_vec, False, net.aabb_train)
    np.testing.assert_allclose(npy(prod), oracle.vm_forward(G["x"], cm, cv, False, aabb), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(npy(net.get_color_feat(x, net.aabb_train)), G[tag + "color_feat"], rtol=1e-4, atol=2e-6)
    # pre-normalised coordinates (the reference's own calling convention for get_*_feat) give the same answer
    xn = 2 * (x - net.aabb_train[:3]) / (net.aabb_train[3:] - net.aabb_train[:3]) - 1
    assert torch.equal(net.get_sigma_feat(xn), sf)


@pytest.mark.parametrize("tag", ["", "shrunk_"])
def test_field_forward_and_autograd_match_reference(tag):
    net = golden_net()
    if tag:
        net.aabb_train.copy_(to(G["aabb_shrunk"]))
    x, d = to(G["x"]), to(G["d"])
    sigma, rgb = net(x, d)
    np.testing.assert_allclose(npy(sigma), G[tag + "sigma"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(npy(rgb), G[tag + "rgb"], rtol=1e-4, atol=5e-6)
    np.testing.assert_allclose(npy(net.density(x)["sigma"]), G[tag + "sigma"], rtol=1e-4, atol=1e-6)
    msk = torch.zeros(x.shape[0], dtype=torch.bool, device=dev())
    msk[::3] = True
    np.testing.assert_allclose(npy(net.color(x, d, mask=msk)), G[tag + "color_masked"], rtol=1e-4, atol=5e-6)
    ((sigma * to(G[tag + "g_sigma"])).sum() + (rgb * to(G[tag + "g_rgb"])).sum()).backward()
    for i in range(3):
        for name in ("sigma_mat", "sigma_vec", "color_mat", "color_vec"):
            ref = G["%sgrad_%s%d" % (tag, name, i)]
            p = getattr(net, name)[i]
            assert p.grad.stride() == p.stride()
            np.testing.assert_allclose(npy(p.grad), ref, rtol=1e-4, atol=2e-5 * max(1.0, np.abs(ref).max()), err_msg="%s%d" % (name, i))
    np.testing.assert_allclose(npy(net.basis_mat.weight.grad), G[tag + "grad_basis_mat"], rtol=1e-4, atol=2e-5)
    for l in range(3):
        ref = G["%sgrad_color_net%d" % (tag, l)]
        np.testing.assert_allclose(npy(net.color_net[l].weight.grad), ref, rtol=1e-4, atol=2e-5 * max(1.0, np.abs(ref).max()))


def test_vm_backward_kernel_matches_oracle_double_accumulation():
    from seal3d_b200 import _lib
    net = golden_net()
    rng = np.random.default_rng(5)
    M = 4096
    x = rng.uniform(-1.02, 1.02, (M, 3)).astype(np.float32)
    for which, reduce in (("sigma", True), ("color", False)):
        mats, vecs = getattr(net, which + "_mat"), getattr(net, which + "_vec")
        R = mats[0].shape[1]
        g = rng.normal(size=(M,) if reduce else (M, 3 * R)).astype(np.float32)
        dims = _lib.host_i32([[m.shape[2], m.shape[3], v.shape[2]] for m, v in zip(mats, vecs)])
        grads = [torch.zeros_like(p) for p in list(mats) + list(vecs)]
        _lib.call("s3d_vm_backward", to(x), M, None, *mats, *vecs, dims[1], R, int(reduce), to(g), *grads)
        om, ov = factor_lists(net, which)
        gm, gv = oracle.vm_backward(x, om, ov, reduce, g)
        for i in range(3):
            np.testing.assert_allclose(npy(grads[i])[0], gm[i], rtol=1e-4, atol=1e-5 * np.abs(gm[i]).max())
            np.testing.assert_allclose(npy(grads[3 + i])[0, :, :, 0], gv[i], rtol=1e-4, atol=1e-5 * np.abs(gv[i]).max())


def test_vm_lookup_fp16_io_is_the_fp32_lookup_rounded_once():
    """s3d_vm_forward_f16 / _backward_f16 (the colour features of an fp16 step): the forward is bit for bit the fp32 lookup cast
    to fp16 (what the reference's autocast basis_mat does to it, tensoRF/network.py:155); the backward with an fp16 output
    gradient equals the fp32 kernel fed the same values"""
    from seal3d_b200 import _lib
    net = golden_net()
    rng = np.random.default_rng(11)
    M = 20000
    x = to(rng.uniform(-1.02, 1.02, (M, 3)).astype(np.float32))
    mats, vecs = net.color_mat, net.color_vec
    R = mats[0].shape[1]
    dims = _lib.host_i32([[m.shape[2], m.shape[3], v.shape[2]] for m, v in zip(mats, vecs)])
    out32 = torch.empty(M, 3 * R, device=dev())
    out16 = torch.empty(M, 3 * R, device=dev(), dtype=torch.float16)
    _lib.call("s3d_vm_forward", x, M, None, *mats, *vecs, dims[1], R, 0, out32)
    _lib.call("s3d_vm_forward_f16", x, M, None, *mats, *vecs, dims[1], R, out16)
    assert torch.equal(out16, out32.half())
    g16 = to(rng.normal(size=(M, 3 * R)).astype(np.float32)).half()
    ga = [torch.zeros_like(p) for p in list(mats) + list(vecs)]
    gb = [torch.zeros_like(p) for p in list(mats) + list(vecs)]
    _lib.call("s3d_vm_backward", x, M, None, *mats, *vecs, dims[1], R, 0, g16.float(), *ga)
    _lib.call("s3d_vm_backward_f16", x, M, None, *mats, *vecs, dims[1], R, g16, *gb)
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max())      # same products, atomics in a different order
    # through the module: "half" output + fp16 upstream gradient
    out = net._lookup(x, mats, vecs, "half")
    assert out.dtype == torch.float16 and torch.equal(out, out16)
    net.zero_grad(set_to_none=True)
    out.backward(g16)
    for p, a in zip(list(mats) + list(vecs), ga):
        assert float((p.grad - a).abs().max()) <= 1e-5 * float(a.abs().max())


def test_out_of_box_and_argument_errors():
    from seal3d_b200 import _lib
    net = golden_net()
    far = to(np.array([[1.5, 0.0, 0.0], [0.0, -1.7, 0.2], [3e9, -3e9, 7.0], [np.nan, 0.0, 0.0]], np.float32))
    assert not net.get_sigma_feat(far).any()                      # zero padding; huge / NaN coordinates do not fault
    assert net._lookup(far[2:3], net.color_mat, net.color_vec, False).abs().max() == 0
    assert net.get_sigma_feat(far[:0]).shape == (0,)
    with pytest.raises(_lib.S3DError):                            # factor images in the reference's contiguous layout are refused
        net._lookup(far, [p.data.contiguous() for p in net.sigma_mat], list(net.sigma_vec), True)
    dims = _lib.host_i32([[4, 4, 4]] * 3)
    out = torch.empty(4, device=dev())
    with pytest.raises(_lib.S3DError):                            # R must be a multiple of 4
        _lib.call("s3d_vm_forward", far, 4, None, *([out] * 6), dims[1], 6, 1, out)
    with pytest.raises(_lib.S3DError):                            # reduce needs R/4 a power of two (48 channels is the product form)
        _lib.call("s3d_vm_forward", far, 4, None, *([out] * 6), dims[1], 48, 1, out)


def test_upsample_and_shrink_match_reference():
    from seal3d_b200 import raymarching
    net = golden_net()
    net.upsample_model([int(v) for v in G["up_resolution"]])
    for i in range(3):
        for name in ("sigma_mat", "sigma_vec", "color_mat", "color_vec"):
            np.testing.assert_allclose(npy(getattr(net, name)[i]), G["up_%s%d" % (name, i)], rtol=1e-5, atol=1e-6)
    assert net.resolution == [int(v) for v in G["up_resolution"]]
    sf = net.get_sigma_feat(to(G["x"]))                           # the re-allocated factors are usable by the kernels
    assert torch.isfinite(sf).all()
    # shrink: occupied box of density-grid cells -> cropped factors and aabb
    net2 = golden_net(prefix="pre_shrink_")
    (x0, x1), (y0, y1), (z0, z1) = G["shrink_density_grid_cells"]
    cells = torch.stack(torch.meshgrid(torch.arange(x0, x1), torch.arange(y0, y1), torch.arange(z0, z1), indexing="ij"), -1).reshape(-1, 3)
    idx = raymarching.morton3D(cells.int().to(dev())).long()
    net2.density_grid.zero_()
    net2.density_grid[0, idx] = 50.0
    net2.density_thresh, net2.mean_density = 10.0, 20.0
    net2.shrink_model()
    np.testing.assert_allclose(npy(net2.aabb_train), G["shrink_aabb"], rtol=1e-6, atol=1e-7)
    for i in range(3):
        for name in ("sigma_mat", "sigma_vec", "color_mat", "color_vec"):
            np.testing.assert_array_equal(npy(getattr(net2, name)[i]), G["shrink_%s%d" % (name, i)])


def test_full_size_properties():
    """resolution 300 (main_SealTensoRF.py --resolution1), 2^20 samples: the lookup is linear in every factor image, so
    out(2*mat) = 2*out exactly in fp32, and by Euler's theorem <grad_mat, mat> summed over planes = <g, out> = the same for lines"""
    from seal3d_b200 import _lib
    from seal3d_b200.tensorf import TensoRFNetwork
    torch.manual_seed(3)
    net = TensoRFNetwork(resolution=[300, 300, 300], bound=1).to(dev())
    M = 1 << 20
    x = torch.rand(M, 3, device=dev()) * 2 - 1
    for which, reduce in (("sigma", True), ("color", False)):
        mats, vecs = list(getattr(net, which + "_mat")), list(getattr(net, which + "_vec"))
        out = net._lookup(x, mats, vecs, reduce)
        with torch.no_grad():
            out2 = net._lookup(x, [m * 2 for m in mats], vecs, reduce)
        assert torch.equal(out2, out * 2)
        g = torch.randn_like(out)
        out.backward(g)
        want = float((g.double() * out.detach().double()).sum())
        got_m = sum(float((p.grad.double() * p.detach().double()).sum()) for p in mats)
        got_v = sum(float((p.grad.double() * p.detach().double()).sum()) for p in vecs)
        scale = float((g.double() * out.detach().double()).abs().sum())
        assert abs(got_m - want) < 1e-5 * scale and abs(got_v - want) < 1e-5 * scale, (which, want, got_m, got_v, scale)


def test_colour_head_on_tensor_cores_matches_the_linear_path():
    """under fp16 autocast basis_mat and the 150-128-128-3 colour MLP run on the tcgen05 kernels (s3d_linear_*, wide FFMLP:
    _linear_tc / _mlp_head_tc) instead of autocast F.linear: same rgb and the same gradients for every parameter, to the
    fp16 rounding both paths share (the reference's own fp16 step is the F.linear path, tensoRF/network.py:148-178)"""
    from seal3d_b200 import tensorf as tf
    torch.manual_seed(1)
    net = golden_net(res=[40, 36, 44])
    x = (torch.rand(5000, 3, device=dev()) * 1.9 - 0.95)
    d = torch.nn.functional.normalize(torch.randn(5000, 3, device=dev()), dim=-1)
    g = torch.randn(5000, 3, device=dev()) * 64.0          # a loss-scaled upstream gradient
    outs = []
    for tc in (True, False):
        net.zero_grad(set_to_none=True)
        orig = tf.TensoRFNetwork.num_layers if hasattr(tf.TensoRFNetwork, "num_layers") else None
        if not tc:
            net.hidden_dim_saved, net.hidden_dim = net.hidden_dim, -1          # disables the tensor-core branch of _color_mlp
        try:
            with torch.autocast("cuda", dtype=torch.float16):
                sigma, rgb = net(x, d)
        finally:
            if not tc:
                net.hidden_dim = net.hidden_dim_saved
        rgb.float().backward(g)
        grads = {n: p.grad.detach().clone().float() for n, p in net.named_parameters() if p.grad is not None}
        outs.append((rgb.float().detach(), grads))
    (rgb_tc, g_tc), (rgb_ref, g_ref) = outs
    np.testing.assert_allclose(npy(rgb_tc), npy(rgb_ref), rtol=0, atol=3e-3)
    assert set(g_tc) == set(g_ref) and any(k.startswith("color_net") for k in g_tc) and "basis_mat.weight" in g_tc
    for k in g_ref:
        a, b = npy(g_tc[k]), npy(g_ref[k])
        assert np.abs(a - b).max() <= 2e-2 * np.abs(b).max() + 1e-6, (k, np.abs(a - b).max(), np.abs(b).max())


def test_distillation_with_tensorf_teacher_and_student():
    """config 5 in miniature: TensoRF teacher (proxy-mapped, colour edit) -> TensoRF student on shared samples through the same
    trainer as the NGP backbone; the loss falls, the arena keeps the channels_last layout, two learning rates are applied"""
    from seal3d_b200 import synth
    from seal3d_b200.seal import TensoRFTeacherNetwork, TensoRFStudentNetwork, SealBBoxMapper
    from seal3d_b200.trainer import DistillTrainer
    torch.manual_seed(0)
    teacher = TensoRFTeacherNetwork(resolution=[64] * 3, bound=1).to(dev())
    student = TensoRFStudentNetwork(resolution=[64] * 3, bound=1).to(dev())
    with torch.no_grad():
        for p in teacher.sigma_vec:
            p.add_(0.5)                     # a non-degenerate teacher density: sigma_feat ~ 3 * 16 * 0.1 * 0.5
        for p in teacher.sigma_mat:
            p.add_(0.1)
    bits, _ = synth.lego_like_occupancy()
    md, tris = synth.bbox_edit()
    md = dict(md)
    md["hsv"] = np.array([0.3, 0.0, 0.0], np.float32)
    mapper = SealBBoxMapper(md, tris, device=dev())
    for net in (teacher, student):
        net.density_bitfield.copy_(to(bits))
        net.init_mapper(mapper)
        net.hack_bitfield()
    tr = DistillTrainer(student, teacher, lr=(2e-2, 1e-3), update_interval=0)
    assert student.color_mat[0].stride() == (64 * 64 * 48, 1, 64 * 48, 48)
    o, d = synth.rays_for_step(0, 2048)
    before = [p.detach().clone() for p in (student.sigma_mat[0], student.color_net[0].weight)]
    losses = []
    for it in range(20):
        losses.append(npy(tr.distill_step(to(o), to(d), perturb=False, force_all_rays=True)).copy())
    losses = np.array(losses)
    assert np.isfinite(losses).all() and losses[-1].sum() < 0.8 * losses[0].sum(), losses
    moved = [float((p.detach() - b).abs().max()) for p, b in zip((student.sigma_mat[0], student.color_net[0].weight), before)]
    assert 0 < moved[1] <= 20 * 1e-3 * 2 and moved[0] > moved[1], moved       # Adam moves ~lr per step: lr2 on the MLP, lr1 on the factors
    # pretraining stage on cached teacher values: nothing is frozen for TensoRF (SealNeRF/trainer.py:476-483)
    pts = to(np.random.default_rng(1).uniform(-0.3, 0.5, (8192, 3)).astype(np.float32))
    dirs = torch.nn.functional.normalize(torch.randn(8192, 3, device=dev()), dim=-1)
    with torch.no_grad():
        mx, mdirs, mask = teacher._map_samples(pts, dirs)
        sig_t, rgb_t = teacher(mx, mdirs)
        rgb_t = teacher._map_colors(mx, mdirs, rgb_t.float().contiguous(), mask)
    w0 = student.color_net[1].weight.detach().clone()
    l0 = float(tr.pretrain_step(pts, dirs, sig_t.contiguous(), rgb_t.contiguous())[0])
    for _ in range(10):
        l1 = float(tr.pretrain_step(pts, dirs, sig_t.contiguous(), rgb_t.contiguous())[0])
    assert l1 < l0 and not torch.equal(w0, student.color_net[1].weight.detach())
