"""CPU: host-side logic -- the C-ABI library loads and exports every symbol the header declares, the five
reference-named extension modules import and expose the reference's 22 functions, the synthetic workload is
deterministic, and the data-parallel sharding + single all-reduce reproduces the single-process gradient
(world_size 2, gloo)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "seal-3d_b200")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "seal3d_b200.h")).read()
    return sorted(set(re.findall(r"^int (s3d_\w+)\(", src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(PKG, "libseal3d_b200.so")):
        g.build()
    lib = ctypes.CDLL(os.path.join(PKG, "libseal3d_b200.so"))
    syms = _header_symbols()
    assert len(syms) >= 34
    for s in syms:
        assert hasattr(lib, s), "header declares %s but the library does not export it" % s
    from seal3d_b200 import _lib
    assert set(_lib.exported_symbols()) == set(syms), set(_lib.exported_symbols()) ^ set(syms)


REF_SURFACE = {
    "_raymarching": ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
                     "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "composite_rays"],
    "_gridencoder": ["grid_encode_forward", "grid_encode_backward", "grad_total_variation"],
    "_shencoder": ["sh_encode_forward", "sh_encode_backward"],
    "_freqencoder": ["freq_encode_forward", "freq_encode_backward"],
    "_ffmlp": ["ffmlp_forward", "ffmlp_inference", "ffmlp_backward", "allocate_splitk", "free_splitk"],
}


def test_reference_named_modules_export_the_22_functions():
    """the drop-in boundary is the set of compiled modules the reference wrappers import first"""
    import __graft_entry__ as g
    if not all(os.path.exists(os.path.join(PKG, m + ".so")) for m in REF_SURFACE):
        g.build()
    sys.path.insert(0, PKG)
    try:
        n = 0
        for mod, fns in REF_SURFACE.items():
            m = __import__(mod)
            for f in fns:
                assert callable(getattr(m, f)), (mod, f)
                n += 1
        assert n == 22
        import _gridencoder
        with pytest.raises(RuntimeError):   # CPU tensors are rejected like the reference's CHECK_CUDA
            _gridencoder.grid_encode_forward(torch.zeros(4, 3), torch.zeros(8, 2), torch.zeros(2, dtype=torch.int32), torch.zeros(1, 4, 2),
                                             4, 3, 2, 1, 1.0, 16, None, 0, False, 0)
    finally:
        sys.path.remove(PKG)


def test_missing_library_fails_loudly(monkeypatch):
    from seal3d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libseal3d_b200.so")
    with pytest.raises(_lib.S3DError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_synthetic_workload_is_deterministic():
    from seal3d_b200 import synth
    o1, d1 = synth.rays_for_step(3, 1000)
    o2, d2 = synth.rays_for_step(3, 1000)
    assert np.array_equal(o1, o2) and np.array_equal(d1, d2)
    np.testing.assert_allclose(np.linalg.norm(d1, axis=1), 1, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(o1, axis=1), synth.CAMERA_RADIUS, rtol=1e-5)
    bits, grid = synth.lego_like_occupancy()
    assert bits.shape == (128 ** 3 // 8,) and 0.01 < np.unpackbits(bits).mean() < 0.1
    offs, pls = synth.grid_offsets()
    assert offs[-1] == 6119864 and abs(pls - 2 ** (7 / 15)) < 1e-12
    md, tris = synth.bbox_edit()
    assert tris.shape == (12, 3, 3) and np.allclose(md["map_bound"], [[0.15, -0.15, -0.15], [0.45, 0.15, 0.15]])


def test_shard_bounds_cover_everything():
    from seal3d_b200.parallel import shard_bounds
    for n in (0, 1, 7, 4096, 262145):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from seal3d_b200 import parallel
    r, _, w = parallel.init_from_env("gloo")
    rng = np.random.default_rng(0)
    offsets, pls = oracle.grid_offsets(num_levels=4, log2_hashmap_size=10, desired_resolution=64)
    n = int(offsets[-1])
    x = rng.uniform(0, 1, (600, 3)).astype(np.float32)
    g = rng.normal(size=(4, 600, 2)).astype(np.float32)
    lo, hi = parallel.shard_bounds(600, r, w)
    part = oracle.grid_encode_backward(np.ascontiguousarray(g[:, lo:hi]), x[lo:hi], (n, 2), offsets, pls, 16)
    flat = torch.from_numpy(part.reshape(-1).copy())
    parallel.allreduce_sum_(flat)
    t = parallel.max_over_ranks(1.0 + r, torch.device("cpu"))
    if r == 0:
        full = oracle.grid_encode_backward(g, x, (n, 2), offsets, pls, 16)
        q.put((float(np.abs(flat.numpy() - full.reshape(-1)).max()), float(np.abs(full).max()), t))
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_single_process_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    err, scale, t = q.get(timeout=120)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert err <= 1e-5 * scale and t == 2.0


class _StubStudent:
    def __init__(self):
        import torch
        self.mean_count = 0
        self.step_counter = torch.zeros(16, 2, dtype=torch.int32)


class _StubTrainer:
    """a trainer whose every step contains one all-reduce, like FusedDistillTrainer._reduce_and_step"""

    def __init__(self):
        import torch
        self.student, self.calls, self.grad = _StubStudent(), 0, torch.ones(64)

    def distill_step(self, o, d, perturb=True, force_all_rays=False, prefetch=None):
        import torch
        import torch.distributed as dist
        self.calls += 1
        g = self.grad.clone()
        dist.all_reduce(g)
        self.student.step_counter[self.calls % 16, 0] = 1000 + self.calls
        return torch.tensor([float(g[0]), float(self.calls)])

    def refresh_occupancy(self):
        self.student.mean_count = 1


def _bench_flow_worker(rank, world, port, q, rank0_extra_step):
    import datetime
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import argparse
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=8))
    import bench
    tr = _StubTrainer()
    batches = [(torch.zeros(4, 3), torch.zeros(4, 3)) for _ in range(4)]
    args = argparse.Namespace(rays=4, steps=3, warmup=3, no_roofline=False)
    rows = []
    try:
        r = bench.timed_legs(tr, batches, batches, args, rank, world, torch.device("cpu"), pipelined=True,
                             profile_hook=(lambda: rows.clear(), lambda: [("s3d_stub", 1.0)] * 3))
        if rank0_extra_step and rank == 0:
            tr.distill_step(*batches[0])     # what round 1's bench did: a step (= a collective) on rank 0 alone
        dist.barrier()
        q.put((rank, "ok", tr.calls, r["ms"] > 0 and r["ms_e2e"] > 0, r["breakdown"], float(r["last"][0])))
    except Exception as e:       # gloo reports the unmatched collective as a timeout
        q.put((rank, "error", type(e).__name__, None, None, None))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def _run_bench_flow(extra):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + (17 if extra else 0)) % 90
    ps = [ctx.Process(target=_bench_flow_worker, args=(r, 2, port, q, extra)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    return out


def test_bench_control_flow_issues_matched_collectives_gloo():
    """bench.timed_legs (warm-up, both timed legs, the per-kernel breakdown) on 2 gloo ranks with a trainer whose step
    all-reduces: both ranks must take the same number of steps and finish; round 1's rank-0-only breakdown steps hung here"""
    out = _run_bench_flow(False)
    assert [o[1] for o in out] == ["ok", "ok"], out
    assert out[0][2] == out[1][2] == 3 + 6 + 4 + 3 + 3 + 3, out    # warm-up, pipeline priming, collective warm-up, leg 1, leg 2, breakdown
    assert all(o[3] for o in out) and out[0][5] == 2.0              # the all-reduce summed both ranks
    assert out[0][4]["s3d_stub"]["calls_per_step"] == 1.0


def test_bench_control_flow_test_detects_an_unmatched_collective_gloo():
    """the negative control: one extra step on rank 0 alone must NOT pass silently (gloo times out)"""
    out = _run_bench_flow(True)
    assert any(o[1] == "error" for o in out), out


def test_checkpoint_model_roundtrip_in_reference_format(tmp_path):
    """nerf/utils.py:1015-1136: dict layout, keys of SURVEY appendix B, latest-checkpoint lookup, bare state dicts;
    TensoRF factors are written contiguous in the reference shape and land in channels_last storage, at the file's resolution"""
    import torch
    from seal3d_b200 import checkpoint as ck
    from seal3d_b200.seal import StudentNetwork, TensoRFStudentNetwork
    m = StudentNetwork(bound=1)
    m.mean_count, m.mean_density = 4321, 0.25
    ck.save_checkpoint(str(tmp_path / "ngp_ep0002.pth"), model=m, epoch=2)
    ck.save_checkpoint(str(tmp_path / "ngp_ep0010.pth"), model=m, epoch=10)
    assert ck.latest_checkpoint(str(tmp_path), "ngp").endswith("ngp_ep0010.pth")
    raw = torch.load(str(tmp_path / "ngp_ep0010.pth"), weights_only=False)
    assert {"epoch", "global_step", "stats", "mean_count", "mean_density", "model"} <= set(raw) and "optimizer" not in raw
    assert tuple(raw["model"]["encoder.embeddings"].shape) == (6119864, 2) and tuple(raw["model"]["color_net.0.weight"].shape) == (64, 63)
    assert raw["model"]["encoder.offsets"].tolist()[:3] == [0, 4920, 18744] and tuple(raw["model"]["density_bitfield"].shape) == (262144,)
    m2 = StudentNetwork(bound=1)
    missing, unexpected, _ = ck.load_checkpoint(str(tmp_path / "ngp_ep0010.pth"), model=m2)
    assert not missing and not unexpected and m2.mean_count == 4321 and m2.mean_density == 0.25
    assert torch.equal(m2.encoder_color.embeddings, m.encoder_color.embeddings)
    torch.save(raw["model"], str(tmp_path / "bare.pth"))                    # a bare state dict loads too (:1082-1085)
    m3 = StudentNetwork(bound=1)
    ck.load_checkpoint(str(tmp_path / "bare.pth"), model=m3)
    assert torch.equal(m3.sigma_net[1].weight, m.sigma_net[1].weight)
    t = TensoRFStudentNetwork(resolution=[12, 14, 16])
    ck.save_checkpoint(str(tmp_path / "t.pth"), model=t)
    saved = torch.load(str(tmp_path / "t.pth"), weights_only=False)["model"]
    assert tuple(saved["color_mat.0"].shape) == (1, 48, 14, 12) and saved["color_mat.0"].is_contiguous()
    t2 = TensoRFStudentNetwork(resolution=[8, 8, 8])                        # e.g. a shrunk / upsampled checkpoint
    ck.load_checkpoint(str(tmp_path / "t.pth"), model=t2)
    assert t2.resolution == [12, 14, 16] and t2.color_mat[0].stride() == (14 * 12 * 48, 1, 12 * 48, 48)
    assert torch.equal(t2.color_mat[0], t.color_mat[0]) and torch.equal(t2.sigma_vec[2], t.sigma_vec[2])


def test_pretraining_lattice_and_euler_directions_match_scipy():
    """schedule.sample_points vs the reference's recipe (SealNeRF/trainer.py:609-635): torch.arange lattice per box and
    scipy's Rotation.from_euler('xyz', grid, degrees=True).apply([1 - 1e-5, 0, 0]) -- scipy is what the reference calls"""
    import numpy as np
    import torch
    from scipy.spatial.transform import Rotation
    from seal3d_b200.schedule import sample_points
    bounds = np.array([[[0.0, -0.1, 0.2], [0.031, -0.05, 0.26]], [[0.5, 0.5, 0.5], [0.52, 0.53, 0.51]]], np.float32)
    pts, dirs = sample_points(bounds, 0.01, 90)
    want = []
    for lo, hi in bounds:
        X, Y, Z = torch.meshgrid(*[torch.arange(float(lo[d]), float(hi[d]), step=0.01) for d in range(3)], indexing="ij")
        want.append(torch.stack([X, Y, Z], -1).reshape(-1, 3))
    assert torch.equal(pts, torch.cat(want).float()) and pts.shape[0] > 100
    a = np.arange(0, 360, 90)
    e = np.stack(np.meshgrid(a, a, a, indexing="ij"), -1).reshape(-1, 3)
    ref = Rotation.from_euler("xyz", e, degrees=True).apply(np.array([1 - 1e-5, 0, 0]))
    assert dirs.shape == (2 * 64, 3)
    np.testing.assert_allclose(dirs[:64].numpy(), ref, atol=1e-7)
    np.testing.assert_allclose(dirs[64:].numpy(), ref, atol=1e-7)


def test_gradient_arena_chunks_tile_the_levels():
    """fused.FusedNGP.grad_chunks: whole 4-level blocks, contiguous arena slices, the MLP gradients ride on the last slice"""
    import torch
    from seal3d_b200 import synth
    from seal3d_b200.fused import FusedNGP
    off, _ = synth.grid_offsets()
    stub = type("S", (), {})()
    stub.offsets, stub.L, stub.grad = torch.from_numpy(off), 16, torch.zeros(int(off[-1]) * 4 + 11392)
    for n in (1, 2, 3, 4, 8):
        ch = FusedNGP.grad_chunks(stub, n)
        assert 1 <= len(ch) <= min(n, 4) and ch[0][0] == 0 and ch[0][2] == 0 and ch[-1][1] == 16 and ch[-1][3] == stub.grad.numel()
        for a, b in zip(ch, ch[1:]):
            assert a[1] == b[0] and a[3] == b[2] and a[1] % 4 == 0 and a[3] == int(off[a[1]]) * 4
    two = FusedNGP.grad_chunks(stub, 2)
    assert abs((two[0][3] - two[0][2]) - (two[1][3] - two[1][2])) < 0.4 * stub.grad.numel()      # about equal bytes


def test_c_abi_argument_errors_are_reported_before_any_launch():
    """the C-ABI validates shapes before touching the device: < 0 return codes (S3D_EINVAL -22 / S3D_ENOTSUP -95) for the
    configurations the reference rejects with std::runtime_error / TORCH_CHECK; zero-sized batches return 0.  No kernel is
    launched by any of these calls, so they run without a GPU."""
    import ctypes as C
    from seal3d_b200 import _lib
    lib = _lib.lib()
    EINVAL, ENOTSUP, n = -22, -95, None
    dims = (C.c_int * 9)(*([4] * 9))
    # grid encoder: unknown dtype, unsupported (D, C), empty batch
    assert lib.s3d_grid_encode_forward(n, n, n, n, 16, 3, 2, 16, 0.5, 16, n, 0, 0, 0, 7, n) == EINVAL
    assert lib.s3d_grid_encode_forward(n, n, n, n, 16, 6, 2, 16, 0.5, 16, n, 0, 0, 0, 0, n) == EINVAL
    assert lib.s3d_grid_encode_forward(n, n, n, n, 0, 3, 2, 16, 0.5, 16, n, 0, 0, 0, 0, n) == 0
    assert lib.s3d_grad_total_variation(n, n, n, n, 1.0, 16, 3, 2, 16, 0.5, 16, 0, 0, 1, n) == ENOTSUP      # fp16 TV is not a reference path
    # FFMLP: hidden width in {16, 32, 64, 128, 256} (ffmlp.cu:653-658 throws otherwise), input width a multiple of 16, output <= 256
    assert lib.s3d_ffmlp_forward(n, n, 128, 32, 16, 100, 2, 0, 6, n, n, n) == ENOTSUP
    assert lib.s3d_ffmlp_forward(n, n, 128, 32, 16, 512, 2, 0, 6, n, n, n) == ENOTSUP
    assert lib.s3d_ffmlp_forward(n, n, 128, 32, 300, 128, 2, 0, 6, n, n, n) == EINVAL
    assert lib.s3d_ffmlp_forward(n, n, 128, 30, 16, 64, 2, 0, 6, n, n, n) == EINVAL
    assert lib.s3d_ffmlp_forward(n, n, 0, 32, 16, 256, 3, 0, 6, n, n, n) == 0
    assert lib.s3d_ffmlp_backward(n, n, n, n, 128, 32, 16, 64, 2, 2, 6, 1, n, n, n, n) == ENOTSUP        # sine activation has no backward
    assert lib.s3d_ffmlp_backward(n, n, n, n, 128, 32, 16, 128, 2, 2, 6, 1, n, n, n, n) == ENOTSUP
    # VM lookups: rank multiple of 4; the reduced form needs a power-of-two group
    assert lib.s3d_vm_forward(n, 8, n, n, n, n, n, n, n, dims, 6, 1, n, n) == EINVAL
    assert lib.s3d_vm_forward(n, 8, n, n, n, n, n, n, n, dims, 48, 1, n, n) == ENOTSUP
    assert lib.s3d_vm_forward(n, 0, n, n, n, n, n, n, n, dims, 16, 1, n, n) == 0
    assert lib.s3d_vm_resize(n, 0, 4, n, 8, 8, 16, n) == EINVAL
    # fused field: level ranges are whole 4-level groups, at most 16 levels
    assert lib.s3d_ngp_scatter_levels(n, n, 8, 1.0, n, n, 16, 0.5, 16, 1.0, 2, 8, n) == EINVAL
    assert lib.s3d_ngp_scatter_levels(n, n, 8, 1.0, n, n, 20, 0.5, 16, 1.0, 0, 8, n) == ENOTSUP
    assert lib.s3d_ngp_scatter_levels(n, n, 8, 1.0, n, n, 16, 0.5, 16, 1.0, 8, 8, n) == 0
    assert lib.s3d_ngp_encode(n, 8, 1.0, n, 12, n, 16, 0.5, 16, n, 0, n) == EINVAL                          # entry stride is 8 or 16 bytes
    # density grid / rays
    assert lib.s3d_mark_untrained_grid(n, n, 3, 0.5, 0.5, 0, 128, 1.0, n, n) == EINVAL
    assert lib.s3d_mark_untrained_grid(n, n, 5000, 0.5, 0.5, 1, 128, 1.0, n, n) == ENOTSUP
    assert lib.s3d_get_rays(n, 2, 1.0, 1.0, 0.5, 0.5, 4, 4, n, 1, 7, n, n, n) == EINVAL                     # all-pixels form needs N = H*W
    assert lib.s3d_march_rays_train(n, n, n, 1.0, 0.0, 0, 8, 1, 128, 8, n, n, n, n, n, n, n, n, n) == EINVAL  # max_steps = 0
    assert lib.s3d_march_rays_train(n, n, n, 1.0, 0.0, 1024, 0, 1, 128, 8, n, n, n, n, n, n, n, n, n) == 0


def test_schedule_psnr_matches_the_reference_meter():
    """nerf/utils.py:207-235: psnr = -10 log10(mean((pred - truth)^2)) on [0,1] images"""
    import numpy as np
    import torch
    from seal3d_b200.schedule import SealStudentSchedule
    rng = np.random.default_rng(0)
    a, b = rng.uniform(0, 1, (50, 40, 3)).astype(np.float32), rng.uniform(0, 1, (50, 40, 3)).astype(np.float32)
    want = -10 * np.log10(np.mean((a - b) ** 2))
    assert abs(SealStudentSchedule.psnr(torch.from_numpy(a), torch.from_numpy(b)) - want) < 1e-4
    assert abs(SealStudentSchedule.psnr(torch.from_numpy(a), torch.from_numpy(a + 0.1)) - 20.0) < 1e-3


def test_patch_indices_form_contiguous_pixel_blocks():
    """nerf/utils.py:73-92 (patch sampling of get_rays): N // p^2 patches of p x p pixels, inside the image"""
    import torch
    from seal3d_b200.utils import patch_indices
    gen = torch.Generator().manual_seed(0)
    H, W, p = 60, 90, 4
    inds = patch_indices(H, W, 100, p, generator=gen)
    assert inds.shape == (6 * 16,) and inds.dtype == torch.int64
    blocks = inds.view(6, p, p)
    rows, cols = blocks // W, blocks % W
    assert (rows[:, 1:, :] - rows[:, :-1, :] == 1).all() and (cols[:, :, 1:] - cols[:, :, :-1] == 1).all()
    assert rows.min() >= 0 and rows.max() < H and cols.min() >= 0 and cols.max() < W


def test_clock_sampler_helper_process_protocol(tmp_path):
    """bench.ClockSamplerProcess's child (NVML polled from a helper process so that the training process's launches are not
    disturbed): ready / start / stop / quit over the pipes, with a stand-in pynvml (no GPU here)"""
    import json
    import subprocess
    import sys
    import textwrap
    import time
    import bench
    (tmp_path / "pynvml.py").write_text(textwrap.dedent('''
        NVML_CLOCK_SM = 1
        def nvmlInit(): pass
        def nvmlDeviceGetHandleByUUID(u): return 1
        def nvmlDeviceGetHandleByIndex(i): return 1
        def nvmlDeviceGetMaxClockInfo(h, c): return 1965
        def nvmlDeviceGetClockInfo(h, c): return 1950
        def nvmlDeviceGetCurrentClocksEventReasons(h): return 0x4 | 0x40
    '''))
    env = dict(os.environ, PYTHONPATH=str(tmp_path))
    p = subprocess.Popen([sys.executable, "-c", bench._SAMPLER_CHILD, "GPU-test", "0"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1, env=env)
    try:
        assert p.stdout.readline().strip() == "ready"
        p.stdin.write("start\n"); p.stdin.flush()
        time.sleep(0.25)
        p.stdin.write("stop\n"); p.stdin.flush()
        d = json.loads(p.stdout.readline())
        assert len(d["sm"]) >= 3 and set(d["sm"]) == {1950.0} and d["mx"] == 1965.0
        assert d["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
        p.stdin.write("start\n"); p.stdin.flush()        # a second window starts empty
        p.stdin.write("stop\n"); p.stdin.flush()
        assert len(json.loads(p.stdout.readline())["sm"]) <= 2
        p.stdin.write("quit\n"); p.stdin.flush()
        assert p.wait(5) == 0
    finally:
        if p.poll() is None:
            p.kill()
    # without a GPU the wrapper reports "not available" and bench falls back to the in-process sampler
    h = bench.ClockSamplerProcess(0)
    assert not h.ok()
    h.close()


def test_shards_tile_the_table_and_reduction_rate_block():
    """host arithmetic behind the peer-memory step and the bench line: shard_bounds tiles [0, n) for every world size, the largest
    and smallest shard differ by at most one entry; bench.reduction_rate_block scales a counted batch to a launch"""
    from seal3d_b200.parallel import shard_bounds
    import bench
    for n in (0, 1, 7, 6119864, 50021):
        for world in (1, 2, 3, 4, 8, 16):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 0
    b = bench.reduction_rate_block((400, 10, 300), 1000.0, 2.0)       # 40 reductions per sample, 1000 samples, 2 ms
    assert b["reductions_per_sample"] == 40 and b["hashed_level_reductions_per_sample"] == 30 and b["reductions_per_launch"] == 40000
    assert abs(b["achieved"] - 40000 / 2e-3 / 1e9) < 1e-12 and b["peak"] == bench.RED_RATE_PAIRED_G
    assert abs(b["frac"] - (40000 / (bench.RED_RATE_PAIRED_G * 1e9) * 1e3) / 2.0) < 1e-12


def test_host_pointer_arrays_for_the_peer_entry_points():
    import numpy as np
    import torch
    from seal3d_b200 import _lib
    t = [torch.zeros(4), torch.zeros(8)]
    arr, addr = _lib.host_ptrs(t + [12345])
    assert arr.dtype == np.uint64 and list(arr) == [t[0].data_ptr(), t[1].data_ptr(), 12345] and addr == arr.ctypes.data
