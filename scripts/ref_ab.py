"""Same-box A/B of the UNMODIFIED reference kernels (oracle/_ref, built by oracle/build_ref.py) against this repo's kernels, op
by op on identical inputs, and of the composed reference `-O` distillation step (tests/refstep.RefDistillStep) against the
fused trainer at the benchmark's batch -- BASELINE.md's "R" column.  Development / measurement tool (GPU box only):

    python scripts/ref_ab.py [--rays 262144] [--out gpurun_out/r2/ref_vs_ours.json]

Every op is called at the extension-module level with pre-allocated outputs (`_ref_gridencoder.grid_encode_forward(...)` vs
`_gridencoder.grid_encode_forward(...)`: same signature, same tensors), timed with CUDA events on the current stream after
warm-up, an L2 flush (256 MB memset) before every timed launch, median of the timed launches."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=262144)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2", "ref_vs_ours.json"))
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    import bench
    import refext
    import refstep
    from seal3d_b200 import synth, _lib
    from seal3d_b200.fused import FusedDistillTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, iters=7, warm=2, setup=None):
        ts = []
        for it in range(warm + iters):
            if setup:
                setup()
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it >= warm:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    ours = {n: refstep._ours(n) for n in ("gridencoder", "raymarching", "shencoder", "ffmlp")}
    ref = {n: refext.load(n) for n in ("gridencoder", "raymarching", "shencoder", "ffmlp")}
    rows = []

    def ab(name, units, unit_name, make):
        """make(mod) -> (fn, setup or None) for either backend"""
        t = {}
        for who, mods in (("reference", ref), ("ours", ours)):
            fn, setup = make(mods)
            t[who] = timeit(fn, setup=setup)
        row = {"op": name, "units": units, "unit": unit_name, "reference_ms": t["reference"], "ours_ms": t["ours"], "speedup": t["reference"] / t["ours"]}
        rows.append(row)
        print("%-44s ref %8.3f ms   ours %8.3f ms   x%.2f" % (name, t["reference"], t["ours"], row["speedup"]), flush=True)

    # ---- hash grid, 2^22 random points (BASELINE config 1 at roofline size) --------------------------------------------
    offsets_np, pls = synth.grid_offsets()
    offsets = torch.from_numpy(offsets_np).to(dev)
    S = float(np.log2(pls))
    B = 1 << 22
    g = torch.Generator(device=dev).manual_seed(8)
    x = torch.rand(B, 3, device=dev, generator=g)
    emb32 = torch.rand(int(offsets_np[-1]), 2, device=dev, generator=g) * 2e-4 - 1e-4
    for dt, tag in ((torch.float32, "fp32"), (torch.float16, "fp16")):
        emb = emb32.to(dt)
        out = torch.empty(16, B, 2, device=dev, dtype=dt)
        grad = torch.randn(16, B, 2, device=dev, generator=g).to(dt)
        gemb = torch.zeros_like(emb)
        ab("grid_encode_forward %s 2^22" % tag, B, "points", lambda m: (lambda: m["gridencoder"].grid_encode_forward(x, emb, offsets, out, B, 3, 2, 16, S, 16, None, 0, False, 0), None))
        ab("grid_encode_backward %s 2^22" % tag, B, "points", lambda m: (lambda: m["gridencoder"].grid_encode_backward(grad, x, emb, offsets, gemb, B, 3, 2, 16, S, 16, None, None, 0, False, 0), gemb.zero_))
    del out, grad, gemb, x

    # ---- marcher / compositor / SH on the benchmark's ray batch ----------------------------------------------------------
    bits_np, _ = synth.lego_like_occupancy()
    bits = torch.from_numpy(bits_np).to(dev)
    o_np, d_np = synth.rays_for_step(0, args.rays)
    o, d = torch.from_numpy(o_np).to(dev), torch.from_numpy(d_np).to(dev)
    N = args.rays
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=dev)
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    ab("near_far_from_aabb", N, "rays", lambda m: (lambda: m["raymarching"].near_far_from_aabb(o, d, aabb, N, 0.2, nears, fars), None))
    from seal3d_b200 import raymarching as rm
    xs, ds, dl, rays = rm.march_rays_train(o, d, 1.0, bits, 1, 128, nears, fars, None, -1, True, 128, True)
    M = xs.shape[0]
    noises = torch.rand(N, device=dev)
    xyzs, dirs, deltas = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev), torch.zeros(M, 2, device=dev)
    rays_t = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    ab("march_rays_train (M = %d)" % M, N, "rays", lambda m: (lambda: m["raymarching"].march_rays_train(o, d, bits, 1.0, 0.0, 1024, N, 1, 128, M, nears, fars, xyzs, dirs, deltas, rays_t, counter, noises), counter.zero_))
    sig = torch.rand(M, device=dev) * 20
    rgb = torch.rand(M, 3, device=dev)
    ws, dep, img = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
    ab("composite_rays_train_forward", M, "samples", lambda m: (lambda: m["raymarching"].composite_rays_train_forward(sig, rgb, dl, rays, M, N, 1e-4, ws, dep, img), None))
    gws, gimg = torch.randn(N, device=dev), torch.randn(N, 3, device=dev)
    gsig, grgb = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
    ab("composite_rays_train_backward", M, "samples", lambda m: (lambda: m["raymarching"].composite_rays_train_backward(gws, gimg, sig, rgb, dl, rays, ws, img, M, N, 1e-4, gsig, grgb), None))
    sh = torch.empty(M, 16, device=dev)
    ab("sh_encode_forward degree 4", M, "samples", lambda m: (lambda: m["shencoder"].sh_encode_forward(ds, sh, M, 3, 4, None), None))
    del sh

    # ---- FFMLP 2^21 rows, 32 -> 64 -> 64 -> 16 (testing/test_ffmlp.py shape) ----------------------------------------------
    Bm = 1 << 21
    nl, hid = 2, 64
    w = ((torch.rand(hid * (32 + hid * (nl - 1) + 16), device=dev, generator=g) * 2 - 1) * (3 / 64) ** 0.5).half()
    xin = torch.rand(Bm, 32, device=dev, generator=g).half()
    fbuf = torch.empty(nl, Bm, hid, device=dev, dtype=torch.float16)
    yout = torch.empty(Bm, 16, device=dev, dtype=torch.float16)
    ab("ffmlp_forward 2^21 x (32-64-64-16)", Bm, "rows", lambda m: (lambda: m["ffmlp"].ffmlp_forward(xin, w, Bm, 32, 16, hid, nl, 0, 6, fbuf, yout), None))
    ref["ffmlp"].allocate_splitk(nl + 1)
    gy = torch.randn(Bm, 16, device=dev, generator=g).half()
    bbuf = torch.zeros(nl, Bm, hid, device=dev, dtype=torch.float16)
    gin = torch.zeros(1, device=dev, dtype=torch.float16)
    gw = torch.zeros_like(w)
    ab("ffmlp_backward 2^21 x (32-64-64-16)", Bm, "rows", lambda m: (lambda: m["ffmlp"].ffmlp_backward(gy, xin, w, fbuf, Bm, 32, 16, hid, nl, 0, 6, False, bbuf, gin, gw), None))
    del xin, fbuf, yout, gy, bbuf

    # ---- the composed step ---------------------------------------------------------------------------------------------
    teacher, student = bench.build_world(dev, "fp16")
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, update_interval=0)
    batches = []
    for b in range(4):
        oo, dd = synth.rays_for_step(b, N)
        batches.append((torch.from_numpy(oo).to(dev), torch.from_numpy(dd).to(dev)))
    for i in range(3):
        tr.distill_step(*batches[i % 4], perturb=True, force_all_rays=(i < 2))
    tr.refresh_occupancy()
    mean_count = int(tr.student.mean_count)
    for i in range(3):
        tr.distill_step(*batches[i % 4], perturb=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        tr.distill_step(*batches[i % 4], perturb=True)
    e1.record()
    torch.cuda.synchronize()
    ours_ms = e0.elapsed_time(e1) / args.steps
    bits_now = tr.student.density_bitfield.clone()
    rstep = refstep.RefDistillStep(synth.field_params("teacher"), synth.field_params("student"), bits_now, dev, lr=1e-2, map_samples=teacher._map_samples)
    for i in range(3):
        rstep.step(*batches[i % 4], mean_count)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        rstep.step(*batches[i % 4], mean_count)
    e1.record()
    torch.cuda.synchronize()
    ref_ms = e0.elapsed_time(e1) / args.steps
    step = {"rays_per_step": N, "samples_budget": mean_count, "reference_ms_per_step": ref_ms, "ours_ms_per_step": ours_ms,
            "reference_rays_per_s": N / ref_ms * 1e3, "ours_rays_per_s": N / ours_ms * 1e3, "speedup": ref_ms / ours_ms,
            "reference_step": "reference extensions (oracle/_ref) through the reference's own wrappers + torch autocast nn.Linear + autograd + GradScaler + torch.optim.Adam (tests/refstep.RefDistillStep); proxy mapping by this repo's kernel",
            "ours_step": "fused.FusedDistillTrainer.distill_step, inline march (no pipelining), no occupancy refresh inside the timed steps"}
    print("distillation step @ %d rays: reference %.2f ms (%.2f M rays/s)   ours %.2f ms (%.2f M rays/s)   x%.2f" %
          (N, ref_ms, N / ref_ms / 1e3, ours_ms, N / ours_ms / 1e3, ref_ms / ours_ms), flush=True)
    prop = torch.cuda.get_device_properties(dev)
    res = {"device": prop.name, "sm_count": prop.multi_processor_count, "ops": rows, "step": step,
           "protocol": "CUDA events on the launching stream, 256 MB L2 flush before every timed launch, median of 7 after 2 warm-ups; step: mean of %d steps after 3 warm-ups" % args.steps}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        res["clocks"] = {"sm_mhz": pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)}
    except Exception:
        pass
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
