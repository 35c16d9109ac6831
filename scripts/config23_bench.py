"""Rates of BASELINE.json configs 2 and 3 on one GPU (development / measurement tool; the bench line of bench.py is config 4's
distillation step):

  config 2   main_nerf.py Lego NGP -O: training rays/s of the photometric step (finetune_step against fixed target images,
             no teacher in the step) at N = 4 096 (the reference's --num_rays default) and N = 2^18, and full-image render
             rays/s of one 800 x 800 view (640 000 rays): the single-pass render (renderer.render_single_pass through the fused
             field) and the reference-style host loop (run_cuda eval branch, nerf/renderer.py:323-372) through the same field
  config 3   main_SealNeRF.py bbox edit: pretraining points/s (pretrain_step, tables only, SealNeRF/trainer.py:456-469) at
             2^21 points per step, fine-tuning rays/s at N = 4 096 against teacher-rendered targets, and the teacher
             proxy_dataset render rate (SealNeRF/provider.py:19-70: one proxied 800 x 800 view)

    python scripts/config23_bench.py [--out gpurun_out/r2/config23.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2", "config23.json"))
    args = ap.parse_args()
    import bench
    from seal3d_b200 import synth
    from seal3d_b200.fused import FusedDistillTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = {}

    def rate(fn, units, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        return {"ms": ms, "units_per_call": units, "rate_per_s": units / ms * 1e3}

    teacher, student = bench.build_world(dev, "fp16")
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, update_interval=16)
    # ---- config 2: photometric training ---------------------------------------------------------------------------------
    for N in (4096, 1 << 18):
        batches = []
        for b in range(4):
            o, d = synth.rays_for_step(100 + b, N)
            batches.append((to(o), to(d), torch.rand(N, 3, device=dev), torch.rand(N, device=dev)))
        for i in range(4):
            tr.finetune_step(*batches[i % 4], perturb=True, force_all_rays=(i < 2))
        tr.refresh_occupancy()
        it = [0]

        def step():
            o, d, im, dp = batches[it[0] % 4]
            it[0] += 1
            tr.finetune_step(o, d, im, dp, perturb=True)
        r = rate(step, N, 32 if N <= 4096 else 16)
        r["samples_per_step"] = float(tr.student.step_counter[:, 0].float().max().item())
        res["config2_train_rays_%d" % N] = r
        print("config 2 train N=%d: %.3f ms/step  %.2f M rays/s" % (N, r["ms"], r["rate_per_s"] / 1e6), flush=True)
    # ---- config 2: full-image render ------------------------------------------------------------------------------------
    o, d = synth.full_image_rays(0)
    o, d = to(o), to(d)
    n_img = o.shape[0]
    student.eval()
    r = rate(lambda: tr.S.render_image(o, d), n_img, 5, warm=2)
    res["config2_render_single_pass"] = r
    print("config 2 render 800x800 single pass: %.2f ms  %.2f M rays/s" % (r["ms"], r["rate_per_s"] / 1e6), flush=True)
    orig = student._field
    student._field = tr.S.field
    try:
        with torch.no_grad():
            r = rate(lambda: student.run_cuda(o.view(1, -1, 3), d.view(1, -1, 3), perturb=False), n_img, 3, warm=1)
    finally:
        student._field = orig
    res["config2_render_reference_style_loop"] = r
    print("config 2 render 800x800 host loop (nerf/renderer.py:323-372 schedule): %.2f ms  %.2f M rays/s" % (r["ms"], r["rate_per_s"] / 1e6), flush=True)
    # ---- config 3: proxy_dataset render (teacher, proxy-mapped), pretraining, fine-tuning ----------------------------------
    teacher.eval()
    Tf = tr.T
    r = rate(lambda: Tf.render_image(o, d), n_img, 5, warm=2)
    res["config3_proxy_dataset_render"] = r
    print("config 3 proxied teacher view (800x800): %.2f ms  %.2f M rays/s" % (r["ms"], r["rate_per_s"] / 1e6), flush=True)
    student.train()
    P = 1 << 21
    g = torch.Generator(device=dev).manual_seed(3)
    pts = torch.rand(P, 3, device=dev, generator=g) * 0.6 - 0.3 + torch.tensor([0.3, 0.0, 0.0], device=dev)
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev, generator=g), dim=-1).contiguous()
    sig_t, rgb_t, _ = Tf.forward(pts, dirs)
    r = rate(lambda: tr.pretrain_step(pts, dirs, sig_t, rgb_t), P, 16)
    res["config3_pretrain_points_2097152"] = r
    print("config 3 pretraining: %.3f ms per 2^21 points  %.1f M points/s" % (r["ms"], r["rate_per_s"] / 1e6), flush=True)
    N = 4096
    out = Tf.render_image(o[:N * 8], d[:N * 8])
    batches = [(o[i * N:(i + 1) * N].contiguous(), d[i * N:(i + 1) * N].contiguous(), out["image"][i * N:(i + 1) * N].contiguous(),
                out["depth"][i * N:(i + 1) * N].contiguous()) for i in range(8)]
    it = [0]

    def ft():
        b = batches[it[0] % 8]
        it[0] += 1
        tr.finetune_step(*b, perturb=True)
    r = rate(ft, N, 32)
    res["config3_finetune_rays_4096"] = r
    print("config 3 fine-tuning N=4096: %.3f ms/step  %.2f M rays/s" % (r["ms"], r["rate_per_s"] / 1e6), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
