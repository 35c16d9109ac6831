"""Step-time experiments on the benchmark's distillation step under the current environment knobs (development tool).
   python scripts/r2/stepbench.py [--rays N] [--steps K] [--tag text] [--no-prefetch]
Prints: tag, ms/step (CUDA events around K pipelined steps), and the per-launch breakdown of 3 un-pipelined steps."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--tag", default="")
    ap.add_argument("--no-prefetch", action="store_true")
    ap.add_argument("--breakdown", action="store_true")
    args = ap.parse_args()
    from seal3d_b200 import synth, _lib
    from seal3d_b200.fused import FusedDistillTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    teacher, student = bench.build_world(dev, "fp16")
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, world_size=1, update_interval=16)
    res = []
    for b in range(4):
        o, d = synth.rays_for_step(b, args.rays)
        res.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)))
    for i in range(5):
        tr.distill_step(*res[i % 4], perturb=True, force_all_rays=(i < 2))
    pre = not args.no_prefetch
    for i in range(3):
        tr.distill_step(*res[i % 4], perturb=True, prefetch=res[(i + 1) % 4] if (pre and i < 2) else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        tr.distill_step(*res[i % 4], perturb=True, prefetch=res[(i + 1) % 4] if (pre and i + 1 < args.steps) else None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    line = "%-40s %.3f ms/step  %.2f M rays/s  loss %s" % (args.tag, ms, args.rays / ms / 1e3, tr.loss_buf.cpu().numpy())
    if args.breakdown:
        _lib.PROFILE = []
        for i in range(3):
            tr.distill_step(*res[i % 4], perturb=True)
        torch.cuda.synchronize()
        agg = {}
        for name, a, b in _lib.PROFILE:
            agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) / 3
        _lib.PROFILE = None
        line += "  | " + " ".join("%s=%.3f" % (k.replace("s3d_", ""), v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.02)
    print(line, flush=True)


if __name__ == "__main__":
    main()
