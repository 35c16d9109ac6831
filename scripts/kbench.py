"""Kernel-variant experiments on the benchmark's own distillation step (GPU only; development tool, not part of the product).

    python scripts/kbench.py --rays 262144 --set 0:0,1,2 --set 1:0,1,2

For every value of every knob (s3d_debug_variant(which, value), other knobs at their defaults) it runs a few steps and
prints the per-kernel time from the CUDA events _lib.PROFILE records around each C-ABI launch."""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402


def measure(tr, resident, nprof=4):
    from seal3d_b200 import _lib
    for i in range(2):
        tr.distill_step(*resident[i % len(resident)], perturb=True)
    _lib.PROFILE = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(nprof):
        tr.distill_step(*resident[i % len(resident)], perturb=True)
    torch.cuda.synchronize()
    agg = {}
    for name, a, b in _lib.PROFILE:
        agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) / nprof
    _lib.PROFILE = None
    e0.record()
    for i in range(8):
        tr.distill_step(*resident[i % len(resident)], perturb=True)
    e1.record()
    torch.cuda.synchronize()
    agg["step_ms"] = e0.elapsed_time(e1) / 8
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=262144)
    ap.add_argument("--set", action="append", default=[], help="which:v0,v1,...")
    ap.add_argument("--defaults", default="", help="which=value,... applied before every measurement")
    args = ap.parse_args()
    from seal3d_b200 import synth, _lib
    from seal3d_b200.fused import FusedDistillTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    teacher, student = bench.build_world(dev, "fp16")
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, world_size=1, update_interval=16)
    resident = []
    for b in range(4):
        o, d = synth.rays_for_step(b, args.rays)
        resident.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)))
    for i in range(4):
        tr.distill_step(*resident[i % 4], perturb=True, force_all_rays=(i < 2))
    if tr.student.mean_count <= 0:
        tr.refresh_occupancy()
    lib = _lib.lib()
    has_knobs = hasattr(lib, "s3d_debug_variant")   # only present in experiment builds
    if has_knobs:
        lib.s3d_debug_variant.argtypes = [ctypes.c_int, ctypes.c_int]
    defaults = dict((int(k), int(v)) for k, v in (kv.split("=") for kv in args.defaults.split(",") if kv))
    keys = ("s3d_ngp_encode_pair", "s3d_ngp_scatter", "s3d_ngp_mlp_backward", "s3d_ngp_mlp_forward", "s3d_march_rays_train", "step_ms")

    def run(label):
        r = measure(tr, resident)
        print(json.dumps({"variant": label, **{k.replace("s3d_", ""): round(r.get(k, 0.0), 4) for k in keys}}), flush=True)

    g = {p: bench.roofline_grid_encode(dev, p) for p in ("fp32", "fp16")}
    print(json.dumps({"grid_forward": {p: {"ms": round(g[p]["launch_ms"], 4), "frac": round(g[p]["frac"], 4)} for p in g}}), flush=True)
    for k, v in defaults.items():
        if has_knobs:
            lib.s3d_debug_variant(k, v)
    run("defaults %s" % defaults)
    for spec in (args.set if has_knobs else []):
        which, vals = spec.split(":")
        for v in vals.split(","):
            lib.s3d_debug_variant(int(which), int(v))
            run("%s=%s" % (which, v))
        # back to the default of this knob
        lib.s3d_debug_variant(int(which), defaults.get(int(which), 0))


if __name__ == "__main__":
    main()
