"""BASELINE config 5 in numbers: TensoRF teacher -> TensoRF student distillation steps (colour edit in a bbox) on one B200,
through trainer.DistillTrainer (VM lookup kernels with fp16 colour features, basis_mat on s3d_linear_*, one-launch frequency
encodings, the 150-128-128-3 MLP on the wide FFMLP kernels, marcher / compositor / proxy / loss / Adam kernels).  Prints one JSON
line; "torch_gemm_and_glue_ms" = step time minus the time inside this library's launches.
Development / measurement tool; the headline bench (bench.py) stays on the NGP backbone the metric is quoted on."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from seal3d_b200 import synth, _lib
    from seal3d_b200.seal import TensoRFTeacherNetwork, TensoRFStudentNetwork, SealBBoxMapper
    from seal3d_b200.trainer import DistillTrainer
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    res = int(os.environ.get("VM_RES", 300))
    n = int(os.environ.get("RAYS", 65536))
    prec = os.environ.get("PRECISION", "fp16")
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    teacher = TensoRFTeacherNetwork(resolution=[res] * 3, bound=1, density_thresh=10).to(dev)
    student = TensoRFStudentNetwork(resolution=[res] * 3, bound=1, density_thresh=10).to(dev)
    bits, grid = synth.lego_like_occupancy()
    md, tris = synth.bbox_edit(hsv=[0.3, 0.0, 0.0])
    mapper = SealBBoxMapper(md, tris, device=dev)
    for net in (teacher, student):
        net.density_bitfield.copy_(to(bits))
        net.density_grid.copy_(to(grid))
        net.init_mapper(mapper)
        net.hack_bitfield()
    teacher.eval()
    tr = DistillTrainer(student, teacher, lr=(2e-2, 1e-3), precision=prec, loss_scale=(128.0 if prec == "fp16" else 1.0), update_interval=0)
    batches = [tuple(to(a) for a in synth.rays_for_step(b, n)) for b in range(4)]
    for i in range(4):
        tr.distill_step(*batches[i % 4], perturb=True, force_all_rays=True)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    steps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = tr.distill_step(*batches[i % 4], perturb=True, force_all_rays=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    agg = {}
    for name, a, b in _lib.PROFILE:
        agg[name] = agg.get(name, 0.0) + a.elapsed_time(b) / steps
    _lib.PROFILE = None
    M = float(student.step_counter[:, 0].float().max().item())
    ours = sum(agg.values())
    print(json.dumps({"config": "tensorf-distill-step/res%d/%s" % (res, prec), "rays_per_step": n, "samples_per_step": M, "ms_per_step": round(ms, 3),
                      "rays_per_s": round(n / ms * 1e3), "loss": [float(v) for v in loss.cpu()],
                      "our_kernels_ms": round(ours, 3), "torch_gemm_and_glue_ms": round(ms - ours, 3),
                      "kernel_breakdown_ms": {k.replace("s3d_", ""): round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01}}))


if __name__ == "__main__":
    main()
