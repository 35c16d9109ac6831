"""2-rank check on real GPUs: two ranks, each with half of a ray batch, end in the same parameters (bitwise identical
replicas) and close to a single process that ran the whole batch (fp32 reduction order only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from seal3d_b200 import parallel, synth  # noqa: E402
from seal3d_b200.fused import FusedDistillTrainer  # noqa: E402


def run(dev, world, rank, steps=4, n=8192):
    teacher, student = bench.build_world(dev, "fp16")
    # loss_scale fixed so that both configurations scale gradients identically
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, loss_scale=32.0 * n, world_size=world, update_interval=0)
    captured = []
    orig = tr.S.adam_step

    def capture(lr, grad_scale=1.0, **kw):
        if not captured:
            captured.append((tr.S.grad.clone() * grad_scale))     # the all-reduced, unscaled gradient of step 0
        return orig(lr, grad_scale=grad_scale, **kw)

    tr.S.adam_step = capture
    for i in range(steps):
        o, d = synth.rays_for_step(i, n)
        lo, hi = parallel.shard_bounds(n, rank, world)
        tr.distill_step(torch.from_numpy(o[lo:hi]).to(dev), torch.from_numpy(d[lo:hi]).to(dev), perturb=False, force_all_rays=True)
    return student.encoder.embeddings.detach().clone(), student.sigma_net[0].weight.detach().clone(), captured[0]


def main():
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    e, w, g = run(dev, world, rank)
    # replicas identical?
    e0 = e.clone()
    torch.distributed.broadcast(e0, 0)
    same = bool(torch.equal(e0, e))
    flag = torch.tensor([1.0 if same else 0.0], device=dev)
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        # NOTE: the mean over rays differs (each rank normalises by its own N): single-process reference uses the same
        # per-shard normalisation by running the two shards as world=1 with gradient accumulation disabled -> compare
        # against a run of world=1 on the full batch scaled accordingly is not bitwise meaningful; report replica identity
        # and the distance to the full-batch run.
        e1, w1, g1 = run(dev, 1, 0)
        gd = (g1 - g).abs().max().item() / g1.abs().max().item()
        print("replicas_identical=%s  step-0 gradient: max|full_batch - dp2| / max|grad| = %.3e  (parameters after 4 Adam steps differ by %.3e: "
              "Adam with eps=1e-15 turns rounding-level gradient differences of near-zero entries into +-lr steps)" % (bool(flag.item()), gd, (e1 - e).abs().max().item()))
        assert bool(flag.item()) and gd < 1e-3
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
