"""Multi-GPU check of the peer-memory optimizer step (fused.FusedDistillTrainer._peer_step) against the all-reduce path, on real GPUs:
    torchrun --nproc-per-node N scripts/dp_peer_check.py
Both paths run the same ray shards from the same state with fixed-point (deterministic) gradient accumulation, so the only
difference left is the cross-rank summation: for N = 2 that is a + b either way and the parameters must agree BIT FOR BIT; for
N > 2 NCCL's order is its own and the comparison is a tolerance.  In both cases the replicas of the peer path must be identical
and the gathered Adam moments must equal the all-reduce path's.  Then both paths are timed on the benchmark's batch size."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from seal3d_b200 import parallel, synth  # noqa: E402
from seal3d_b200.fused import FusedDistillTrainer  # noqa: E402


def run(dev, world, rank, peer, steps=4, n=8192, deterministic=True, timing=False):
    os.environ["S3D_PEER_ADAM"] = "1" if peer else "0"
    teacher, student = bench.build_world(dev, "fp16")
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, loss_scale=32.0 * n, world_size=world, update_interval=0 if not timing else 16,
                             deterministic=deterministic)
    assert (tr.peer is not None) == bool(peer), "peer mode requested=%s active=%s" % (peer, tr.peer is not None)
    if timing:
        batches = []
        for b in range(4):
            o, d = synth.rays_for_step(1000 * rank + b, n)
            batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)))
        for i in range(6):
            tr.distill_step(*batches[i % 4], perturb=True, prefetch=batches[(i + 1) % 4])
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            tr.distill_step(*batches[(i + 2) % 4], perturb=True, prefetch=batches[(i + 3) % 4] if i + 1 < steps else None)
        e1.record(); torch.cuda.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    for i in range(steps):
        o, d = synth.rays_for_step(i, n)
        lo, hi = parallel.shard_bounds(n, rank, world)
        tr.distill_step(torch.from_numpy(o[lo:hi]).to(dev), torch.from_numpy(d[lo:hi]).to(dev), perturb=False, force_all_rays=True)
    tr.gather_optimizer_state()
    S = tr.S
    return [student.encoder.embeddings.detach().clone(), student.encoder_color.embeddings.detach().clone(), S.mlp32.clone(), S.m4.clone(), S.v4.clone(),
            (tr.table8 if tr.table8 is not None else S.table4).clone().view(torch.int32)]


def all_ranks_equal(t, dev):
    t0 = t.clone()
    dist.broadcast(t0, 0)
    f = torch.tensor([1.0 if torch.equal(t0, t) else 0.0], device=dev)
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    return bool(f.item())


def main():
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    names = ["sigma table", "colour table", "mlp", "exp_avg", "exp_avg_sq", "fp16 table"]
    a = run(dev, world, rank, peer=False)
    b = run(dev, world, rank, peer=True)
    ok = True
    for n_, x, y in zip(names, a, b):
        same_ranks = all_ranks_equal(y, dev)
        xf, yf = x.float(), y.float()
        bit = bool(torch.equal(x, y))
        rel = float((xf - yf).abs().max() / xf.abs().max().clamp_min(1e-30))
        if rank == 0:
            print("%-14s replicas identical: %-5s  peer == all-reduce bitwise: %-5s  max rel diff %.3e" % (n_, same_ranks, bit, rel), flush=True)
        ok = ok and same_ranks and (bit if world == 2 else True)
    t_nccl = run(dev, world, rank, peer=False, steps=20, n=262144, deterministic=False, timing=True)
    t_peer = run(dev, world, rank, peer=True, steps=20, n=262144, deterministic=False, timing=True)
    if rank == 0:
        print("world %d  262144 rays/rank: all-reduce path %.3f ms/step (%.1f M rays/s)   peer path %.3f ms/step (%.1f M rays/s)" %
              (world, t_nccl, world * 262144 / t_nccl / 1e3, t_peer, world * 262144 / t_peer / 1e3), flush=True)
        print("PEER CHECK", "OK" if ok else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
