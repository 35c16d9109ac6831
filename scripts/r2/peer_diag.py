"""what the peer-memory step's pieces cost (development tool): torchrun --nproc-per-node N scripts/r2/peer_diag.py"""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from seal3d_b200 import parallel, synth, _lib
from seal3d_b200.fused import FusedDistillTrainer
rank, local, world = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
teacher, student = bench.build_world(dev, "fp16")
tr = FusedDistillTrainer(student, teacher, lr=1e-2, world_size=world, update_interval=16)
assert tr.peer is not None
n = 262144
batches = []
for b in range(4):
    o, d = synth.rays_for_step(1000 * rank + b, n)
    batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)))
for i in range(6):
    tr.distill_step(*batches[i % 4], perturb=True)
torch.cuda.synchronize(); dist.barrier()
def t(fn, reps):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return parallel.max_over_ranks(e0.elapsed_time(e1) / reps, dev)
P, S = tr.peer, tr.S
res = {}
res["barrier"] = t(lambda: P["g"].barrier(0), 200)
res["zero_arena"] = t(lambda: S.grad.zero_(), 50)
def kern():
    _lib.call("s3d_ngp_peer_adam_tables", P["grad"][1], P["sigma"][1], P["color"][1], P["shadow"][1], P["world"], P["rank"], S.m4, S.v4,
              S._tbl_stride, P["shard"][0], P["shard"][1], 1e-2, 0.9, 0.99, 1e-15, 5, 1.0)
res["peer_kernel_zero_grads"] = t(kern, 20)
def full():
    tr._peer_step(1.0, True)
res["peer_step_empty_grads"] = t(full, 20)
# full training step with / without the exchange
res["step"] = t(lambda: tr.distill_step(*batches[0], perturb=True), 20)
orig = tr._peer_step
tr._peer_step = lambda scale, train_mlp: S.grad.zero_()
res["step_without_exchange"] = t(lambda: tr.distill_step(*batches[0], perturb=True), 20)
tr._peer_step = orig
if rank == 0:
    print("world", world, {k: round(v, 4) for k, v in res.items()}, flush=True)
dist.barrier(); dist.destroy_process_group()
