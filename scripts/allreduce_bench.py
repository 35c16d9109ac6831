"""All-reduce of the gradient arena (24.5 M values) over NCCL: fp32 (98 MB, what the trainer does) vs bf16 / fp16 (49 MB), CUDA-event
timed on every rank, max over ranks.  torchrun --nproc-per-node N scripts/allreduce_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from seal3d_b200 import parallel

rank, local, world = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
n = 6119864 * 4 + 12496
for dt in (torch.float32, torch.bfloat16, torch.float16):
    x = torch.randn(n, device=dev).to(dt)
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / 20, dev)
    if rank == 0:
        b = x.numel() * x.element_size()
        print("all_reduce %-8s %6.1f MB  world %d  %.3f ms  busbw %.1f GB/s" % (str(dt).replace("torch.", ""), b / 1e6, world, ms, 2 * (world - 1) / world * b / ms / 1e6), flush=True)
dist.barrier()
dist.destroy_process_group()
