"""Data-parallel plumbing: one process per GPU, rays sharded across ranks, ONE all-reduce of the gradient
arena per step.  The reference has no working multi-GPU path (dead DDP code,
nerf/utils.py:330-332), so this is new functionality; W=1 must equal the single-process result and W>1 the
single-process result on the concatenated batch up to fp32 reduction order."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT) -> (rank, local_rank, world)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of n units for `rank`; the first n % world ranks get one extra unit"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(flat):
    """the step's single collective: in-place sum of the flat gradient arena over all ranks"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
    return flat


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class PeerBuffer:
    """One symmetric-memory allocation per rank + rendezvous (torch.distributed._symmetric_memory: cuMem allocations whose
    handles the ranks exchange and map into their own address space): `.local` is this rank's tensor, `.ptrs[r]` the device address
    of rank r's copy as seen from THIS process -- kernels launched here load and store through them over NVLink.  `.barrier()`
    enqueues a cross-rank barrier on the current stream (signal pads in the same allocation, system-scope release / acquire), so
    peer reads and writes of consecutive kernels are ordered without a host round trip or an NCCL launch.  Collective: every
    rank of the group constructs its PeerBuffers in the same order."""

    def __init__(self, numel, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.local = symm.empty(int(numel), dtype=dtype, device=device)
        try:
            self.hdl = symm.rendezvous(self.local, group)
        except TypeError:
            self.hdl = symm.rendezvous(self.local, group.group_name)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        if len(self.ptrs) != self.world or self.ptrs[self.rank] != self.local.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned unexpected pointers")

    def barrier(self, channel=0):
        self.hdl.barrier(channel=channel)
