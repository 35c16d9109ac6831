"""diagnostic: why is the pipelined schedule slow when the host does not sync every step?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from seal3d_b200 import synth
from seal3d_b200.fused import FusedDistillTrainer
dev = torch.device("cuda", 0)
teacher, student = bench.build_world(dev, "fp16")
tr = FusedDistillTrainer(student, teacher, lr=1e-2, world_size=1, update_interval=16)
res = []
for b in range(4):
    o, d = synth.rays_for_step(b, 262144)
    res.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)))
for i in range(5):
    tr.distill_step(*res[i % 4], perturb=True, force_all_rays=(i < 2))
if tr.student.mean_count <= 0:
    tr.refresh_occupancy()
torch.cuda.synchronize()
for mode in ("inline", "ahead", "ahead+sync"):
    st0 = torch.cuda.memory_stats()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    cpu = []
    torch.cuda.synchronize()
    evs[0].record()
    for i in range(20):
        t0 = time.perf_counter()
        nxt = res[(i + 1) % 4] if mode != "inline" and i < 19 else None
        l = tr.distill_step(*res[i % 4], perturb=True, prefetch=nxt)
        if mode == "ahead+sync":
            l.cpu()
        cpu.append((time.perf_counter() - t0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats()
    gpu = [evs[i].elapsed_time(evs[i + 1]) for i in range(20)]
    print(mode, "gpu ms/step", ["%.1f" % g for g in gpu], "cpu enqueue ms", ["%.1f" % c for c in cpu])
    print("   device allocs +%d, retries +%d, reserved %.1f GB" % (st1["num_device_alloc"] - st0["num_device_alloc"], st1["num_alloc_retries"] - st0["num_alloc_retries"], st1["reserved_bytes.all.current"] / 2**30))
