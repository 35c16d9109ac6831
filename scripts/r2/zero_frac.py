"""how many samples of the benchmark step reach the gradient scatter with an all-zero feature gradient (rays that terminated)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from seal3d_b200 import synth, _lib
from seal3d_b200.fused import FusedDistillTrainer
dev = torch.device("cuda", 0)
teacher, student = bench.build_world(dev, "fp16")
tr = FusedDistillTrainer(student, teacher, lr=1e-2, update_interval=16)
orig = _lib.call
def spy(name, *a):
    if name == "s3d_ngp_scatter":
        df = a[1]
        z = (df == 0).all(dim=1).float().mean().item()
        zl = (df.view(-1, 2, 16, 2) == 0).all(dim=3).all(dim=1).float().mean().item()
        print("samples %d  all-zero gradient rows %.4f  zero (sample, level) pairs %.4f" % (df.shape[0], z, zl), flush=True)
    return orig(name, *a)
_lib.call = spy
import seal3d_b200.fused as fz
fz._lib.call = spy
for i in range(20):
    o, d = synth.rays_for_step(i, 262144)
    tr.distill_step(torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), perturb=True)
