"""Timeline of k_ngp_mlp_bwd (CTA 0, iterations 2..7) from clock64 stamps; needs a library built with
S3D_NVCC_EXTRA=-DS3D_TRACE.  Development tool."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    from seal3d_b200 import synth, _lib
    from seal3d_b200.fused import FusedDistillTrainer
    dev = torch.device("cuda", 0)
    teacher, student = bench.build_world(dev, "fp16")
    tr = FusedDistillTrainer(student, teacher, lr=1e-2, world_size=1, update_interval=16)
    o, d = synth.rays_for_step(0, 65536)
    o, d = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    for i in range(4):
        tr.distill_step(o, d, perturb=True, force_all_rays=(i < 2))
    torch.cuda.synchronize()
    lib = _lib.lib()
    n = 6 * 3 * 64
    buf = (ctypes.c_longlong * n)()
    lib.s3d_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = lib.s3d_debug_trace(buf, n)
    t = np.array(buf, dtype=np.int64).reshape(6, 3, 64)
    print("rc", rc)
    names = {0: "issuer", 1: "fwd-set", 2: "bwd-set"}
    for it in range(6):
        t0 = t[it][t[it] > 0].min()
        print("== iteration", it + 2, " length", int(t[it].max() - t0))
        for role in range(3):
            ev = [(k, int(t[it, role, k] - t0)) for k in range(64) if t[it, role, k] > 0]
            print("  %-8s" % names[role], " ".join("%d:%d" % e for e in ev))
    # per-round latencies, averaged over iterations 3..6
    def avg(f):
        return float(np.mean([f(t[it]) for it in range(1, 5)]))
    for k in range(5):
        print("round %d: F publish->issuer sync %5.0f  issue+commit %5.0f  commit->F wake %5.0f | F epilogue (wake->next publish) %5.0f" % (
            k, avg(lambda x: x[0, 4 * k] - x[1, 2 * k]), avg(lambda x: x[0, 4 * k + 1] - x[0, 4 * k]), avg(lambda x: x[1, 2 * k + 1] - x[0, 4 * k + 1]),
            avg(lambda x: (x[1, 2 * k + 2] if k < 4 else x[1, 20]) - x[1, 2 * k + 1])))
        print("         B publish->issuer sync %5.0f  issue+commit %5.0f  commit->B wake %5.0f | B epilogue (wake->next publish) %5.0f" % (
            avg(lambda x: x[0, 4 * k + 2] - x[2, 2 * k]), avg(lambda x: x[0, 4 * k + 3] - x[0, 4 * k + 2]), avg(lambda x: x[2, 2 * k + 1] - x[0, 4 * k + 3]),
            avg(lambda x: (x[2, 2 * k + 2] if k < 4 else x[2, 20]) - x[2, 2 * k + 1])))


if __name__ == "__main__":
    main()
