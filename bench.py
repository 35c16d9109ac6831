"""bench.py -- training rays/s of the Seal-3D distillation hot path on B200 (BASELINE.json metric).

A "step" = one teacher->student distillation step on a batch of synthetic Lego-shaped rays
(800x800 cameras, NGP L16/F2 hash grids, analytic occupancy, bbox edit):
    near/far -> march (student occupancy) -> proxy map + teacher field (no grad) -> student field ->
    composite both -> MSE(rgb)+L1(depth) -> backward -> [all-reduce of the gradient arena] -> fused Adam
(+ the density-grid refresh every 16 steps).  Ray batches shard across ranks (weak scaling: every
GPU gets --rays rays per step); the only collective is the per-step gradient all-reduce.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference      # the CPU restatement of the same step on the host cores

Prints ONE JSON line (rank 0).  `value` = rays/s with the ray batches already resident in HBM;
`e2e` = the same through the public trainer call with host (pinned) ray buffers copied in and the
loss read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (rank 0 only) is meant to use every host thread it can,
# and libgomp / the BLAS behind numpy read the variable when they are loaded -- so it is set before numpy is imported
if "reference" in sys.argv[1:]:
    _ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_ncpu)

import numpy as np

# the sample budget M changes a little at every occupancy refresh: round large torch allocations up to 1/16 of a power of
# two so the refreshed buffers reuse the cached blocks instead of going to cudaMalloc inside the timed region
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "roundup_power2_divisions:16")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "seal3d-distill-step/lego800x800-synthetic/ngp-L16-F2-T19/bbox-edit"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=int(os.environ.get("S3D_BENCH_RAYS", 262144)), help="rays per step per GPU")
    ap.add_argument("--precision", default=os.environ.get("S3D_BENCH_PRECISION", "fp16"), choices=["fp32", "fp16"])
    ap.add_argument("--engine", default=os.environ.get("S3D_BENCH_ENGINE", "fused"), choices=["fused", "autograd"],
                    help="fused = csrc/field.cu kernels (tcgen05 MLP, interleaved tables); autograd = op-by-op kernels + torch GEMMs")
    ap.add_argument("--cpu-rays", type=int, default=1024, help="rays per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="march every batch inline instead of one step ahead on a side stream")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ CPU arm


class CpuDistill:
    """The same distillation step composed from the oracle's C / numpy restatement (oracle/), on the host cores."""

    def __init__(self, n_rays):
        import oracle
        from seal3d_b200 import synth
        self.oracle, self.synth, self.n = oracle, synth, n_rays
        self.bits, _ = synth.lego_like_occupancy()
        self.offsets, self.pls = synth.grid_offsets()
        tp, sp = synth.field_params("teacher"), synth.field_params("student")
        mk = lambda p: oracle.NGPField(p["emb_sigma"], p["emb_color"], p["w_s0"], p["w_s1"], p["w_c0"], p["w_c1"], p["w_c2"], self.offsets, self.pls)
        self.teacher, self.student = mk(tp), mk(sp)
        self.md, self.tris = synth.bbox_edit()
        self.params = [self.student.es, self.student.ec] + self.student.w
        self.m = [np.zeros_like(p) for p in self.params]
        self.v = [np.zeros_like(p) for p in self.params]
        self.t = 0
        self.aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)

    def loss_terms(self, i):
        """(MSE, L1 depth) of the unperturbed step i before any update -- what smoke() checks the fused engine against"""
        return self.step(i, perturb=False, update=False)

    def step(self, i, perturb=True, update=True):
        oc = self.oracle
        o, d = self.synth.rays_for_step(i, self.n)
        n, f = oc.near_far_from_aabb(o, d, self.aabb, 0.2)
        noise = np.random.default_rng(i).uniform(0, 1, self.n).astype(np.float32) if perturb else np.zeros(self.n, np.float32)
        # count first (M=0 writes nothing), then allocate exactly
        _, _, _, rays, cnt = oc.march_rays_train(o, d, 1.0, self.bits, 1, 128, n, f, noise, M=0)
        M = int(cnt[0])
        xyzs, dirs, deltas, rays, _ = oc.march_rays_train(o, d, 1.0, self.bits, 1, 128, n, f, noise, M=max(M, 1))
        mx, mdirs, mask = oc.seal_bbox_map_to_origin(xyzs, dirs, self.md, self.tris)
        sig_t, rgb_t = self.teacher.forward(mx, mdirs)
        ws_t, dep_t, img_t = oc.composite_rays_train_forward(sig_t, rgb_t, deltas, rays)
        img_t = img_t + (1 - ws_t)[:, None]
        sig_s, rgb_s = self.student.forward(xyzs, dirs, keep=True)
        ws, dep, comp = oc.composite_rays_train_forward(sig_s, rgb_s, deltas, rays)
        loss, g_img, _ = oc.finetune_loss(comp + (1 - ws)[:, None], dep, img_t, dep_t)
        if not update:
            return float(((comp + (1 - ws)[:, None] - img_t) ** 2).mean(-1).mean()), float(np.abs(dep - dep_t).mean())
        g_ws = -g_img.sum(1)
        gs, gc = oc.composite_rays_train_backward(g_ws, g_img, sig_s, rgb_s, deltas, rays, ws, comp)
        gr = self.student.backward(gs, gc)
        grads = [gr["emb_sigma"], gr["emb_color"], gr["w_s0"], gr["w_s1"], gr["w_c0"], gr["w_c1"], gr["w_c2"]]
        self.t += 1
        b1, b2, lr, eps = 0.9, 0.99, 1e-2, 1e-15
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            m *= b1; m += (1 - b1) * g
            v *= b2; v += (1 - b2) * g * g
            p -= (lr / (1 - b1 ** self.t)) * m / (np.sqrt(v) / np.sqrt(1 - b2 ** self.t) + eps)
        return float(loss), M


def cpu_arm(n_rays, steps, warmup):
    import oracle
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    oracle.set_num_threads(ncpu)        # explicit: an inherited OMP_NUM_THREADS=1 (torchrun) must not decide the baseline
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncpu)
    except Exception:
        pass
    c = CpuDistill(n_rays)
    for i in range(warmup):
        c.step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        _, M = c.step(warmup + i)
    dt = time.perf_counter() - t0
    return dict(value=n_rays * steps / dt, unit="rays/s", cores=oracle.num_threads(), kind="port",
                sample="%d distillation steps of %d rays (%d samples/step) with the oracle's C/numpy restatement (oracle/)" % (steps, n_rays, M),
                ms_per_step=1e3 * dt / steps)


# ------------------------------------------------------------------------------------------ clocks


_SAMPLER_CHILD = r"""
import json, select, sys, time
import pynvml as nv
nv.nvmlInit()
key = sys.argv[1]
try:
    h = nv.nvmlDeviceGetHandleByUUID(key if key.startswith("GPU-") else "GPU-" + key)
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))
mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
print("ready", flush=True)
on, sm, reasons, buf, done = False, [], set(), b"", False
import os
while not done:
    r, _, _ = select.select([0], [], [], 0.02)
    if r:
        chunk = os.read(0, 4096)        # raw reads: a buffered readline could swallow a second command that select never reports
        if not chunk:
            break
        buf += chunk
        while b"\n" in buf:
            line, buf = buf.split(b"\n", 1)
            line = line.strip().decode()
            if line == "start":
                on, sm, reasons = True, [], set()
            elif line == "stop":
                on = False
                print(json.dumps({"sm": sm, "mx": mx, "reasons": sorted(reasons)}), flush=True)
            elif line == "quit":
                done = True
    if on:
        try:
            sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for bit, name in BITS.items():
                if mask & bit:
                    reasons.add(name)
        except Exception:
            pass
"""


class ClockSamplerProcess:
    """The same NVML counters as ClockSampler, polled every 20 ms (six or seven samples in a 20-step timed leg) by a HELPER PROCESS: NVML calls made from the training process
    itself contend with its kernel launches on the driver's locks -- on 8 GPUs rank 0's in-process sampler cost the whole job
    4 % of the leg it ran in (every rank waits for the slowest at the step's barriers).  The helper is started early (NVML
    initialisation takes a few hundred ms) and told over a pipe when the timed region starts and stops."""

    def __init__(self, gpu_index):
        import torch
        self.p = None
        try:
            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            self.p = subprocess.Popen([sys.executable, "-c", _SAMPLER_CHILD, uuid, str(gpu_index)], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True, bufsize=1)
            import select
            r, _, _ = select.select([self.p.stdout], [], [], 20.0)
            if not r or self.p.stdout.readline().strip() != "ready":
                raise RuntimeError("sampler helper did not start")
        except Exception:
            self.close()
            self.p = None

    def ok(self):
        return self.p is not None

    def start(self):
        self.p.stdin.write("start\n")
        self.p.stdin.flush()

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        try:
            self.p.stdin.write("stop\n")
            self.p.stdin.flush()
            import select
            r, _, _ = select.select([self.p.stdout], [], [], 5.0)      # never hang the bench on a dead helper
            if not r:
                raise RuntimeError("sampler helper did not answer")
            d = json.loads(self.p.stdout.readline())
            if d["sm"]:
                out.update(sm_mhz=float(np.median(d["sm"])), sm_max_mhz=d["mx"], reasons=d["reasons"], samples=len(d["sm"]), source="nvml (helper process)")
        except Exception:
            pass
        self.close()
        return out

    def close(self):
        if self.p is not None:
            try:
                self.p.stdin.write("quit\n")
                self.p.stdin.flush()
                self.p.wait(2)
            except Exception:
                try:
                    self.p.kill()
                except Exception:
                    pass
            self.p = None


class ClockSampler:
    """SM clock + clock-event (throttle) reasons of one GPU, sampled DURING the timed region: NVML in a thread (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; an nvidia-smi process per sample is too slow
    for a 0.2 s region on an 8-GPU box), nvidia-smi -lms as the fallback when pynvml is missing."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.p = self.f = self.thread = None
        self.sm, self.reasons, self.mx = [], set(), None
        try:
            import threading
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.stop_flag = False
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)   # the first query of a process is slow (~0.1 s): not inside the timed region
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.mx, reasons=sorted(self.reasons), samples=len(self.sm), source="nvml")
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


# ------------------------------------------------------------------------------------------ GPU arm


def build_world(dev, precision, seed_rank=0):
    import torch
    from seal3d_b200 import synth
    from seal3d_b200.seal import TeacherNetwork, StudentNetwork, SealBBoxMapper
    from seal3d_b200.trainer import DistillTrainer
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    bits, grid = synth.lego_like_occupancy()
    # density_thresh 10 = main_nerf.py:49-50 / main_SealNeRF.py defaults
    t, s = TeacherNetwork(bound=1, density_thresh=10).to(dev), StudentNetwork(bound=1, density_thresh=10).to(dev)
    for net, kind in ((t, "teacher"), (s, "student")):
        fp = synth.field_params(kind)
        net.encoder.embeddings.data.copy_(to(fp["emb_sigma"]))
        net.encoder_color.embeddings.data.copy_(to(fp["emb_color"]))
        for lin, k in ((net.sigma_net[0], "w_s0"), (net.sigma_net[1], "w_s1"), (net.color_net[0], "w_c0"), (net.color_net[1], "w_c1"), (net.color_net[2], "w_c2")):
            lin.weight.data.copy_(to(fp[k]))
        net.density_bitfield.copy_(to(bits))
        net.density_grid.copy_(to(grid))
    md, tris = synth.bbox_edit()
    mapper = SealBBoxMapper(md, tris, device=dev)
    for net in (t, s):
        net.init_mapper(mapper)
        net.hack_bitfield()
    t.eval()
    for p in t.parameters():
        p.requires_grad_(False)
    return t, s


def roofline_grid_encode(dev, precision):
    """dominant kernel of the step: the hash-grid gather.  Timed alone on B = 2^22 random points (inputs 48 MB +
    outputs > L2 between launches), CUDA events on the launching stream; algorithmic bytes per point from
    SURVEY.md 8(d): 12 + 16*8*2*s + 16*2*s (s = bytes per table element)."""
    import torch
    from seal3d_b200 import _lib, synth
    offsets, pls = synth.grid_offsets()
    B = 1 << 22
    dt = torch.float16 if precision == "fp16" else torch.float32
    s_el = 2 if precision == "fp16" else 4
    g = torch.Generator(device=dev).manual_seed(8)
    x = torch.rand(B, 3, device=dev, generator=g)
    emb = (torch.rand(int(offsets[-1]), 2, device=dev, generator=g) * 2e-4 - 1e-4).to(dt)
    out = torch.empty(16, B, 2, device=dev, dtype=dt)
    off = torch.from_numpy(offsets).to(dev)
    S = float(np.log2(pls))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("s3d_grid_encode_forward", x, emb, off, out, B, 3, 2, 16, S, 16, None, 0, 0, 0, 0 if s_el == 4 else 1)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(e0.elapsed_time(e1) * 1e-3)
    per_pt = 12 + 16 * 8 * 2 * s_el + 16 * 2 * s_el
    t = float(np.mean(times))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = B * per_pt / t / 1e9
    kname = "k_grid_forward<%s,3,2,all-levels>" % ("half" if s_el == 2 else "float")
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if kname in tj:
            traffic = tj[kname]["dram_bytes_per_sample"] * B
    except Exception:
        pass
    return dict(kernel=kname, bound="hbm", achieved=ach, peak=peak,
                unit="GB/s", frac=ach / peak, traffic=traffic, peak_source="MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6.65 TB/s",
                points=B, bytes_per_point=per_pt, launch_ms=t * 1e3)


# algorithmic bytes per sample of the step's kernels (DESIGN.md "Kernels"): what one launch must move, counted
# the way 8(d) counts the fp16 configuration -- a table / gradient-table entry is 2 bytes per value, a read-modify-write once
ALGO_BYTES = {
    # xyz 12 + 16 levels x 8 corners x 8 B (both fp16 tables of one model, interleaved) + 128 B feature row
    "s3d_ngp_encode": ("k_ngp_encode", 12 + 16 * 8 * 8 + 128),
    # teacher + student on one sample: xyz 12 + 16 x 8 corners x 16 B (four fp16 tables) + two 128 B feature rows
    "s3d_ngp_encode_pair": ("k_ngp_encode_pair", 12 + 16 * 8 * 16 + 256),
    # xyz 12 + 128 B dfeats row + 16 x 8 corners x 8 B (8d's 2 x 512 B scatter term: two fp16-valued tables).  The kernel
    # itself accumulates into fp32 entries (16 B per corner = 2188 B per sample): reported separately as `achieved_fp32_entries`
    "s3d_ngp_scatter": ("k_ngp_scatter", 12 + 128 + 16 * 8 * 8),
    # feature row 128 + dirs 12 + sigma 4 + rgb 12
    "s3d_ngp_mlp_forward": ("k_ngp_mlp_fwd_ts", 128 + 12 + 16),
    # feature row 128 + dirs 12 + g_sigma 4 + g_rgb 12 + dfeats 128
    "s3d_ngp_mlp_backward": ("k_ngp_mlp_bwd", 128 + 12 + 16 + 128),
    "s3d_grid_encode_forward": ("k_grid_forward", 12 + 16 * 8 * 4 + 64),       # one fp16 table
    "s3d_grid_encode_backward": ("k_grid_backward", 12 + 16 * 8 * 4 + 64),
}
ALGO_BYTES_ALT = {"s3d_ngp_scatter": ("achieved_fp32_entries", 12 + 128 + 16 * 8 * 16)}
# measured on one B200 (scripts/r2/red_micro.cu, profiles/r2_red_rate_micro_run36.log), in 1e9 reductions per second, into a 98 MB table:
RED_RATE_RANDOM_G = 152.0   # isolated random entries -- the same for RED.32 / .64 / .128 / packed fp16: width does not matter
RED_RATE_PAIRED_G = 213.0   # two (or four) adjacent 16-byte entries of one sector: the pattern of a grid cell's x-corner pairs


def reduction_rate_block(red, samples_per_step, launch_ms):
    """red = (reductions, samples, reductions on hashed levels) counted by s3d_ngp_scatter_count on one batch -> the scatter's
    reduction-rate roofline for a launch of samples_per_step samples that took launch_ms"""
    n_red = red[0] * samples_per_step / red[1]
    floor_ms = n_red / (RED_RATE_PAIRED_G * 1e9) * 1e3
    return {
        "reductions_per_sample": red[0] / red[1], "hashed_level_reductions_per_sample": red[2] / red[1], "reductions_per_launch": n_red,
        "achieved": n_red / (launch_ms * 1e-3) / 1e9, "peak": RED_RATE_PAIRED_G, "unit": "G reductions/s", "frac": floor_ms / launch_ms,
        "floor_ms": floor_ms, "random_address_rate": RED_RATE_RANDOM_G,
        "note": "the gradient table stays in L2, so this kernel is bound by the chip's global-reduction RATE, not by bytes: isolated random "
                "reductions retire at 152 G/s whatever their width (RED.32 / .64 / .128 / f16x2), same-sector pairs -- the two x-corners of a cell, "
                "which is how this kernel's reductions come -- at 213 G/s (scripts/r2/red_micro.cu, profiles/r2_red_rate_micro_run36.log); "
                "frac = (reductions / 213 G/s) / launch time"}


def whole_step_roofline(n_rays, samples_per_step, rays_per_s_per_gpu):
    """the north star's second figure: the step as a fraction of the HBM-bandwidth roofline, with the algorithmic bytes
    per ray of SURVEY.md 8(d) (fp16 tables): 68 B per ray + per sample 32 (sample buffer) + 2 x 512 teacher gathers + 2 x 512
    student gathers + 2 x 512 student scatter + 16 (sigma, rgb) = 3120 B"""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    spr = samples_per_step / n_rays
    bytes_per_ray = 68.0 + 3120.0 * spr
    bound = peak * 1e9 / bytes_per_ray
    return {"bytes_per_ray": bytes_per_ray, "samples_per_ray": spr, "hbm_bound_rays_per_s_per_gpu": bound,
            "frac": rays_per_s_per_gpu / bound, "peak_GBps": peak,
            "note": "algorithmic bytes per ray of SURVEY.md 8(d) / measured HBM copy bandwidth; tables are L2-resident, so this is not DRAM traffic"}


def step_roofline(breakdown, samples_per_step, step_ms):
    """roofline of the step's dominant kernel: algorithmic bytes per launch / mean launch duration (CUDA events on the
    launching stream, measured live above), against the measured HBM copy bandwidth"""
    top = max(breakdown.items(), key=lambda kv: kv[1]["ms_per_step"])
    name, st = top
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    out = {"kernel": ALGO_BYTES.get(name, (name, None))[0], "bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": None,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
           "launch_ms": st["ms_per_call"], "launches_per_step": st["calls_per_step"], "share_of_step": st["ms_per_step"] / step_ms}
    if name in ALGO_BYTES:
        per = ALGO_BYTES[name][1]
        ach = samples_per_step * per / (st["ms_per_call"] * 1e-3) / 1e9
        out.update(achieved=ach, frac=ach / peak, bytes_per_sample=per, samples_per_launch=samples_per_step)
        if name in ALGO_BYTES_ALT:
            k, alt = ALGO_BYTES_ALT[name]
            out[k] = {"bytes_per_sample": alt, "achieved": samples_per_step * alt / (st["ms_per_call"] * 1e-3) / 1e9,
                      "frac": samples_per_step * alt / (st["ms_per_call"] * 1e-3) / 1e9 / peak}
        try:   # DRAM traffic of this kernel from the committed ncu --set full capture, scaled by samples to this launch
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            t = tj.get(ALGO_BYTES[name][0])
            if t:
                out.update(traffic=t["dram_bytes_per_sample"] * samples_per_step, traffic_source=tj.get("_source"),
                           algorithmic_bytes=per * samples_per_step)
        except Exception:
            pass
    else:
        out.update(achieved=None, frac=None)
    return out


class _Timer:
    """start/stop pair: CUDA events on the current stream of a CUDA device, the host clock otherwise (the CPU / gloo test of
    this file's control flow, tests/test_host_cpu.py)"""

    def __init__(self, dev):
        import torch
        self.cuda = dev.type == "cuda"
        if self.cuda:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def start(self):
        if self.cuda:
            self.e0.record()
        else:
            self.t0 = time.perf_counter()

    def stop(self):
        if self.cuda:
            self.e1.record()
        else:
            self.t1 = time.perf_counter()

    def ms(self):      # after a synchronize
        return self.e0.elapsed_time(self.e1) if self.cuda else 1e3 * (self.t1 - self.t0)


def timed_legs(tr, resident, host, args, rank, world, dev, pipelined, profile_hook=None, helper=None):
    """Everything of the GPU arm that happens after the world is built: warm-up, leg 1 (resident inputs), leg 2 (end to end),
    the per-kernel breakdown.  EVERY rank executes EVERY trainer step in here -- each step contains the gradient all-reduce,
    so a step taken by a subset of the ranks is an unmatched collective (the round-1 driver runs at N = 2, 4, 8 hung on
    exactly that).  `tr` needs distill_step(o, d, perturb=, force_all_rays=, prefetch=) -> loss tensor, refresh_occupancy(),
    student.mean_count and student.step_counter; `profile_hook` = (begin(), end() -> [(name, ms)]) around the breakdown steps;
    `helper` = rank 0's ClockSamplerProcess (started by the caller well before the timed region)."""
    import torch
    import torch.distributed as dist
    from seal3d_b200 import parallel
    pool = len(resident)
    n = args.rays

    def barrier():
        if world > 1:
            dist.barrier()
        if dev.type == "cuda":
            torch.cuda.synchronize()

    # warm-up (first steps run in exact mode and set mean_count)
    for i in range(args.warmup):
        o, d = resident[i % pool]
        tr.distill_step(o, d, perturb=True, force_all_rays=(i < 2))
    if tr.student.mean_count <= 0:
        tr.refresh_occupancy()
    if pipelined:   # the side stream, its allocator pool and the pinned-copy path are created on first use: outside the timed legs
        for src in (resident, host):
            for i in range(3):
                tr.distill_step(*src[i % pool], perturb=True, prefetch=src[(i + 1) % pool] if i < 2 else None)
    if world > 1:   # NCCL picks / tunes its all-reduce channels over the first collectives of this size: keep that out of the timed legs too
        for i in range(4):
            tr.distill_step(*resident[i % pool], perturb=True)
    samples_per_step = float(tr.student.step_counter[:, 0].float().max().item())

    # -- leg 1: resident inputs ---------------------------------------------------------------
    barrier()
    # rank 0 samples its GPU's clocks during the timed region (the line reports that GPU).  The other ranks do not: eight
    # processes polling NVML at once contend on the driver's global lock and slow every rank's kernel launches -- measured on
    # 8 GPUs as 7.78 ms/step in this leg against 7.18 in the next one, which has no sampler
    clocks = None
    if dev.type == "cuda" and rank == 0:
        if helper is not None and helper.ok():
            try:
                helper.start()
                clocks = helper
            except Exception:       # the helper died after it reported ready: sample in-process instead
                clocks = None
        if clocks is None:
            clocks = ClockSampler(dev.index)
    from seal3d_b200 import _lib
    launches0 = _lib.LAUNCHES
    tm = _Timer(dev)
    tm.start()
    for i in range(args.steps):
        o, d = resident[i % pool]
        if pipelined:   # the next batch is marched on a side stream under this step's field kernels (fused.py: _prefetch)
            tr.distill_step(o, d, perturb=True, prefetch=resident[(i + 1) % pool] if i + 1 < args.steps else None)
        else:
            tr.distill_step(o, d, perturb=True)
    tm.stop()
    barrier()
    ms = parallel.max_over_ranks(tm.ms(), dev)
    clk = clocks.stop() if clocks is not None else None
    gpu_launches = int(_lib.LAUNCHES - launches0)     # kernels launched through the C-ABI inside the timed region of leg 1

    # -- leg 2: end to end through the public call, host buffers in, loss out, every step -------
    barrier()
    tm.start()
    last = None
    for i in range(args.steps):
        ho, hd = host[i % pool]
        if pipelined:   # pinned host buffers go in as they are: this step's copy + march were issued during the previous step
            last = tr.distill_step(ho, hd, perturb=True, prefetch=host[(i + 1) % pool] if i + 1 < args.steps else None).cpu()
        else:
            o, d = ho.to(dev, non_blocking=True), hd.to(dev, non_blocking=True)
            last = tr.distill_step(o, d, perturb=True).cpu()
    tm.stop()
    barrier()
    ms_e2e = parallel.max_over_ranks(tm.ms(), dev)

    # -- per-kernel breakdown of one step (events around every C-ABI launch; separate from the timed legs) -------
    breakdown = None
    if not args.no_roofline:
        barrier()
        nprof = 3
        if profile_hook:
            profile_hook[0]()
        for i in range(nprof):
            o, d = resident[i % pool]
            tr.distill_step(o, d, perturb=True)
        barrier()
        agg = {}
        for name, ms_k in (profile_hook[1]() if profile_hook else []):
            agg.setdefault(name, [0.0, 0])
            agg[name][0] += ms_k
            agg[name][1] += 1
        breakdown = {k: {"ms_per_step": v[0] / nprof, "calls_per_step": v[1] / nprof, "ms_per_call": v[0] / v[1]} for k, v in agg.items()}
    barrier()
    return dict(ms=ms, ms_e2e=ms_e2e, clocks=clk, gpu_launches=gpu_launches, samples_per_step=samples_per_step, last=last, breakdown=breakdown)


def gpu_arm(args):
    import torch
    from seal3d_b200 import synth, parallel, _lib
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from seal3d_b200.trainer import DistillTrainer
    from seal3d_b200.fused import FusedDistillTrainer
    teacher, student = build_world(dev, args.precision)
    if args.engine == "fused":
        tr = FusedDistillTrainer(student, teacher, lr=1e-2, world_size=world, update_interval=16)
    else:
        tr = DistillTrainer(student, teacher, lr=1e-2, precision=args.precision, loss_scale=(128.0 if args.precision == "fp16" else 1.0),
                            world_size=world, update_interval=16)
    n = args.rays
    pool = 4
    host = []
    for b in range(pool):
        o, d = synth.rays_for_step(1000 * rank + b, n)
        host.append((torch.from_numpy(o).pin_memory(), torch.from_numpy(d).pin_memory()))
    resident = [(o.to(dev), d.to(dev)) for o, d in host]
    pipelined = args.engine == "fused" and not args.no_prefetch

    def prof_begin():
        _lib.PROFILE = []

    def prof_end():
        rows = [(name, a.elapsed_time(b)) for name, a, b in _lib.PROFILE]
        _lib.PROFILE = None
        return rows

    helper = ClockSamplerProcess(dev.index) if rank == 0 else None
    try:
        r = timed_legs(tr, resident, host, args, rank, world, dev, pipelined, profile_hook=(prof_begin, prof_end), helper=helper)
    finally:
        if helper is not None:
            helper.close()
    if rank != 0:
        return None
    red = None
    if args.engine == "fused" and not args.no_roofline:
        # the scatter's real unit of work: global reductions per launch (the same run detection as the kernel, counted by
        # s3d_ngp_scatter_count on one batch marched like the step marches it; no collective, rank 0 only)
        xyzs = tr._march(resident[0][0], resident[0][1], True, False)[0]
        cnt = torch.zeros(2, dtype=torch.int64, device=dev)
        S = tr.S
        _lib.call("s3d_ngp_scatter_count", xyzs, xyzs.shape[0], S.bound, S.offsets, S.L, S.S, S.H, cnt)
        red = (int(cnt[0].item()), int(xyzs.shape[0]), int(cnt[1].item()))
    ms, ms_e2e, samples_per_step, last, breakdown = r["ms"], r["ms_e2e"], r["samples_per_step"], r["last"], r["breakdown"]
    rays_total = n * world * args.steps
    line = {
        "metric": "training rays/sec (Lego 800x800, NGP L16/F2, distill on)", "value": rays_total / (ms * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": n, "global_rays_per_step": n * world, "samples_per_step_per_gpu": samples_per_step,
                   "samples_per_ray": samples_per_step / n, "parallelism": ("dp%d (ray shards; gradients reduced + Adam + tables broadcast by one kernel over NVLink peer memory, no collective call in the step)" % world)
                   if getattr(tr, "peer", None) is not None else ("dp%d (ray shards, one grad all-reduce/step)" % world),
                   "schedule": "fused teacher+student on shared samples; occupancy refresh every 16 steps" + ("; next batch marched on a side stream under the current step" if pipelined else ""),
                   "l2_note": "each step streams > 126 MB (samples + 4 tables + arena) so successive steps do not reuse L2 contents"},
        "e2e": {"value": rays_total / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 24 * n, "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": r["gpu_launches"],
        "clocks": r["clocks"], "loss_last": [float(v) for v in last.numpy()] if last is not None else None,
    }
    line["config"]["engine"] = args.engine
    line["step_hbm_roofline"] = whole_step_roofline(n, samples_per_step, n * args.steps / (ms * 1e-3))
    if breakdown:
        line["kernel_breakdown_ms_per_step"] = {k: round(v["ms_per_step"], 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms_per_step"])}
        line["roofline"] = step_roofline(breakdown, samples_per_step, ms / args.steps)
        if red and line["roofline"].get("kernel") == "k_ngp_scatter":
            line["roofline"]["reduction_rate"] = reduction_rate_block(red, samples_per_step, line["roofline"]["launch_ms"])
    return line


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        # the reference has no CPU path of its own (every op requires CUDA tensors); the CPU arm is the
        # oracle's restatement of the same step, on all host threads OpenMP / BLAS give it
        r = cpu_arm(args.cpu_rays, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": "training rays/sec (Lego 800x800, NGP L16/F2, distill on)", "value": r["value"], "unit": "rays/s",
                "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": max(1, min(args.warmup, 1)), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_step": args.cpu_rays}, "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    import torch
    try:
        line = gpu_arm(args)
        if line is not None:
            finish_line(args, line)
    finally:
        # every rank leaves through here: the ranks that have nothing to print wait for rank 0's single-GPU extras, then
        # the process group is torn down on all of them (no rank exits with a communicator another rank still uses)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            try:
                torch.distributed.barrier()
            finally:
                torch.distributed.destroy_process_group()


def reference_gpu_block():
    """the committed same-box measurement of the UNMODIFIED reference kernels (scripts/ref_ab.py on a B200: per-op times and the
    composed reference `-O` step next to this repo's).  bench.py itself never loads oracle/_ref on the GPU arm."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r2_ref_vs_ours.json")))
    except Exception:
        return None
    st = j["step"]
    return {"source": "profiles/r2_ref_vs_ours.json (scripts/ref_ab.py, committed measurement, not re-run here)", "device": j.get("device"),
            "step_rays": st["rays_per_step"], "reference_ms_per_step": st["reference_ms_per_step"], "reference_rays_per_s": st["reference_rays_per_s"],
            "ours_ms_per_step_same_run": st["ours_ms_per_step"], "speedup_same_run": st["speedup"],
            "ops": {r["op"]: {"reference_ms": round(r["reference_ms"], 4), "ours_ms": round(r["ours_ms"], 4)} for r in j["ops"]}}


def finish_line(args, line):
    import torch
    rg = reference_gpu_block()
    if rg:
        line["reference_gpu"] = rg
    if not args.no_roofline:
        d0 = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        line["grid_encode_forward_roofline"] = {p: roofline_grid_encode(d0, p) for p in ("fp32", "fp16")}
        if "roofline" not in line:
            line["roofline"] = line["grid_encode_forward_roofline"][args.precision]
    if not args.no_cpu_baseline and line["n_gpus"] == 1:
        line["cpu_baseline"] = cpu_arm(args.cpu_rays, 2, 1)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
