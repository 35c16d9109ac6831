"""CPU oracle for the Seal-3D hot path -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrappers over ``liboracle.so`` (``seal_oracle.c``, a CPU restatement of the
reference kernels, each function citing the reference file:line it follows) plus the few
pieces that are plain numpy (the NGP field of nerf/network.py composed from the C ops, the
distillation losses, Adam).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package never does: it
fails loudly when its CUDA library is missing.

Pinning status: see the header of seal_oracle.c and DESIGN.md ("Oracle").
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "seal_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        base = [cc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", so, src, "-lm"]
        try:
            subprocess.check_call(base[:1] + ["-fopenmp"] + base[1:])
        except (subprocess.CalledProcessError, OSError):
            subprocess.check_call(base)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    """OpenMP thread count of the C restatement (bench.py's CPU arm sets it explicitly: torchrun exports OMP_NUM_THREADS=1)"""
    lib().orc_set_num_threads(C.c_int(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


u32, f32, i64, cint = C.c_uint32, C.c_float, C.c_int64, C.c_int

# ---------------------------------------------------------------- raymarching


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_p(rays_o), _p(rays_d), f32(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3D(_p(coords), u32(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), u32(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), u32(N), f32(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises=None, dt_gamma=0.0, max_steps=1024,
                     M=None):
    """Returns xyzs[M,3], dirs[M,3], deltas[M,2], rays[N,3] (id, offset, count), counter[2].
    Deterministic ray-major slot order (see seal_oracle.c)."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    if noises is None:
        noises = np.zeros(N, np.float32)
    if M is None:
        M = N * max_steps
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    lib().orc_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32(bound), f32(dt_gamma), u32(max_steps), u32(N),
                               u32(Cc), u32(H), u32(M), _p(_f32(nears)), _p(_f32(fars)), _p(xyzs), _p(dirs), _p(deltas),
                               _p(rays), _p(counter), _p(_f32(noises)))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), u32(M), u32(N), f32(T_thresh),
                                           _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_rays_train_backward(_p(_f32(grad_ws)), _p(_f32(grad_image)), _p(sigmas), _p(rgbs), _p(deltas),
                                            _p(rays), _p(_f32(ws)), _p(_f32(image)), u32(M), u32(N), f32(T_thresh),
                                            _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises=None,
               dt_gamma=0.0, max_steps=1024, align=-1):
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    if noises is None:
        noises = np.zeros(n_alive, np.float32)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _p(_i32(rays_alive)), _p(_f32(rays_t)), _p(_f32(rays_o)),
                         _p(_f32(rays_d)), f32(bound), f32(dt_gamma), u32(max_steps), u32(Cc), u32(H), _p(bitfield),
                         _p(_f32(nears)), _p(_f32(fars)), _p(xyzs), _p(dirs), _p(deltas), _p(_f32(noises)))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In place on rays_alive (int32), rays_t, weights_sum, depth, image (float32, contiguous)."""
    for a, t in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32),
                 (image, np.float32)):
        assert a.dtype == t and a.flags.c_contiguous
    lib().orc_composite_rays(u32(n_alive), u32(n_step), f32(T_thresh), _p(rays_alive), _p(rays_t), _p(_f32(sigmas)),
                             _p(_f32(rgbs)), _p(_f32(deltas)), _p(weights_sum), _p(depth), _p(image))


# ---------------------------------------------------------------------------- density-grid refresh
def pcg_hash(v):
    """the counter-based hash the device kernels draw from (csrc/density.cu: pcg): the reference uses torch's unseeded global
    generator (nerf/renderer.py:479, 492, 496), whose stream cannot be reproduced, so parity of update_extra_state is defined
    on an INJECTED stream -- this function regenerates the device's stream so a test can inject it here"""
    v = np.asarray(v, dtype=np.uint64)
    v = (v * 747796405 + 2891336453) & 0xFFFFFFFF
    w = (((v >> ((v >> 28) + 4)) ^ v) * 277803737) & 0xFFFFFFFF
    return (((w >> 22) ^ w) & 0xFFFFFFFF).astype(np.uint32)


def density_draws(seed, n_uniform, n_occ, H):
    """-> dict(coords int32 [n_uniform,3] in [0,H), occ_u32 uint32 [n_occ], jitter float32 [n_uniform+n_occ,3] in [0,1))
    for one cascade, from seed (the per-cascade seed, (seed + 7919 * cas) & 0xffffffff in renderer.update_extra_state)"""
    seed = np.uint32(seed & 0xFFFFFFFF)
    draw = lambda sd, ctr: pcg_hash(np.uint32(sd) ^ pcg_hash(ctr))
    c = draw(seed, np.arange(n_uniform * 3, dtype=np.uint64))
    coords = ((c.astype(np.uint64) * H) >> 32).astype(np.int32).reshape(n_uniform, 3)
    occ = draw(int(seed) ^ 0x9e3779b9, np.arange(n_occ, dtype=np.uint64))
    n = n_uniform + n_occ
    j = draw(int(seed) ^ 0x85ebca6b, np.arange(n * 3, dtype=np.uint64))
    jitter = ((j >> 8).astype(np.float32) * np.float32(1.0 / 16777216.0)).reshape(n, 3)
    return dict(coords=coords, occ_u32=occ, jitter=jitter)


def density_cells_and_xyz(density_grid_cas, H, bound_cas, draws=None, jitter=None):
    """nerf/renderer.py:457-479 (full sweep, draws=None: every cell once, jitter [H^3,3] injected) and :487-509 (partial
    update: uniform coords + occupied cells indexed by injected draws) for ONE cascade -> (cells int64 [n] morton, xyz [n,3])"""
    f32 = np.float32
    if draws is None:
        cells = np.arange(H ** 3, dtype=np.int64)
        coords = morton3D_invert(cells.astype(np.int32))
    else:
        idx_u = morton3D(draws["coords"]).astype(np.int64)                          # :492-493
        occ = np.nonzero(density_grid_cas > 0)[0]                                  # :495 (ascending)
        if occ.shape[0] > 0:
            k = ((draws["occ_u32"].astype(np.uint64) * np.uint64(occ.shape[0])) >> np.uint64(32)).astype(np.int64)   # :496 randint(0, Nz)
            idx_o = occ[k]                                                          # :497
        else:      # the reference raises here (randint(0, 0)); the device repeats the uniform half
            idx_o = idx_u[np.arange(draws["occ_u32"].shape[0]) % max(idx_u.shape[0], 1)]
        cells = np.concatenate([idx_u, idx_o])                                      # :500
        coords = np.concatenate([draws["coords"], morton3D_invert(idx_o.astype(np.int32))])   # :498, :501
        jitter = draws["jitter"]
    xyzs = f32(2) * coords.astype(f32) / f32(H - 1) - f32(1)                        # :470 / :503
    hgs = f32(bound_cas) / f32(H)                                                   # :475 / :505 (python floats; exact for powers of two)
    cas = xyzs * (f32(bound_cas) - hgs)                                             # :477 / :507
    cas = cas + (jitter.astype(f32) * f32(2) - f32(1)) * hgs                        # :479 / :509
    return cells, cas.astype(f32)


def density_grid_update(density_grid, tmp_grid, decay=0.95, density_thresh=0.01):
    """nerf/renderer.py:521-530: EMA-max where both valid, mean of clamp(grid, 0), threshold, packbits
    -> (grid float32, mean_density float, thresh float, bitfield uint8)"""
    g = density_grid.astype(np.float32).copy()
    valid = (g >= 0) & (tmp_grid >= 0)
    g[valid] = np.maximum(g[valid] * np.float32(decay), tmp_grid[valid])
    mean = float(np.float32(np.clip(g, 0, None).astype(np.float64).mean()))
    thresh = min(mean, float(density_thresh))
    return g, mean, thresh, packbits(g.reshape(-1), thresh)


def update_extra_state(density_grid, density_fn, H, bound, density_scale, density_thresh, iter_density, per_cascade, step_counter,
                       local_step, decay=0.95, duplicates="max"):
    """nerf/renderer.py:445-538 for all cascades.  per_cascade[c] = dict(jitter=...) for a full sweep (iter_density < 16) or the
    dict of density_draws for a partial update.  density_fn(xyz [n,3] float32) -> sigma [n] (WITHOUT density_scale, like
    self.density(...)['sigma']).  A cell drawn more than once keeps the largest of its values (the reference's index_put
    keeps an arbitrary one of them; the device takes the max so the result is deterministic); duplicates="last" = the
    serial last-write-wins of CPU torch, which is what the committed run of the reference method did (cpu_extra_state.npz).
    -> dict(grid, mean_density, thresh, bitfield, mean_count (None when local_step == 0), cells, xyz)"""
    C = density_grid.shape[0]
    tmp = -np.ones_like(density_grid, dtype=np.float32)
    all_cells, all_xyz = [], []
    for cas in range(C):
        bound_cas = min(2 ** cas, bound)
        if iter_density < 16:
            cells, xyz = density_cells_and_xyz(density_grid[cas], H, bound_cas, None, per_cascade[cas]["jitter"])
        else:
            cells, xyz = density_cells_and_xyz(density_grid[cas], H, bound_cas, per_cascade[cas])
        sig = np.asarray(density_fn(xyz), dtype=np.float32).reshape(-1) * np.float32(density_scale)   # :481-482 / :511-512
        if duplicates == "max":
            np.maximum.at(tmp[cas], cells, sig)                                                       # :483 / :513
        else:
            tmp[cas][cells] = sig
        all_cells.append(cells); all_xyz.append(xyz)
    g, mean, thresh, bits = density_grid_update(density_grid, tmp, decay, density_thresh)
    total = min(16, local_step)                                                                       # :533
    mean_count = int(step_counter[:total, 0].sum() / total) if total > 0 else None                    # :534-535
    return dict(grid=g, mean_density=mean, thresh=thresh, bitfield=bits, mean_count=mean_count, cells=all_cells, xyz=all_xyz, tmp=tmp)


def get_rays(poses, intrinsics, H, W, inds=None):
    """nerf/utils.py:53-140: poses [B,4,4], intrinsics (fx, fy, cx, cy), inds int64 [N] / [B,N] or None (all pixels)
    -> rays_o, rays_d [B,N,3]"""
    poses = _f32(poses).reshape(-1, 4, 4)
    B = poses.shape[0]
    fx, fy, cx, cy = [float(v) for v in intrinsics]
    if inds is not None:
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        rows = 1 if inds.ndim == 1 else inds.shape[0]
        N = inds.shape[-1]
    else:
        rows, N = 1, H * W
    ro, rd = np.empty((B, N, 3), np.float32), np.empty((B, N, 3), np.float32)
    lib().orc_get_rays(_p(poses), u32(B), f32(fx), f32(fy), f32(cx), f32(cy), u32(W), _p(inds), u32(rows), u32(N), _p(ro), _p(rd))
    return ro, rd


def mark_untrained_count(poses, intrinsic, cascade=1, H=128, bound=1.0):
    """nerf/renderer.py:379-443: per-cell camera count [cascade, H^3] (morton order); the grid gets -1 where it is 0"""
    poses = _f32(poses).reshape(-1, 4, 4)
    fx, fy, cx, cy = [float(v) for v in intrinsic]
    count = np.empty((cascade, H ** 3), np.int32)
    lib().orc_mark_untrained_count(_p(poses), u32(poses.shape[0]), f32(cx / fx), f32(cy / fy), u32(cascade), u32(H), f32(bound), _p(count))
    return count


# ---------------------------------------------------------------- gridencoder


class level_scales:
    """with oracle.level_scales(float32[L]): ...  -- use the given per-level scales instead of libm's exp2f (see
    seal_oracle.c: the one value of the path that differs between CUDA's and glibc's math library)"""

    def __init__(self, scales):
        self.s = None if scales is None else np.ascontiguousarray(scales, dtype=np.float32)

    def __enter__(self):
        lib().orc_set_level_scales(_p(self.s))
        return self

    def __exit__(self, *a):
        lib().orc_set_level_scales(None)



def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    """Level offsets exactly as gridencoder/grid.py:100-127 computes them.  Returns (offsets int32[L+1],
    per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


def grid_encode_forward(inputs, emb, offsets, per_level_scale, H, calc_grad_inputs=False, gridtype=0,
                        align_corners=False, interp=0, half_accum=False):
    """Returns outputs [L,B,C] (the reference's kernel layout) and dy_dx [B,L*D*C] or None."""
    inputs, emb, offsets = _f32(inputs), _f32(emb), _i32(offsets)
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    out = np.empty((L, B, Cc), np.float32)
    dy_dx = np.empty((B, L * D * Cc), np.float32) if calc_grad_inputs else None
    lib().orc_grid_encode_forward(_p(inputs), _p(emb), _p(offsets), _p(out), u32(B), u32(D), u32(Cc), u32(L), f32(S),
                                  u32(H), _p(dy_dx), u32(gridtype), cint(int(align_corners)), u32(interp),
                                  cint(int(half_accum)))
    return out, dy_dx


def grid_encode_backward(grad, inputs, emb_shape, offsets, per_level_scale, H, dy_dx=None, gridtype=0,
                         align_corners=False, interp=0):
    """grad [L,B,C].  Returns grad_embeddings [sO,C] (+ grad_inputs [B,D] when dy_dx is given)."""
    grad, inputs, offsets = _f32(grad), _f32(inputs), _i32(offsets)
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, emb_shape[1]
    S = np.float32(np.log2(per_level_scale))
    ge = np.zeros(emb_shape, np.float32)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    lib().orc_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(ge), u32(B), u32(D), u32(Cc), u32(L), f32(S),
                                   u32(H), _p(None if dy_dx is None else _f32(dy_dx)), _p(gi), u32(gridtype),
                                   cint(int(align_corners)), u32(interp))
    return (ge, gi) if dy_dx is not None else ge


def grad_total_variation(inputs, emb, grad, offsets, weight, per_level_scale, H, gridtype=0, align_corners=False):
    inputs, emb, offsets = _f32(inputs), _f32(emb), _i32(offsets)
    assert grad.dtype == np.float32 and grad.flags.c_contiguous
    B, D = inputs.shape
    L, Cc = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    lib().orc_grad_total_variation(_p(inputs), _p(emb), _p(grad), _p(offsets), f32(weight), u32(B), u32(D), u32(Cc),
                                   u32(L), f32(S), u32(H), u32(gridtype), cint(int(align_corners)))


def round_to_half(a):
    a = _f32(a)
    out = np.empty_like(a)
    lib().orc_round_to_half(_p(a), _p(out), i64(a.size))
    return out


# ---------------------------------------------------------------- shencoder / freqencoder


def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    inputs = _f32(inputs)
    B, D = inputs.shape
    out = np.empty((B, degree * degree), np.float32)
    dy_dx = np.empty((B, D * degree * degree), np.float32) if calc_grad_inputs else None
    lib().orc_sh_encode_forward(_p(inputs), _p(out), u32(B), u32(D), u32(degree), _p(dy_dx))
    return out, dy_dx


def sh_encode_backward(grad, inputs, degree, dy_dx):
    grad, inputs, dy_dx = _f32(grad), _f32(inputs), _f32(dy_dx)
    B, D = inputs.shape
    gi = np.zeros((B, D), np.float32)
    lib().orc_sh_encode_backward(_p(grad), _p(inputs), u32(B), u32(D), u32(degree), _p(dy_dx), _p(gi))
    return gi


def freq_encode_forward(inputs, degree):
    inputs = _f32(inputs)
    B, D = inputs.shape
    Cc = D + D * 2 * degree
    out = np.empty((B, Cc), np.float32)
    lib().orc_freq_encode_forward(_p(inputs), u32(B), u32(D), u32(degree), u32(Cc), _p(out))
    return out


def freq_encode_backward(grad, outputs, D, degree):
    grad, outputs = _f32(grad), _f32(outputs)
    B, Cc = grad.shape
    gi = np.empty((B, D), np.float32)
    lib().orc_freq_encode_backward(_p(grad), _p(outputs), u32(B), u32(D), u32(degree), u32(Cc), _p(gi))
    return gi


# ---------------------------------------------------------------- ffmlp


def ffmlp_forward(inputs, weights, in_dim, out_dim, hidden, num_layers, act=0, out_act=6, round_half_act=False,
                  want_buffer=True):
    inputs, weights = _f32(inputs), _f32(weights)
    B = inputs.shape[0]
    fb = np.empty((num_layers, B, hidden), np.float32) if want_buffer else None
    out = np.empty((B, out_dim), np.float32)
    lib().orc_ffmlp_forward(_p(inputs), _p(weights), u32(B), u32(in_dim), u32(out_dim), u32(hidden), u32(num_layers),
                            u32(act), u32(out_act), _p(fb), _p(out), cint(int(round_half_act)))
    return out, fb


def ffmlp_backward(grad, inputs, weights, forward_buffer, in_dim, out_dim, hidden, num_layers, act=0,
                   calc_grad_inputs=True):
    grad, inputs, weights, fb = _f32(grad), _f32(inputs), _f32(weights), _f32(forward_buffer)
    B = inputs.shape[0]
    bb = np.empty((num_layers, B, hidden), np.float32)
    gi = np.empty((B, in_dim), np.float32) if calc_grad_inputs else None
    gw = np.empty_like(weights)
    lib().orc_ffmlp_backward(_p(grad), _p(inputs), _p(weights), _p(fb), u32(B), u32(in_dim), u32(out_dim), u32(hidden),
                             u32(num_layers), u32(act), _p(bb), _p(gi), _p(gw))
    return gw, gi, bb


# ---------------------------------------------------------------- Seal proxy mapping


def seal_map_mask(points, bounds, tris, test_dir=None):
    points, bounds, tris = _f32(points), _f32(bounds).reshape(-1, 2, 3), _f32(tris).reshape(-1, 3, 3)
    P = points.shape[0]
    mask = np.zeros(P, np.uint8)
    lib().orc_seal_map_mask(_p(points), i64(P), _p(bounds), u32(bounds.shape[0]), _p(tris), u32(tris.shape[0]),
                            _p(None if test_dir is None else _f32(test_dir)), _p(mask))
    return mask.astype(bool)


def seal_bbox_map_to_origin(points, dirs, map_data, tris, test_dir=None):
    """map_data: dict with transform[4,4] (inverse), rotation[3,3] (inverse), scale[3] (=1/scale), center[3],
    map_bound [2,3]|[nb,2,3], optional empty_bound[2,3] + map_source[3] (seal_utils.py:222-236)."""
    points, dirs = _f32(points), _f32(dirs)
    P = points.shape[0]
    bounds = _f32(map_data["map_bound"]).reshape(-1, 2, 3)
    tris = _f32(tris).reshape(-1, 3, 3)
    op, od = np.empty_like(points), np.empty_like(dirs)
    mask = np.zeros(P, np.uint8)
    src = _f32(map_data["empty_bound"]) if "map_source" in map_data else None
    ms = _f32(map_data["map_source"]) if "map_source" in map_data else None
    lib().orc_seal_bbox_map_to_origin(_p(points), _p(dirs), i64(P), _p(_f32(map_data["transform"])),
                                      _p(_f32(map_data["rotation"])), _p(_f32(map_data["scale"])),
                                      _p(_f32(map_data["center"])), _p(bounds), u32(bounds.shape[0]), _p(tris),
                                      u32(tris.shape[0]), _p(None if test_dir is None else _f32(test_dir)), _p(src),
                                      _p(ms), _p(op), _p(od), _p(mask))
    return op, od, mask.astype(bool)


def seal_modify_hsv(rgb, mod):
    rgb = _f32(rgb)
    out = np.empty_like(rgb)
    lib().orc_seal_modify_hsv(_p(rgb), i64(rgb.shape[0]), _p(_f32(mod)), _p(out))
    return out


def seal_modify_rgb(rgb, target_rgb, light_offset=0.0):
    rgb = _f32(rgb)
    out = np.empty_like(rgb)
    lib().orc_seal_modify_rgb(_p(rgb), i64(rgb.shape[0]), _p(_f32(target_rgb)), f32(light_offset), _p(out))
    return out


# ---------------------------------------------------------------- NGP field (nerf/network.py:99-128) in numpy


class NGPField:
    """fp32 restatement of NeRFNetwork.forward / backward (nerf/network.py:99-128, activation.py:5-17):
    sigma = exp(h[0]), geo = h[1:16]; rgb = sigmoid(MLP([SH16 | geo15 | grid2(32)])).  Tables and weights
    are plain numpy arrays with the state-dict shapes of SURVEY.md appendix B."""

    def __init__(self, emb_sigma, emb_color, w_s0, w_s1, w_c0, w_c1, w_c2, offsets, per_level_scale, H=16, bound=1.0):
        self.es, self.ec = _f32(emb_sigma), _f32(emb_color)
        self.w = [_f32(w_s0), _f32(w_s1), _f32(w_c0), _f32(w_c1), _f32(w_c2)]
        self.offsets, self.pls, self.H, self.bound = _i32(offsets), per_level_scale, H, bound

    def _enc(self, x, emb):
        u = ((_f32(x) + np.float32(self.bound)) / np.float32(2 * self.bound)).astype(np.float32)
        out, _ = grid_encode_forward(u, emb, self.offsets, self.pls, self.H)
        return u, np.ascontiguousarray(out.transpose(1, 0, 2).reshape(u.shape[0], -1))

    def forward(self, x, d, keep=False):
        ws0, ws1, wc0, wc1, wc2 = self.w
        u, f_s = self._enc(x, self.es)
        h1 = np.maximum(f_s @ ws0.T, 0)
        h2 = h1 @ ws1.T
        sigma = np.exp(h2[:, 0])
        sh, _ = sh_encode_forward(d, 4)
        _, f_c = self._enc(x, self.ec)
        cin = np.concatenate([sh, h2[:, 1:], f_c], axis=1)
        c1 = np.maximum(cin @ wc0.T, 0)
        c2 = np.maximum(c1 @ wc1.T, 0)
        rgb = 1.0 / (1.0 + np.exp(-(c2 @ wc2.T)))
        if keep:
            self._saved = (u, f_s, h1, h2, cin, c1, c2, rgb)
        return sigma.astype(np.float32), rgb.astype(np.float32)

    def density(self, x):
        ws0, ws1 = self.w[0], self.w[1]
        _, f_s = self._enc(x, self.es)
        h2 = np.maximum(f_s @ ws0.T, 0) @ ws1.T
        return np.exp(h2[:, 0]).astype(np.float32), h2[:, 1:].astype(np.float32)

    def backward(self, g_sigma, g_rgb):
        """Returns dict of grads for emb_sigma, emb_color and the 5 weights (after forward(keep=True))."""
        ws0, ws1, wc0, wc1, wc2 = self.w
        u, f_s, h1, h2, cin, c1, c2, rgb = self._saved
        g_o = _f32(g_rgb) * rgb * (1 - rgb)
        g_wc2 = g_o.T @ c2
        g_c2 = (g_o @ wc2) * (c2 > 0)
        g_wc1 = g_c2.T @ c1
        g_c1 = (g_c2 @ wc1) * (c1 > 0)
        g_wc0 = g_c1.T @ cin
        g_cin = g_c1 @ wc0
        g_h2 = np.empty_like(h2)
        g_h2[:, 0] = _f32(g_sigma) * np.exp(np.clip(h2[:, 0], -15, 15))
        g_h2[:, 1:] = g_cin[:, 16:31]
        g_fc = g_cin[:, 31:]
        g_ws1 = g_h2.T @ h1
        g_h1 = (g_h2 @ ws1) * (h1 > 0)
        g_ws0 = g_h1.T @ f_s
        g_fs = g_h1 @ ws0
        L = self.offsets.shape[0] - 1

        def tab(gf, shape):
            g = np.ascontiguousarray(_f32(gf).reshape(gf.shape[0], L, -1).transpose(1, 0, 2))
            return grid_encode_backward(g, u, shape, self.offsets, self.pls, self.H)

        return dict(emb_sigma=tab(g_fs, self.es.shape), emb_color=tab(g_fc, self.ec.shape), w_s0=g_ws0, w_s1=g_ws1,
                    w_c0=g_wc0, w_c1=g_wc1, w_c2=g_wc2)


# ---------------------------------------------------------------- TensoRF VM field (tensoRF/network.py:99-183)


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _vm_dims(mats, vecs):
    """mats[i] [R, H_i, W_i], vecs[i] [R, D_i] (reference shapes [1,R,H,W] / [1,R,D,1] squeezed) -> int32 [9] = H, W, D per plane"""
    return _i32([[m.shape[1], m.shape[2], v.shape[1]] for m, v in zip(mats, vecs)]).reshape(-1)


def _vm_squeeze(mats, vecs):
    mats = [_f32(np.asarray(m).reshape(m.shape[-3], m.shape[-2], m.shape[-1])) for m in mats]
    vecs = [_f32(np.asarray(v).reshape(v.shape[1], v.shape[2]) if np.asarray(v).ndim == 4 else v) for v in vecs]
    return mats, vecs


def vm_forward(xyz, mats, vecs, reduce, aabb=None):
    """get_sigma_feat (reduce=True, returns [M]) / the mat_feat * vec_feat product of get_color_feat (reduce=False,
    returns [M, 3R]); xyz are world coordinates if aabb is given (normalised like network.py:158), else already in [-1,1]."""
    xyz = _f32(xyz).reshape(-1, 3)
    mats, vecs = _vm_squeeze(mats, vecs)
    R, M = mats[0].shape[0], xyz.shape[0]
    out = np.empty(M if reduce else (M, 3 * R), np.float32)
    ab = _f32(aabb) if aabb is not None else None
    lib().orc_vm_forward(_p(xyz), u32(M), _p(ab), _ptr_array(mats), _ptr_array(vecs), _p(_vm_dims(mats, vecs)), u32(R),
                         cint(1 if reduce else 0), _p(out))
    return out


def vm_backward(xyz, mats, vecs, reduce, g, aabb=None):
    """gradients (reference layouts [R,H,W] / [R,D]) of the planes and lines for an upstream gradient g"""
    xyz = _f32(xyz).reshape(-1, 3)
    mats, vecs = _vm_squeeze(mats, vecs)
    R, M = mats[0].shape[0], xyz.shape[0]
    g = _f32(g)
    gm, gv = [np.empty_like(m) for m in mats], [np.empty_like(v) for v in vecs]
    ab = _f32(aabb) if aabb is not None else None
    lib().orc_vm_backward(_p(xyz), u32(M), _p(ab), _ptr_array(mats), _ptr_array(vecs), _p(_vm_dims(mats, vecs)), u32(R),
                          cint(1 if reduce else 0), _p(g), _ptr_array(gm), _ptr_array(gv))
    return gm, gv


class TensoRFField:
    """fp32 restatement of tensoRF/network.py NeRFNetwork.forward (:153-183): sigma = exp(sum of plane x line products),
    rgb = sigmoid(MLP([freq2(basis_mat . products) | freq2(d)])) with bias-free 150-128-128-3 layers; backward (after
    forward(keep=True)) returns the gradients of every parameter in the reference layouts."""

    def __init__(self, sigma_mat, sigma_vec, color_mat, color_vec, basis_mat, color_net, aabb=(-1, -1, -1, 1, 1, 1), degree=2):
        self.sm, self.sv = _vm_squeeze(sigma_mat, sigma_vec)
        self.cm, self.cv = _vm_squeeze(color_mat, color_vec)
        self.basis = _f32(basis_mat)
        self.w = [_f32(w) for w in color_net]
        self.aabb, self.degree = _f32(aabb), degree

    def forward(self, x, d, keep=False):
        sf = vm_forward(x, self.sm, self.sv, True, self.aabb)
        sigma = np.exp(sf)
        prod = vm_forward(x, self.cm, self.cv, False, self.aabb)
        cf = prod @ self.basis.T
        ecf, ed = freq_encode_forward(cf, self.degree), freq_encode_forward(d, self.degree)
        h = [np.concatenate([ecf, ed], axis=1)]
        for l, w in enumerate(self.w):
            z = h[-1] @ w.T
            h.append(np.maximum(z, 0) if l != len(self.w) - 1 else 1.0 / (1.0 + np.exp(-z)))
        if keep:
            self._saved = (_f32(x), sf, prod, cf, ecf, h)
        return sigma.astype(np.float32), h[-1].astype(np.float32)

    def density(self, x):
        return np.exp(vm_forward(x, self.sm, self.sv, True, self.aabb)).astype(np.float32)

    def backward(self, g_sigma, g_rgb):
        x, sf, prod, cf, ecf, h = self._saved
        rgb = h[-1]
        g = _f32(g_rgb) * rgb * (1 - rgb)
        g_w = [None] * len(self.w)
        for l in range(len(self.w) - 1, -1, -1):
            g_w[l] = g.T @ h[l]
            g = g @ self.w[l]
            if l > 0:
                g = g * (h[l] > 0)
        g_ecf = np.ascontiguousarray(g[:, :ecf.shape[1]])
        g_cf = freq_encode_backward(g_ecf, ecf, cf.shape[1], self.degree)
        g_basis = g_cf.T @ prod
        g_prod = g_cf @ self.basis
        g_cm, g_cv = vm_backward(x, self.cm, self.cv, False, g_prod, self.aabb)
        g_sf = _f32(g_sigma) * np.exp(np.clip(sf, -15, 15))      # trunc_exp backward (activation.py:14-17)
        g_sm, g_sv = vm_backward(x, self.sm, self.sv, True, g_sf, self.aabb)
        return dict(sigma_mat=g_sm, sigma_vec=g_sv, color_mat=g_cm, color_vec=g_cv, basis_mat=g_basis, color_net=g_w)


# ---------------------------------------------------------------- distillation losses (numpy)


def pretrain_loss(sigma_s, rgb_s, sigma_t, rgb_t):
    """SealNeRF/trainer.py:456-469: L1(sigma) + L1(rgb), mean reductions.  Returns loss, dL/dsigma_s, dL/drgb_s."""
    ds, dc = sigma_s - sigma_t, rgb_s - rgb_t
    loss = np.abs(ds).mean() + np.abs(dc).mean()
    return loss, np.sign(ds) / ds.size, np.sign(dc) / dc.size


def finetune_loss(image_s, depth_s, image_t, depth_t):
    """nerf/utils.py:484-489,530: mean_rays(mean_c (rgb - gt)^2) + mean|depth - gt_depth|.
    Returns loss, dL/dimage, dL/ddepth (the latter is dropped by the compositor, raymarching.py:275)."""
    N = image_s.shape[0]
    dr = image_s - image_t
    dd = depth_s - depth_t
    loss = (dr ** 2).mean(-1).mean() + np.abs(dd).mean()
    return loss, 2 * dr / (3 * N), np.sign(dd) / N


# ---------------------------------------------------------------------------- optimizer-side state
class Optimizer:
    """The reference's optimisation step restated in numpy float64 -> float32 (test infrastructure):
      * torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15)                     main_SealNeRF.py:283-284
      * LambdaLR  lr_t = lr * 0.1 ** min(t / iters, 1), stepped every step  main_SealNeRF.py:287-288, nerf/utils.py:861-862
      * torch.cuda.amp.GradScaler: scale(loss), step skipped on non-finite gradients, update (growth / backoff)
                                                                            nerf/utils.py:361, 857-859
      * torch_ema.ExponentialMovingAverage(decay=0.95).update()             nerf/utils.py:356-357, 882-883
    torch and torch_ema are third-party (not in the reference tree); Adam / GradScaler / LambdaLR are pinned in
    tests/test_gpu_fused.py against the installed torch itself, the EMA rule is torch_ema's published one (parity unpinned)."""

    def __init__(self, params, lr, iters=None, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000,
                 beta1=0.9, beta2=0.99, eps=1e-15, ema_decay=None):
        self.p = np.array(params, dtype=np.float64)
        self.m, self.v = np.zeros_like(self.p), np.zeros_like(self.p)
        self.lr, self.iters, self.b1, self.b2, self.eps = lr, iters, beta1, beta2, eps
        self.scale, self.growth, self.backoff, self.interval, self.tracker = float(init_scale), growth_factor, backoff_factor, growth_interval, 0
        self.t_sched, self.t_opt = 0, 0
        self.ema_decay, self.ema_n = ema_decay, 0
        self.shadow = self.p.copy() if ema_decay is not None else None

    def step(self, scaled_grad):
        g = np.asarray(scaled_grad, dtype=np.float64)
        lr_t = self.lr * (0.1 ** min(self.t_sched / float(self.iters), 1.0) if self.iters else 1.0)
        found_inf = not np.isfinite(g).all()
        if not found_inf:
            g = g / self.scale
            self.t_opt += 1
            self.m = self.b1 * self.m + (1 - self.b1) * g
            self.v = self.b2 * self.v + (1 - self.b2) * g * g
            bc1, bc2 = 1 - self.b1 ** self.t_opt, 1 - self.b2 ** self.t_opt
            self.p = self.p - (lr_t / bc1) * self.m / (np.sqrt(self.v) / np.sqrt(bc2) + self.eps)
            self.tracker += 1
            if self.tracker == self.interval:
                self.scale *= self.growth
                self.tracker = 0
        else:
            self.scale *= self.backoff
            self.tracker = 0
        self.t_sched += 1
        return found_inf

    def ema_update(self):
        self.ema_n += 1
        d = min(self.ema_decay, (1.0 + self.ema_n) / (10.0 + self.ema_n))
        self.shadow = self.shadow - (1.0 - d) * (self.shadow - self.p)


# ---------------------------------------------------------------------------- brush / anchor / texture colour
def seal_brush_map_to_origin(points, bounds, tris, normal_expand, center, border_points, attenuation_distance, mode="linear",
                             test_dir=None):
    """SealBrushMapper.map_to_origin (seal_utils.py:408-453); returns (points', mask)"""
    points, bounds, tris, border = _f32(points), _f32(bounds).reshape(-1, 2, 3), _f32(tris), _f32(border_points).reshape(-1, 3)
    out, mask = np.empty_like(points), np.empty(points.shape[0], dtype=np.uint8)
    lib().orc_seal_brush_map_to_origin(_p(points), i64(points.shape[0]), _p(bounds), u32(bounds.shape[0]), _p(tris), u32(tris.shape[0]),
                                       _p(_f32(test_dir)) if test_dir is not None else None, _p(_f32(normal_expand)), _p(_f32(center)),
                                       _p(border), u32(border.shape[0]), f32(attenuation_distance), cint(0 if mode == "linear" else 1),
                                       _p(out), _p(mask))
    return out, mask.astype(bool)


def seal_anchor_map_to_origin(points, bounds, tris, v_anchor, v_offset, v_h, len_h, radius, scale, test_dir=None):
    """SealAnchorMapper.map_to_origin (seal_utils.py:514-570); returns (points', valid_mask)"""
    points, bounds, tris = _f32(points), _f32(bounds).reshape(-1, 2, 3), _f32(tris)
    out, mask = np.empty_like(points), np.empty(points.shape[0], dtype=np.uint8)
    lib().orc_seal_anchor_map_to_origin(_p(points), i64(points.shape[0]), _p(bounds), u32(bounds.shape[0]), _p(tris), u32(tris.shape[0]),
                                        _p(_f32(test_dir)) if test_dir is not None else None, _p(_f32(v_anchor)), _p(_f32(v_offset)),
                                        _p(_f32(v_h)), f32(len_h), f32(radius), _p(_f32(scale)), _p(out), _p(mask))
    return out, mask.astype(bool)


def seal_map_color_image(points, rgb, image, image_mask, v_norm, v_o, v_w, v_h, light_offset=0.0):
    """texture branch of SealMapper.map_color (seal_utils.py:58-79)"""
    points, rgb, image, image_mask = _f32(points), _f32(rgb), _f32(image), _f32(image_mask)
    out = np.empty_like(rgb)
    lib().orc_seal_map_color_image(_p(points), _p(rgb), i64(rgb.shape[0]), _p(image), _p(image_mask), u32(image.shape[0]), u32(image.shape[1]),
                                   _p(_f32(v_norm)), _p(_f32(v_o)), _p(_f32(v_w)), _p(_f32(v_h)), f32(light_offset), _p(out))
    return out
