"""Deterministic synthetic workload (no dataset / checkpoint exists on the box): a Lego-shaped
camera rig, an analytic occupancy grid, random-init tables / MLP weights and an axis-aligned
bbox edit.  Definitions follow SURVEY.md section 8(d); the camera recipe mirrors the reference's
``get_rays`` (nerf/utils.py:54-139) and ``rand_poses`` look-at (nerf/provider.py:57-91).

Everything here is host-side numpy / torch-CPU so that the oracle and the CUDA path consume
bit-identical inputs.
"""
import math

import numpy as np

H_IMG = W_IMG = 800
CAMERA_ANGLE_X = 0.6911112            # NeRF-synthetic Lego transforms_train.json
CAMERA_RADIUS = 4.031128 * 0.8        # --scale 0.8  (nerf/provider.py:19-27)
FOCAL = 0.5 * W_IMG / math.tan(0.5 * CAMERA_ANGLE_X)   # nerf/provider.py:264
CX = CY = 0.5 * W_IMG
GRID_H = 128                          # nerf/renderer.py:74
BOUND = 1.0
MAX_STEPS = 1024                      # main_nerf.py:29
MIN_NEAR = 0.2
NUM_LEVELS, LEVEL_DIM, BASE_RES, LOG2_T, DESIRED_RES = 16, 2, 16, 19, 2048


def _normalize(v):
    return v / (np.linalg.norm(v, axis=-1, keepdims=True) + 1e-10)


def make_poses(n=100, seed=0):
    """cam2world [n,4,4] float32 on the upper hemisphere, looking at the origin."""
    rng = np.random.default_rng(seed)
    thetas = rng.uniform(math.pi / 6, 0.47 * math.pi, n)
    phis = rng.uniform(0, 2 * math.pi, n)
    centers = np.stack([CAMERA_RADIUS * np.sin(thetas) * np.sin(phis), CAMERA_RADIUS * np.cos(thetas),
                        CAMERA_RADIUS * np.sin(thetas) * np.cos(phis)], -1)
    fwd = -_normalize(centers)
    up = np.tile(np.array([0.0, -1.0, 0.0]), (n, 1))
    right = _normalize(np.cross(fwd, up))
    up = _normalize(np.cross(right, fwd))
    poses = np.tile(np.eye(4), (n, 1, 1))
    poses[:, :3, :3] = np.stack([right, up, fwd], -1)
    poses[:, :3, 3] = centers
    return poses.astype(np.float32)


_POSES = None


def poses():
    global _POSES
    if _POSES is None:
        _POSES = make_poses()
    return _POSES


def rays_from_pixels(pose, inds):
    """rays_o, rays_d [N,3] float32 for flat pixel ids of one view (get_rays, nerf/utils.py:116-136)."""
    inds = np.asarray(inds, dtype=np.int64)
    i = (inds % W_IMG).astype(np.float32) + np.float32(0.5)
    j = (inds // W_IMG).astype(np.float32) + np.float32(0.5)
    xs = (i - np.float32(CX)) / np.float32(FOCAL)
    ys = (j - np.float32(CY)) / np.float32(FOCAL)
    dirs = np.stack([xs, ys, np.ones_like(xs)], -1)
    dirs = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)
    rays_d = (dirs @ pose[:3, :3].T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, np.ascontiguousarray(rays_d)


def rays_for_step(step, n_rays, views_per_step=None):
    """The ray batch of training step `step`: pixels drawn uniformly (with replacement) from one view
    (n_rays <= 4096, the reference's per-step recipe) or from ceil(n_rays/4096) views for big batches."""
    rng = np.random.default_rng(1 + step)
    ps = poses()
    per_view = 4096 if views_per_step is None else max(1, n_rays // views_per_step)
    ro, rd = [], []
    left = n_rays
    while left > 0:
        k = min(per_view, left)
        v = int(rng.integers(0, ps.shape[0]))
        inds = rng.integers(0, H_IMG * W_IMG, k)
        o, d = rays_from_pixels(ps[v], inds)
        ro.append(o)
        rd.append(d)
        left -= k
    return np.concatenate(ro), np.concatenate(rd)


def full_image_rays(view=0):
    return rays_from_pixels(poses()[view], np.arange(H_IMG * W_IMG))


def _morton3D(x, y, z):
    def expand(v):
        v = v.astype(np.uint32)
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    return expand(x) | (expand(y) << np.uint32(1)) | (expand(z) << np.uint32(2))


def lego_like_density(p):
    """Analytic occupancy: union of a box |x|<.35,|y|<.2,|z|<.5 and a sphere r=.3 at (0,.25,0)."""
    box = (np.abs(p[..., 0]) < 0.35) & (np.abs(p[..., 1]) < 0.2) & (np.abs(p[..., 2]) < 0.5)
    sph = ((p[..., 0]) ** 2 + (p[..., 1] - 0.25) ** 2 + (p[..., 2]) ** 2) < 0.3 ** 2
    return box | sph


def lego_like_occupancy():
    """Returns (bitfield uint8[128^3/8], density_grid float32[1,128^3]) in morton order; 100 inside, 0 outside."""
    H = GRID_H
    ax = np.arange(H)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    centers = (np.stack([X, Y, Z], -1).astype(np.float32) + 0.5) / H * 2 - 1
    occ = lego_like_density(centers)
    idx = _morton3D(X.reshape(-1), Y.reshape(-1), Z.reshape(-1)).astype(np.int64)
    grid = np.zeros(H ** 3, np.float32)
    grid[idx] = np.where(occ.reshape(-1), 100.0, 0.0)
    bits = np.packbits(grid > 10.0, bitorder="little")
    return bits.astype(np.uint8), grid.reshape(1, -1)


def grid_offsets():
    """gridencoder/grid.py:100-127 for the NGP config (D3 L16 C2 H16 T2^19, desired 2048*bound)."""
    pls = np.exp2(np.log2(DESIRED_RES * BOUND / BASE_RES) / (NUM_LEVELS - 1))
    offs, off = [], 0
    for i in range(NUM_LEVELS):
        res = int(np.ceil(BASE_RES * pls ** i))
        n = min(2 ** LOG2_T, (res + 1) ** 3)
        n = int(np.ceil(n / 8) * 8)
        offs.append(off)
        off += n
    offs.append(off)
    return np.array(offs, np.int32), float(pls)


def field_params(kind="student", small=False):
    """Random-init NGP field parameters with the reference's state-dict shapes.
    student: tables U(-1e-4,1e-4) seeds 2/3; teacher: tables N(0,0.1) seeds 4/5 so sigma/rgb are
    non-degenerate.  MLP weights: nn.Linear default init U(+-1/sqrt(fan_in)), seeds 6 (student) / 7."""
    offs, _ = grid_offsets()
    n = int(offs[-1])
    if kind == "student":
        e_s = np.random.default_rng(2).uniform(-1e-4, 1e-4, (n, 2)).astype(np.float32)
        e_c = np.random.default_rng(3).uniform(-1e-4, 1e-4, (n, 2)).astype(np.float32)
        wr = np.random.default_rng(6)
    else:
        e_s = (np.random.default_rng(4).standard_normal((n, 2)) * 0.1).astype(np.float32)
        e_c = (np.random.default_rng(5).standard_normal((n, 2)) * 0.1).astype(np.float32)
        wr = np.random.default_rng(7)

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return wr.uniform(-b, b, (o, i)).astype(np.float32)

    return dict(emb_sigma=e_s, emb_color=e_c, w_s0=lin(64, 32), w_s1=lin(16, 64), w_c0=lin(64, 63), w_c1=lin(64, 64),
                w_c2=lin(3, 64))


_BOX_FACES = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                       [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])


def bbox_edit(lo=(-0.15, -0.15, -0.15), hi=(0.15, 0.15, 0.15), translate=(0.3, 0.0, 0.0), rot_z=0.0,
              scale=(1.0, 1.0, 1.0), map_source=None, hsv=None):
    """map_data + target-box triangles of a SealBBoxMapper-equivalent edit (seal_utils.py:168-236) built
    without trimesh/pytorch3d: source box -> scale about its centre -> rigid transform."""
    lo, hi, scale = np.asarray(lo, np.float64), np.asarray(hi, np.float64), np.asarray(scale, np.float64)
    T = np.eye(4)
    c, s = math.cos(rot_z), math.sin(rot_z)
    T[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    T[:3, 3] = translate
    center = (lo + hi) / 2
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    v = (corners - center) * scale + center
    v = (T[:3, :3] @ v.T).T + T[:3, 3]
    tris = v[_BOX_FACES].astype(np.float32)
    to_b = np.stack([v.min(0), v.max(0)])
    from_b = np.stack([lo, hi])
    md = {
        "force_fill_bound": np.stack([to_b, from_b]).astype(np.float32),
        "map_bound": to_b.astype(np.float32),
        "transform": np.linalg.inv(T).astype(np.float32),
        "rotation": np.linalg.inv(T[:3, :3]).astype(np.float32),
        "scale": (1.0 / scale).astype(np.float32),
        "center": center.astype(np.float32),
    }
    if map_source is not None:
        md["empty_bound"] = from_b.astype(np.float32)
        md["map_source"] = np.asarray(map_source, np.float32)
    if hsv is not None:
        md["hsv"] = np.asarray(hsv, np.float32)
    return md, tris
