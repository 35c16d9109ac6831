// Seal-3D per-sample proxy mapping (teacher side) for sm_100a: bbox mapper + colour edits + bitfield force-fill.
//
// Replaces the ~25 ATen launches, dense [2P,F,3] temporaries and boolean-index gathers of
// SealNeRF/seal_utils.py: SealMapper.map_mask (:132-153), points_in_mesh / moller_trumbore (:630-685),
// SealBBoxMapper.map_to_origin (:237-279), modify_hsv / modify_rgb (:739-769, color_utils.py:31-63) and
// SealNeRFRenderer.hack_bitfield (SealNeRF/renderer.py:21-66) with one kernel per step:
//   * map: one thread per sample -- zero-row test, strict AABB test against every map bound, two opposite
//     Moeller-Trumbore rays against the F triangles held in shared memory ("hit in both directions"),
//     then the inverse affine on x and the inverse rotation on d, optional map_source teleport;
//   * colour: HSV shift, or H/S replacement with V re-lighting around the mean V of the masked samples of
//     the call (one block-reduced sum per call, same semantics as torch.mean over rgbs[mask]).
#include "common.cuh"

namespace {

struct MapConsts {
    float transform[16];  // inverse source->target 4x4, row-major
    float rotation[9];    // inverse 3x3
    float scale[3];       // 1 / scale
    float center[3];      // from_center
    float test_dir[3];
    float src_lo[3], src_hi[3], map_source[3];
    int has_source;
};

__device__ __forceinline__ bool mt_any_hit(const float3 o, const float3 d, const float *__restrict__ tris, uint32_t F) {
    for (uint32_t f = 0; f < F; f++) {
        const float *t = tris + f * 9;
        const float3 v0 = make_float3(t[0], t[1], t[2]);
        const float3 e1 = make_float3(t[3] - v0.x, t[4] - v0.y, t[5] - v0.z);
        const float3 e2 = make_float3(t[6] - v0.x, t[7] - v0.y, t[8] - v0.z);
        const float3 n = make_float3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
        const float invdet = 1.0f / -((d.x * n.x + d.y * n.y + d.z * n.z) + 1e-8f);
        const float3 a0 = make_float3(o.x - v0.x, o.y - v0.y, o.z - v0.z);
        const float3 da0 = make_float3(a0.y * d.z - a0.z * d.y, a0.z * d.x - a0.x * d.z, a0.x * d.y - a0.y * d.x);
        const float u = (da0.x * e2.x + da0.y * e2.y + da0.z * e2.z) * invdet;
        const float v = -(da0.x * e1.x + da0.y * e1.y + da0.z * e1.z) * invdet;
        const float tt = (a0.x * n.x + a0.y * n.y + a0.z * n.z) * invdet;
        if (tt >= 0.0f && u >= 0.0f && v >= 0.0f && (u + v) <= 1.0f) return true;
    }
    return false;
}

// SealMapper.map_mask (seal_utils.py:132-153): non-zero row, strictly inside one of the map bounds, and inside the mesh
// ("hit in both directions" along the test direction, points_in_mesh :667-685)
__device__ __forceinline__ bool in_map_region(const float3 x, const float *__restrict__ bounds, uint32_t nb, const float *__restrict__ s_tris,
                                              uint32_t F, const float3 d0) {
    bool m = false;
    if (x.x != 0.0f && x.y != 0.0f && x.z != 0.0f) {  // points.all(1): padding rows never enter the mask
        for (uint32_t i = 0; i < nb && !m; i++) {
            const float *lo = bounds + i * 6, *hi = lo + 3;
            m = hi[0] > x.x && x.x > lo[0] && hi[1] > x.y && x.y > lo[1] && hi[2] > x.z && x.z > lo[2];
        }
    }
    if (m) m = mt_any_hit(x, d0, s_tris, F) && mt_any_hit(x, make_float3(-d0.x, -d0.y, -d0.z), s_tris, F);
    return m;
}

// seal_utils.py:728-736 project_points
__device__ __forceinline__ float3 project_point(const float3 n, const float3 o, const float3 p) {
    const float3 v = make_float3(p.x - o.x, p.y - o.y, p.z - o.z);
    const float s = (v.x * n.x + v.y * n.y + v.z * n.z) / (n.x * n.x + n.y * n.y + n.z * n.z);
    return make_float3(p.x - s * n.x, p.y - s * n.y, p.z - s * n.z);
}

struct BrushConsts {
    float test_dir[3], normal_expand[3], center[3], att;
    int mode;  // 0 linear, 1 dry
};

// SealBrushMapper.map_to_origin (seal_utils.py:408-453): inside samples move against the brush normal, with the linear
// attenuation towards the stroke border (nearest of the K border points of the projected sample); shared memory holds
// the triangles followed by the border points
__global__ void __launch_bounds__(256)
k_brush_map(const float *__restrict__ points, uint32_t P, const BrushConsts c, const float *__restrict__ bounds, uint32_t nb,
            const float *__restrict__ tris_g, uint32_t F, const float *__restrict__ border_g, uint32_t K,
            float *__restrict__ out_points, uint8_t *__restrict__ mask) {
    extern __shared__ float s_tris[];
    float *s_border = s_tris + F * 9;
    for (uint32_t i = threadIdx.x; i < F * 9; i += blockDim.x) s_tris[i] = tris_g[i];
    for (uint32_t i = threadIdx.x; i < K * 3; i += blockDim.x) s_border[i] = border_g[i];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float3 x = make_float3(points[(size_t)p * 3], points[(size_t)p * 3 + 1], points[(size_t)p * 3 + 2]);
    const bool m = in_map_region(x, bounds, nb, s_tris, F, make_float3(c.test_dir[0], c.test_dir[1], c.test_dir[2]));
    float3 o = x;
    if (m && c.mode == 0) {
        const float3 n = make_float3(c.normal_expand[0], c.normal_expand[1], c.normal_expand[2]);
        const float3 pr = project_point(n, make_float3(c.center[0], c.center[1], c.center[2]), x);
        float best = INFINITY;
        for (uint32_t k = 0; k < K; k++) {
            const float dx = pr.x - s_border[k * 3], dy = pr.y - s_border[k * 3 + 1], dz = pr.z - s_border[k * 3 + 2];
            best = fminf(best, sqrtf(dx * dx + dy * dy + dz * dz));
        }
        o = make_float3(x.x - n.x, x.y - n.y, x.z - n.z);
        if (c.att > best) {
            const float w = fabsf(c.att - best) / c.att;
            o = make_float3(o.x + w * n.x, o.y + w * n.y, o.z + w * n.z);
        }
    }
    out_points[(size_t)p * 3] = o.x; out_points[(size_t)p * 3 + 1] = o.y; out_points[(size_t)p * 3 + 2] = o.z;
    mask[p] = (uint8_t)m;
}

struct AnchorConsts {
    float test_dir[3], v_anchor[3], v_offset[3], v_h[3], scale[3], len_h, radius;
};

// pass 1 of the anchor mapper: does ANY sample fall into the map region (the reference's early exit, seal_utils.py:518-520)
__global__ void __launch_bounds__(256)
k_anchor_any(const float *__restrict__ points, uint32_t P, const AnchorConsts c, const float *__restrict__ bounds, uint32_t nb,
             const float *__restrict__ tris_g, uint32_t F, int *__restrict__ flag) {
    extern __shared__ float s_tris[];
    for (uint32_t i = threadIdx.x; i < F * 9; i += blockDim.x) s_tris[i] = tris_g[i];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    bool m = false;
    if (p < P) {
        const float3 x = make_float3(points[(size_t)p * 3], points[(size_t)p * 3 + 1], points[(size_t)p * 3 + 2]);
        m = in_map_region(x, bounds, nb, s_tris, F, make_float3(c.test_dir[0], c.test_dir[1], c.test_dir[2]));
    }
    if (__any_sync(0xffffffffu, m) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// pass 2 (seal_utils.py:522-560): the cone / plane-side test and the pull towards the anchor run on EVERY sample once the
// flag is set, and the returned mask is the cone mask (reference behaviour, kept)
__global__ void __launch_bounds__(256)
k_anchor_map(const float *__restrict__ points, uint32_t P, const AnchorConsts c, const int *__restrict__ flag,
             float *__restrict__ out_points, uint8_t *__restrict__ mask) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float3 x = make_float3(points[(size_t)p * 3], points[(size_t)p * 3 + 1], points[(size_t)p * 3 + 2]);
    float3 o = x;
    bool valid = false;
    if (*flag) {
        const float3 vh = make_float3(c.v_h[0], c.v_h[1], c.v_h[2]), va = make_float3(c.v_anchor[0], c.v_anchor[1], c.v_anchor[2]);
        const float3 pr = project_point(vh, va, x);
        const float3 vp = make_float3(pr.x - x.x, pr.y - x.y, pr.z - x.z);
        const float dist = sqrtf(vp.x * vp.x + vp.y * vp.y + vp.z * vp.z);
        const float os = dist / c.len_h;
        const float3 po = make_float3(pr.x - os * c.v_offset[0], pr.y - os * c.v_offset[1], pr.z - os * c.v_offset[2]);
        const float3 q = make_float3(po.x - va.x, po.y - va.y, po.z - va.z);
        const float pad = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
        const bool cone = (pad <= c.radius) && (dist / (c.radius - pad) < c.len_h / c.radius * 1.1f);
        const bool side = (vp.x * vh.x + vp.y * vh.y + vp.z * vh.z) > 0.0f;
        valid = cone && side;
        if (valid) {
            const float f = -((c.len_h - dist) / 10.0f);
            const float mp[3] = {po.x - f * vh.x / c.len_h, po.y - f * vh.y / c.len_h, po.z - f * vh.z / c.len_h};
            o = make_float3((mp[0] - va.x) * c.scale[0] + va.x, (mp[1] - va.y) * c.scale[1] + va.y, (mp[2] - va.z) * c.scale[2] + va.z);
        }
    }
    out_points[(size_t)p * 3] = o.x; out_points[(size_t)p * 3 + 1] = o.y; out_points[(size_t)p * 3 + 2] = o.z;
    mask[p] = (uint8_t)valid;
}

__global__ void __launch_bounds__(256)
k_bbox_map(const float *__restrict__ points, const float *__restrict__ dirs, uint32_t P, const MapConsts c,
           const float *__restrict__ bounds, uint32_t nb, const float *__restrict__ tris_g, uint32_t F,
           float *__restrict__ out_points, float *__restrict__ out_dirs, uint8_t *__restrict__ mask) {
    extern __shared__ float s_tris[];
    for (uint32_t i = threadIdx.x; i < F * 9; i += blockDim.x) s_tris[i] = tris_g[i];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const float3 x = make_float3(points[(size_t)p * 3], points[(size_t)p * 3 + 1], points[(size_t)p * 3 + 2]);
    bool m = in_map_region(x, bounds, nb, s_tris, F, make_float3(c.test_dir[0], c.test_dir[1], c.test_dir[2]));
    float3 o = x;
    if (c.has_source && c.src_hi[0] > x.x && x.x > c.src_lo[0] && c.src_hi[1] > x.y && x.y > c.src_lo[1] &&
        c.src_hi[2] > x.z && x.z > c.src_lo[2])
        o = make_float3(c.map_source[0], c.map_source[1], c.map_source[2]);
    float3 dd = make_float3(0, 0, 0);
    if (dirs) dd = make_float3(dirs[(size_t)p * 3], dirs[(size_t)p * 3 + 1], dirs[(size_t)p * 3 + 2]);
    if (m) {
        float t[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float *r = c.transform + i * 4;
            const float tp = r[0] * x.x + r[1] * x.y + r[2] * x.z + r[3];
            t[i] = (tp - c.center[i]) * c.scale[i] + c.center[i];
        }
        o = make_float3(t[0], t[1], t[2]);
        if (dirs) {
            const float3 q = dd;
            dd = make_float3(c.rotation[0] * q.x + c.rotation[1] * q.y + c.rotation[2] * q.z,
                             c.rotation[3] * q.x + c.rotation[4] * q.y + c.rotation[5] * q.z,
                             c.rotation[6] * q.x + c.rotation[7] * q.y + c.rotation[8] * q.z);
        }
    }
    out_points[(size_t)p * 3] = o.x; out_points[(size_t)p * 3 + 1] = o.y; out_points[(size_t)p * 3 + 2] = o.z;
    if (dirs) { out_dirs[(size_t)p * 3] = dd.x; out_dirs[(size_t)p * 3 + 1] = dd.y; out_dirs[(size_t)p * 3 + 2] = dd.z; }
    mask[p] = (uint8_t)m;
}

// color_utils.py:31-43
__device__ __forceinline__ float3 rgb2hsv(float3 c) {
    float cmax = c.x; int idx = 0;
    if (c.y > cmax) { cmax = c.y; idx = 1; }
    if (c.z > cmax) { cmax = c.z; idx = 2; }
    const float cmin = fminf(c.x, fminf(c.y, c.z));
    const float delta = cmax - cmin;
    float h;
    if (delta == 0.0f) h = 0.0f;
    else if (idx == 0) { h = fmodf((c.y - c.z) / delta, 6.0f); if (h < 0.0f) h += 6.0f; }
    else if (idx == 1) h = (c.z - c.x) / delta + 2.0f;
    else h = (c.x - c.y) / delta + 4.0f;
    return make_float3(h / 6.0f, cmax == 0.0f ? 0.0f : delta / cmax, cmax);
}
// color_utils.py:46-63
__device__ __forceinline__ float3 hsv2rgb(float3 hsv) {
    const float c = hsv.z * hsv.y;
    float hm = fmodf(hsv.x * 6.0f, 2.0f);
    if (hm < 0.0f) hm += 2.0f;
    const float x = c * (-fabsf(hm - 1.0f) + 1.0f);
    const float m = hsv.z - c;
    const int idx = ((int)(uint8_t)(int)(hsv.x * 6.0f)) % 6;
    float3 r;
    switch (idx) {
        case 0: r = make_float3(c, x, 0); break;
        case 1: r = make_float3(x, c, 0); break;
        case 2: r = make_float3(0, c, x); break;
        case 3: r = make_float3(0, x, c); break;
        case 4: r = make_float3(x, 0, c); break;
        default: r = make_float3(c, 0, x); break;
    }
    return make_float3(r.x + m, r.y + m, r.z + m);
}

__global__ void k_modify_hsv(float *__restrict__ rgbs, const uint8_t *__restrict__ mask, uint32_t M, float m0, float m1, float m2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || (mask && !mask[i])) return;
    float3 hsv = rgb2hsv(make_float3(rgbs[(size_t)i * 3], rgbs[(size_t)i * 3 + 1], rgbs[(size_t)i * 3 + 2]));
    hsv.x += m0; hsv.y += m1; hsv.z += m2;
    const float3 o = hsv2rgb(hsv);
    rgbs[(size_t)i * 3] = o.x; rgbs[(size_t)i * 3 + 1] = o.y; rgbs[(size_t)i * 3 + 2] = o.z;
}

// mean V of the masked rows, in a fixed order (no float atomics: the edited colours -- the distillation targets -- are then the same
// bits on every run): every CTA leaves one (sum, count) pair, the last CTA to finish adds the pairs in index order.
constexpr uint32_t kStatBlocks = 1024;
// part: per-call scratch, kStatBlocks pairs followed by the ticket word (zero on entry)
__global__ void __launch_bounds__(256) k_v_stats(const float *__restrict__ rgbs, const uint8_t *__restrict__ mask, uint32_t M, float *__restrict__ stats,
                                                 float2 *__restrict__ part) {
    unsigned int *ticket = reinterpret_cast<unsigned int *>(part + kStatBlocks);
    __shared__ float2 s_part[8];
    __shared__ bool s_last;
    float s = 0.0f, n = 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        s += fmaxf(rgbs[(size_t)i * 3], fmaxf(rgbs[(size_t)i * 3 + 1], rgbs[(size_t)i * 3 + 2]));
        n += 1.0f;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); n += __shfl_xor_sync(0xffffffffu, n, d); }
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = make_float2(s, n);
    __syncthreads();
    if (threadIdx.x == 0) {
        float2 t = s_part[0];
        for (int w = 1; w < 8; w++) { t.x += s_part[w].x; t.y += s_part[w].y; }
        part[blockIdx.x] = t;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: warp 0 sums the pairs in a fixed tree (lane-strided partial sums, then a butterfly)
    if (threadIdx.x < 32) {
        float a = 0.0f, b = 0.0f;
        for (uint32_t i = threadIdx.x; i < gridDim.x; i += 32) { const float2 t = __ldcg(part + i); a += t.x; b += t.y; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); b += __shfl_xor_sync(0xffffffffu, b, d); }
        if (threadIdx.x == 0) { stats[0] = a; stats[1] = b; }
    }
}

static cudaError_t v_stats_launch(const float *rgbs, const uint8_t *mask, uint32_t M, float *d_stats, cudaStream_t st) {
    float2 *part = nullptr;
    cudaError_t e = scratch_alloc((void **)&part, (kStatBlocks + 1) * sizeof(float2), st);
    if (e != cudaSuccess) return e;
    cudaMemsetAsync(part + kStatBlocks, 0, sizeof(float2), st);
    k_v_stats<<<min(div_up(M, 256u), kStatBlocks), 256, 0, st>>>(rgbs, mask, M, d_stats, part);
    return cudaFreeAsync(part, st);
}

__global__ void k_modify_rgb(float *__restrict__ rgbs, const uint8_t *__restrict__ mask, uint32_t M, float3 target_hsv,
                             float light, const float *__restrict__ stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || (mask && !mask[i])) return;
    const float mean = stats[0] / fmaxf(stats[1], 1.0f);
    const float v = fmaxf(rgbs[(size_t)i * 3], fmaxf(rgbs[(size_t)i * 3 + 1], rgbs[(size_t)i * 3 + 2]));
    const float3 o = hsv2rgb(make_float3(target_hsv.x, target_hsv.y, fminf(1.0f, fmaxf(0.0f, target_hsv.z + (v - mean) + light))));
    rgbs[(size_t)i * 3] = o.x; rgbs[(size_t)i * 3 + 1] = o.y; rgbs[(size_t)i * 3 + 2] = o.z;
}

struct ImageConsts {
    float norm[3], o[3], ow[3], oh[3], low2, loh2, light;   // low2 = len_ow ** 2 exactly as the reference forms it
    uint32_t H, W;
};

// texture branch of SealMapper.map_color (seal_utils.py:58-79): per-sample target colour = image[pixel of the projected point]
__global__ void k_modify_rgb_image(float *__restrict__ rgbs, const float *__restrict__ points, const uint8_t *__restrict__ mask, uint32_t M,
                                   const ImageConsts c, const float *__restrict__ image, const float *__restrict__ alpha,
                                   const float *__restrict__ stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || (mask && !mask[i])) return;
    const float3 p = make_float3(points[(size_t)i * 3], points[(size_t)i * 3 + 1], points[(size_t)i * 3 + 2]);
    const float3 pr = project_point(make_float3(c.norm[0], c.norm[1], c.norm[2]), make_float3(c.o[0], c.o[1], c.o[2]), p);
    const float3 op = make_float3(pr.x - c.o[0], pr.y - c.o[1], pr.z - c.o[2]);
    float fw = floorf((op.x * c.ow[0] + op.y * c.ow[1] + op.z * c.ow[2]) / c.low2 * (float)c.W);
    float fh = floorf((op.x * c.oh[0] + op.y * c.oh[1] + op.z * c.oh[2]) / c.loh2 * (float)c.H);
    fw = fminf(fmaxf(0.0f, fw), (float)(c.W - 1));
    fh = fminf(fmaxf(0.0f, fh), (float)(c.H - 1));
    const size_t pix = (size_t)fh * c.W + (size_t)fw;
    const float3 col = make_float3(rgbs[(size_t)i * 3], rgbs[(size_t)i * 3 + 1], rgbs[(size_t)i * 3 + 2]);
    const float3 thsv = rgb2hsv(make_float3(image[pix * 3], image[pix * 3 + 1], image[pix * 3 + 2]));
    const float mean = stats[0] / fmaxf(stats[1], 1.0f);
    const float v = fmaxf(col.x, fmaxf(col.y, col.z));
    const float3 mod = hsv2rgb(make_float3(thsv.x, thsv.y, fminf(1.0f, fmaxf(0.0f, thsv.z + (v - mean) + c.light))));
    const float a = alpha[pix];
    rgbs[(size_t)i * 3] = a * mod.x + (1.0f - a) * col.x;
    rgbs[(size_t)i * 3 + 1] = a * mod.y + (1.0f - a) * col.y;
    rgbs[(size_t)i * 3 + 2] = a * mod.z + (1.0f - a) * col.z;
}

__host__ float3 rgb2hsv_host(const float *c) {
    float cmax = c[0]; int idx = 0;
    if (c[1] > cmax) { cmax = c[1]; idx = 1; }
    if (c[2] > cmax) { cmax = c[2]; idx = 2; }
    const float cmin = fminf(c[0], fminf(c[1], c[2]));
    const float delta = cmax - cmin;
    float h;
    if (delta == 0.0f) h = 0.0f;
    else if (idx == 0) { h = fmodf((c[1] - c[2]) / delta, 6.0f); if (h < 0.0f) h += 6.0f; }
    else if (idx == 1) h = (c[2] - c[0]) / delta + 2.0f;
    else h = (c[0] - c[1]) / delta + 4.0f;
    return make_float3(h / 6.0f, cmax == 0.0f ? 0.0f : delta / cmax, cmax);
}

// set every bitfield byte that holds a cell of the box [cmin, cmax) (cell coordinates) to 255
__global__ void k_force_fill(uint8_t *__restrict__ bitfield, int x0, int y0, int z0, int nx, int ny, int nz, uint32_t cas_offset_bytes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)(nx * ny * nz)) return;
    const uint32_t z = i % nz, y = (i / nz) % ny, x = i / (nz * ny);
    auto expand = [](uint32_t v) {
        v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
        v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v; };
    const uint32_t idx = expand(x0 + x) | (expand(y0 + y) << 1) | (expand(z0 + z) << 2);
    bitfield[cas_offset_bytes + (idx >> 3)] = 255;
}

}  // namespace

// All array arguments are device pointers except the small host-side constant blocks:
// transform[16], rotation[9], scale[3], center[3] (map_data of seal_utils.py:222-236), test_dir[3] or NULL,
// src_bound[6] + map_source[3] or NULL.  bounds: device [nb,2,3]; tris: device [F,3,3].
S3D_API int s3d_seal_bbox_map_to_origin(const float *points, const float *dirs, uint32_t P, const float *h_transform,
                                        const float *h_rotation, const float *h_scale, const float *h_center,
                                        const float *d_bounds, uint32_t nb, const float *d_tris, uint32_t F,
                                        const float *h_test_dir, const float *h_src_bound, const float *h_map_source,
                                        float *out_points, float *out_dirs, uint8_t *mask, void *stream) {
    if (P == 0) return 0;
    if (F * 9 * sizeof(float) > 96 * 1024) return S3D_ENOTSUP;
    MapConsts c;
    for (int i = 0; i < 16; i++) c.transform[i] = h_transform[i];
    for (int i = 0; i < 9; i++) c.rotation[i] = h_rotation[i];
    for (int i = 0; i < 3; i++) { c.scale[i] = h_scale[i]; c.center[i] = h_center[i]; }
    const float d0[3] = {0.4395064455f, 0.617598629942f, 0.652231566745f};  // seal_utils.py:677-679
    for (int i = 0; i < 3; i++) c.test_dir[i] = h_test_dir ? h_test_dir[i] : d0[i];
    c.has_source = (h_src_bound && h_map_source) ? 1 : 0;
    for (int i = 0; i < 3; i++) {
        c.src_lo[i] = c.has_source ? h_src_bound[i] : 0.0f;
        c.src_hi[i] = c.has_source ? h_src_bound[3 + i] : 0.0f;
        c.map_source[i] = c.has_source ? h_map_source[i] : 0.0f;
    }
    const size_t smem = (size_t)F * 9 * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_bbox_map, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_bbox_map<<<div_up(P, 256u), 256, smem, as_stream(stream)>>>(points, dirs, P, c, d_bounds, nb, d_tris, F, out_points, out_dirs, mask);
    S3D_RETURN_LAST();
}

// rgbs [M,3] edited in place on rows with mask != 0 (mask NULL = all rows).  hsv_mod[3] host or NULL;
// rgb_target[3] host or NULL (+ light_offset); d_stats: device float[2] scratch, zeroed here.
S3D_API int s3d_seal_map_color(float *rgbs, const uint8_t *mask, uint32_t M, const float *h_hsv_mod, const float *h_rgb_target,
                               float light_offset, float *d_stats, void *stream) {
    if (M == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (h_hsv_mod) k_modify_hsv<<<div_up(M, 256u), 256, 0, st>>>(rgbs, mask, M, h_hsv_mod[0], h_hsv_mod[1], h_hsv_mod[2]);
    if (h_rgb_target) {
        if (!d_stats) return S3D_EINVAL;
        cudaMemsetAsync(d_stats, 0, 2 * sizeof(float), st);
        { cudaError_t e = v_stats_launch(rgbs, mask, M, d_stats, st); if (e != cudaSuccess) return (int)e; }
        k_modify_rgb<<<div_up(M, 256u), 256, 0, st>>>(rgbs, mask, M, rgb2hsv_host(h_rgb_target), light_offset, d_stats);
    }
    S3D_RETURN_LAST();
}

// SealNeRF/renderer.py:21-66: cells floor(((b + bound) / bound / 2) * H) of [lo, hi) in every cascade's bitfield -> 255.
// cell_lo / cell_hi: host int[3] (already floored and clipped by the caller).
S3D_API int s3d_seal_force_fill_bitfield(uint8_t *bitfield, const int *h_cell_lo, const int *h_cell_hi, uint32_t H,
                                         uint32_t cascade_index, void *stream) {
    const int nx = h_cell_hi[0] - h_cell_lo[0], ny = h_cell_hi[1] - h_cell_lo[1], nz = h_cell_hi[2] - h_cell_lo[2];
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    const uint32_t n = (uint32_t)(nx * ny * nz);
    k_force_fill<<<div_up(n, 256u), 256, 0, as_stream(stream)>>>(bitfield, h_cell_lo[0], h_cell_lo[1], h_cell_lo[2], nx, ny, nz,
                                                                  cascade_index * (H * H * H / 8));
    S3D_RETURN_LAST();
}

static const float kDefaultTestDir[3] = {0.4395064455f, 0.617598629942f, 0.652231566745f};  // seal_utils.py:677-679

// SealBrushMapper.map_to_origin (seal_utils.py:408-453).  h_* = small host constant blocks; d_* = device arrays
// (bounds [nb,2,3], tris [F,3,3], border points [K,3]).  mode: 0 = 'linear', 1 = 'dry'.  dirs are not touched by the
// reference (has_dirs = False), so there is no dirs argument.
S3D_API int s3d_seal_brush_map_to_origin(const float *points, uint32_t P, const float *d_bounds, uint32_t nb, const float *d_tris, uint32_t F,
                                         const float *h_test_dir, const float *h_normal_expand, const float *h_center,
                                         const float *d_border_points, uint32_t K, float attenuation_distance, int mode,
                                         float *out_points, uint8_t *mask, void *stream) {
    if (P == 0) return 0;
    if (mode != 0 && mode != 1) return S3D_ENOTSUP;   // 'ease-in' / 'ease-out' raise NotImplementedError in the reference too
    const size_t smem = ((size_t)F * 9 + (size_t)K * 3) * sizeof(float);
    if (smem > 160 * 1024) return S3D_ENOTSUP;
    BrushConsts c;
    for (int i = 0; i < 3; i++) {
        c.test_dir[i] = h_test_dir ? h_test_dir[i] : kDefaultTestDir[i];
        c.normal_expand[i] = h_normal_expand[i];
        c.center[i] = h_center[i];
    }
    c.att = attenuation_distance; c.mode = mode;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_brush_map, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_brush_map<<<div_up(P, 256u), 256, smem, as_stream(stream)>>>(points, P, c, d_bounds, nb, d_tris, F, d_border_points, K, out_points, mask);
    S3D_RETURN_LAST();
}

// SealAnchorMapper.map_to_origin (seal_utils.py:514-570).  d_flag: device int scratch (zeroed here).
S3D_API int s3d_seal_anchor_map_to_origin(const float *points, uint32_t P, const float *d_bounds, uint32_t nb, const float *d_tris, uint32_t F,
                                          const float *h_test_dir, const float *h_v_anchor, const float *h_v_offset, const float *h_v_h,
                                          float len_h, float radius, const float *h_scale, int *d_flag, float *out_points, uint8_t *mask,
                                          void *stream) {
    if (P == 0) return 0;
    if (!d_flag) return S3D_EINVAL;
    const size_t smem = (size_t)F * 9 * sizeof(float);
    if (smem > 160 * 1024) return S3D_ENOTSUP;
    AnchorConsts c;
    for (int i = 0; i < 3; i++) {
        c.test_dir[i] = h_test_dir ? h_test_dir[i] : kDefaultTestDir[i];
        c.v_anchor[i] = h_v_anchor[i]; c.v_offset[i] = h_v_offset[i]; c.v_h[i] = h_v_h[i]; c.scale[i] = h_scale[i];
    }
    c.len_h = len_h; c.radius = radius;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(d_flag, 0, sizeof(int), st);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_anchor_any, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_anchor_any<<<div_up(P, 256u), 256, smem, st>>>(points, P, c, d_bounds, nb, d_tris, F, d_flag);
    k_anchor_map<<<div_up(P, 256u), 256, 0, st>>>(points, P, c, d_flag, out_points, mask);
    S3D_RETURN_LAST();
}

// texture branch of SealMapper.map_color (seal_utils.py:58-79): rgbs [M,3] edited in place on rows with mask != 0
// (NULL = all rows) from the positions `points` [M,3]; d_image [H,W,3], d_alpha [H,W] device; plane normal / o / w / h
// host float[3]; d_stats device float[2] scratch (mean V of the edited rows, zeroed here).
S3D_API int s3d_seal_map_color_image(float *rgbs, const float *points, const uint8_t *mask, uint32_t M, const float *d_image,
                                     const float *d_alpha, uint32_t H, uint32_t W, const float *h_norm, const float *h_o,
                                     const float *h_w, const float *h_h, float light_offset, float *d_stats, void *stream) {
    if (M == 0) return 0;
    if (!d_stats || H == 0 || W == 0) return S3D_EINVAL;
    ImageConsts c;
    float low2 = 0.0f, loh2 = 0.0f;
    for (int i = 0; i < 3; i++) {
        c.norm[i] = h_norm[i]; c.o[i] = h_o[i]; c.ow[i] = h_w[i] - h_o[i]; c.oh[i] = h_h[i] - h_o[i];
    }
    // len_ow**2 as the reference forms it: (sqrt of the squared norm) squared
    for (int i = 0; i < 3; i++) { low2 += c.ow[i] * c.ow[i]; loh2 += c.oh[i] * c.oh[i]; }
    const float low = sqrtf(low2), loh = sqrtf(loh2);
    c.low2 = low * low; c.loh2 = loh * loh;
    c.light = light_offset; c.H = H; c.W = W;
    cudaStream_t st = as_stream(stream);
    cudaMemsetAsync(d_stats, 0, 2 * sizeof(float), st);
    { cudaError_t e = v_stats_launch(rgbs, mask, M, d_stats, st); if (e != cudaSuccess) return (int)e; }
    k_modify_rgb_image<<<div_up(M, 256u), 256, 0, st>>>(rgbs, points, mask, M, c, d_image, d_alpha, d_stats);
    S3D_RETURN_LAST();
}
