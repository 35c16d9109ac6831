// Occupancy-grid ray marching and alpha compositing for sm_100a.
//
// Replaces raymarching/src/raymarching.cu of the reference (entry points raymarching.h:7-18).
// Design differences (not a port):
//   * march_rays_train is count -> single-CTA exclusive scan -> write.  Sample slots are
//     assigned ray-major and deterministically (the reference allocates them with two global
//     atomics per ray, raymarching.cu:405-406, so its slot order changes run to run).
//   * every float operation whose rounding decides which voxel / lattice point a ray visits is
//     written with explicit round-to-nearest intrinsics (__fmaf_rn / __fmul_rn / __fadd_rn) at
//     exactly the places nvcc contracts the reference expressions, so per-ray sample counts
//     and positions are bit-identical by construction rather than by compiler accident.
//   * the double-precision detour of raymarching.cu:374-376 is taken only when H is not a
//     power of two (it is an exact no-op otherwise).
//   * all launches go to the caller's stream (the reference uses the legacy default stream).
#include "common.cuh"
#include <float.h>

namespace {

constexpr float kSqrt3 = 1.7320508075688772f;
constexpr float kRPi = 0.3183098861837907f;

__host__ __device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits10(x) | (expand_bits10(y) << 1) | (expand_bits10(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact_bits10(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// ---------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------

__global__ void k_near_far(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                           const float *__restrict__ aabb, uint32_t N, float min_near, float *__restrict__ nears,
                           float *__restrict__ fars) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float rdx = __fdiv_rn(1.0f, rays_d[n * 3]), rdy = __fdiv_rn(1.0f, rays_d[n * 3 + 1]),
                rdz = __fdiv_rn(1.0f, rays_d[n * 3 + 2]);
    float tn = __fmul_rn(__fsub_rn(aabb[0], ox), rdx), tf = __fmul_rn(__fsub_rn(aabb[3], ox), rdx);
    if (tn > tf) { const float s = tn; tn = tf; tf = s; }
    float yn = __fmul_rn(__fsub_rn(aabb[1], oy), rdy), yf = __fmul_rn(__fsub_rn(aabb[4], oy), rdy);
    if (yn > yf) { const float s = yn; yn = yf; yf = s; }
    bool miss = (tn > yf) || (yn > tf);
    if (!miss) {
        if (yn > tn) tn = yn;
        if (yf < tf) tf = yf;
        float zn = __fmul_rn(__fsub_rn(aabb[2], oz), rdz), zf = __fmul_rn(__fsub_rn(aabb[5], oz), rdz);
        if (zn > zf) { const float s = zn; zn = zf; zf = s; }
        miss = (tn > zf) || (zn > tf);
        if (!miss) {
            if (zn > tn) tn = zn;
            if (zf < tf) tf = zf;
            if (tn < min_near) tn = min_near;
        }
    }
    nears[n] = miss ? FLT_MAX : tn;
    fars[n] = miss ? FLT_MAX : tf;
}

__global__ void k_sph_from_ray(const float *__restrict__ rays_o, const float *__restrict__ rays_d, float radius,
                               uint32_t N, float *__restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float Bh = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    coords[n * 2] = 2 * atan2f(sqrtf(x * x + z * z), y) * kRPi - 1;
    coords[n * 2 + 1] = atan2f(z, x) * kRPi;
}

__global__ void k_morton3D(const int *__restrict__ coords, uint32_t N, int *__restrict__ indices) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    indices[n] = (int)morton3((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}

__global__ void k_morton3D_invert(const int *__restrict__ indices, uint32_t N, int *__restrict__ coords) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int ind = indices[n];  // arithmetic shifts of a signed value, as the reference does
    coords[n * 3] = (int)compact_bits10((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int)compact_bits10((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int)compact_bits10((uint32_t)(ind >> 2));
}

// one thread packs 32 cells (128 B of density, 4 B of bitfield) with two 128-bit loads in flight
__global__ void k_packbits(const float *__restrict__ grid, uint32_t N, float thresh, uint8_t *__restrict__ bitfield) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;  // index of a 4-byte word of the bitfield
    const uint32_t first = w * 4;
    if (first >= N) return;
    if (first + 4 <= N && ((reinterpret_cast<uintptr_t>(grid) | reinterpret_cast<uintptr_t>(bitfield)) & 15) == 0) {
        const float4 *g = reinterpret_cast<const float4 *>(grid) + (size_t)w * 8;
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = __ldg(g + i);
            bits |= (uint32_t)(v.x > thresh) << (4 * i) | (uint32_t)(v.y > thresh) << (4 * i + 1) |
                    (uint32_t)(v.z > thresh) << (4 * i + 2) | (uint32_t)(v.w > thresh) << (4 * i + 3);
        }
        reinterpret_cast<uint32_t *>(bitfield)[w] = bits;
    } else {
        for (uint32_t n = first; n < N && n < first + 4; n++) {
            uint8_t b = 0;
            for (int i = 0; i < 8; i++) b |= (grid[(size_t)n * 8 + i] > thresh) ? (uint8_t)(1u << i) : 0;
            bitfield[n] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------
// marching
// ---------------------------------------------------------------------------------------

constexpr uint32_t kBitmapWords = 36;                       // 35 x 32 lattice bits + 1 flag word per ray
constexpr uint32_t kBitmapBits = (kBitmapWords - 1) * 32;    // 1120 >= 1024 * bound + 1 lattice steps for bound <= 1

struct MarchCfg {
    float bound, dt_gamma, dt_min, dt_max, rH, H3f, Hf, Hm1f, halfH;
    uint32_t C, H;
    int h_pow2;
    const uint8_t *grid;
};

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, sx, sy, sz;  // s* = 0.5*sign(d*)
};

__device__ __forceinline__ MarchCfg make_cfg(const uint8_t *grid, float bound, float dt_gamma, uint32_t max_steps,
                                             uint32_t C, uint32_t H) {
    MarchCfg c;
    c.bound = bound; c.dt_gamma = dt_gamma; c.C = C; c.H = H; c.grid = grid;
    c.Hf = (float)H; c.Hm1f = (float)(H - 1);
    c.rH = __fdiv_rn(1.0f, c.Hf);
    c.H3f = (float)(H * H * H);
    c.dt_min = __fdiv_rn(__fmul_rn(2.0f, kSqrt3), (float)max_steps);                                  // raymarching.cu:345
    c.dt_max = __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, kSqrt3), (float)(1 << (C - 1))), c.Hf);             // raymarching.cu:346
    c.h_pow2 = (H & (H - 1)) == 0;
    c.halfH = 0.5f * c.Hf;
    return c;
}

__device__ __forceinline__ Ray load_ray(const float *__restrict__ o, const float *__restrict__ d) {
    Ray r;
    r.ox = o[0]; r.oy = o[1]; r.oz = o[2];
    r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
    r.rdx = __fdiv_rn(1.0f, r.dx); r.rdy = __fdiv_rn(1.0f, r.dy); r.rdz = __fdiv_rn(1.0f, r.dz);
    r.sx = copysignf(0.5f, r.dx); r.sy = copysignf(0.5f, r.dy); r.sz = copysignf(0.5f, r.dz);
    return r;
}

__device__ __forceinline__ int frexp_exponent(float v) {
    int e;
    frexpf(v, &e);
    return e;
}

__device__ __forceinline__ int cell_coord(const MarchCfg &c, float x, float mip_rbound) {
    const float v = __fmaf_rn(x, mip_rbound, 1.0f);
    const float f = c.h_pow2 ? __fmul_rn(v, c.halfH) : (float)(0.5 * (double)v * (double)c.H);
    return (int)clampf(f, 0.0f, c.Hm1f);
}

__device__ __forceinline__ float exit_dist(const MarchCfg &c, int n, float half_sign, float mip_bound, float x, float rd) {
    // (((n + 0.5 + 0.5*sign) * rH * 2 - 1) * mip_bound - x) * rd      raymarching.cu:390-392
    const float a = __fadd_rn(__fadd_rn((float)n, 0.5f), half_sign);
    const float u = __fmaf_rn(__fmul_rn(a, c.rH), 2.0f, -1.0f);
    return __fmul_rn(__fmaf_rn(u, mip_bound, -x), rd);
}

__device__ __forceinline__ float step_len(const MarchCfg &c, float t) {
    return clampf(__fmul_rn(t, c.dt_gamma), c.dt_min, c.dt_max);
}

// ---- the step lattice in closed form (constant step) ----------------------------------------------------------------------
// The reference advances a ray with a SERIAL chain of float additions, t_{k+1} = fl(t_k + dt) (raymarching.cu:396-399, :462),
// and every sample sits on that lattice; reproducing the samples bit for bit means reproducing the chain.  With a constant
// step d (dt_gamma = 0) the chain has an exact closed form: inside one binade [2^e, 2^(e+1)) every t is a multiple of
// u = ulp(2^e) and fl(t + d) = t + q with q = round(d / u) * u, an exact addition, so j steps are t + j q exactly (one fmaf:
// the result is a multiple of u below 2^(e+1), hence representable).  Only the step that crosses into the next binade rounds
// on a coarser grid; that one is performed as a real addition.  A rounding tie (d mod u == u / 2, which would make q depend on
// the parity of t / u) falls back to the serial loop.  Used to (a) jump from the ray's start to the occupied bounding box and
// (b) replay empty stretches of the lattice in the write pass; a voxel-by-voxel skip is only ~4.6 steps long and stays serial.
struct Binade {
    float lim, q;
    bool ok;
};
__device__ __forceinline__ Binade binade_of(float t, float d) {
    Binade b;
    const int e = (__float_as_int(t) >> 23) & 0xff;
    b.ok = e > 24 && e < 254;
    const float base = __int_as_float(e << 23), u = __int_as_float((e - 23) << 23);
    b.lim = __int_as_float((e + 1) << 23);
    b.q = __fsub_rn(__fadd_rn(base, d), base);
    const float b1 = __fadd_rn(base, u);
    b.ok = b.ok && b.q > 0.0f && __fsub_rn(__fadd_rn(b1, d), b1) == b.q;
    return b;
}
// largest j >= 0 with t + j q < lim   (lim - t is exact: both are multiples of u below 2^(e+1))
__device__ __forceinline__ float steps_inside(const Binade &b, float t) {
    float jm = floorf(__fdividef(__fsub_rn(b.lim, t), b.q));
    while (jm > 0.0f && __fmaf_rn(jm, b.q, t) >= b.lim) jm -= 1.0f;
    while (__fmaf_rn(jm + 1.0f, b.q, t) < b.lim) jm += 1.0f;
    return jm;
}
// exactly n steps
__device__ __forceinline__ void lattice_advance(float d, float &t, uint32_t n) {
    for (int guard = 0; guard < 40 && n > 0; guard++) {
        const Binade b = binade_of(t, d);
        if (!b.ok) break;
        const float jm = steps_inside(b, t);
        if ((float)n <= jm) { t = __fmaf_rn((float)n, b.q, t); return; }
        t = __fadd_rn(__fmaf_rn(jm, b.q, t), d);
        n -= (uint32_t)jm + 1u;
    }
    for (; n > 0; n--) t = __fadd_rn(t, d);
}
// while (t < target) t += d;   returns the number of steps (0 when t >= target already)
__device__ __forceinline__ uint32_t lattice_advance_to(float d, float &t, float target) {
    uint32_t adv = 0;
    for (int guard = 0; guard < 40 && t < target; guard++) {
        const Binade b = binade_of(t, d);
        if (!b.ok) break;
        const float jm = steps_inside(b, t);
        float nf = fminf(ceilf(__fdividef(__fsub_rn(target, t), b.q)), jm + 1.0f);
        nf = fmaxf(nf, 1.0f);
        while (nf > 1.0f && __fmaf_rn(nf - 1.0f, b.q, t) >= target) nf -= 1.0f;
        while (nf <= jm && __fmaf_rn(nf, b.q, t) < target) nf += 1.0f;
        if (nf <= jm) { t = __fmaf_rn(nf, b.q, t); return adv + (uint32_t)nf; }
        t = __fadd_rn(__fmaf_rn(jm, b.q, t), d);
        adv += (uint32_t)jm + 1u;
    }
    while (t < target) { const float tn = __fadd_rn(t, d); adv++; if (tn == t) break; t = tn; }
    return adv;
}

// Bounding box of the occupied cells of cascade 0 (cell units), by atomicMin / atomicMax: mm = {min x, y, z, max x, y, z},
// initialised to {INT_MAX x 3, -1 x 3}.  One thread per bitfield byte (8 morton-consecutive cells).
__global__ void k_occ_bbox(const uint8_t *__restrict__ grid, uint32_t n_bytes, int *__restrict__ mm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-1, -1, -1};
    if (i < n_bytes) {
        const uint32_t bits = grid[i];
        for (uint32_t b = 0; b < 8; b++)
            if ((bits >> b) & 1u) {
                const uint32_t idx = i * 8 + b;
                const int c[3] = {(int)compact_bits10(idx), (int)compact_bits10(idx >> 1), (int)compact_bits10(idx >> 2)};
#pragma unroll
                for (int d = 0; d < 3; d++) { lo[d] = min(lo[d], c[d]); hi[d] = max(hi[d], c[d]); }
            }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const int l = __reduce_min_sync(0xffffffffu, lo[d]), h = __reduce_max_sync(0xffffffffu, hi[d]);
        if ((threadIdx.x & 31) == 0 && h >= 0) { atomicMin(mm + d, l); atomicMax(mm + 3 + d, h); }
    }
}

__global__ void k_occ_bbox_init(int *__restrict__ mm) {
    if (threadIdx.x < 6) mm[threadIdx.x] = threadIdx.x < 3 ? 0x7fffffff : -1;
    if (threadIdx.x == 6) mm[6] = 0;     // the walk list's counter lives behind the box
}

// Clip a ray to the occupied bounding box, widened by kBoxMargin cells on every side (single cascade, constant step only).
// Everything outside the box is empty, so the reference's walk emits nothing there; jumping to the first lattice point inside
// the widened box and stopping at its far side changes which EMPTY lattice points are visited, not the samples: two voxel
// walks that stand in the same empty voxel jump to the same lattice point (the first one past the voxel's exit), so the walks
// coincide again after at most a visit or two -- and the margin keeps those visits away from any occupied cell.
// Returns false when the ray misses the box (no samples).  k receives the lattice index of the new t.
constexpr int kBoxMargin = 3;
__device__ __forceinline__ bool clip_to_occupied(const MarchCfg &c, const Ray &r, const int *__restrict__ mm, float &t, float &far, uint32_t &k) {
    if (mm[3] < 0) return false;                       // nothing occupied
    const float cell = __fmul_rn(2.0f, c.rH), mb = fminf(1.0f, c.bound);
    float te = -INFINITY, tx = INFINITY;
    const float o[3] = {r.ox, r.oy, r.oz}, rd[3] = {r.rdx, r.rdy, r.rdz};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float lo = ((float)(mm[a] - kBoxMargin) * cell - 1.0f) * mb, hi = ((float)(mm[3 + a] + 1 + kBoxMargin) * cell - 1.0f) * mb;
        const float t1 = (lo - o[a]) * rd[a], t2 = (hi - o[a]) * rd[a];
        te = fmaxf(te, fminf(t1, t2));                 // fminf / fmaxf drop a NaN operand (0 * inf on an axis-aligned ray)
        tx = fminf(tx, fmaxf(t1, t2));
    }
    if (!(te <= tx) || !(tx > t)) return false;
    far = fminf(far, tx);
    if (t < te) k += lattice_advance_to(step_len(c, 0.0f), t, te);
    return true;
}

// One visit of the marching loop (raymarching.cu:359-400).  Occupied: returns true with the sample
// position and dt (caller advances t).  Empty: advances t along the step lattice past the voxel.
template <bool COUNT_LATTICE = false>
__device__ __forceinline__ bool march_visit(const MarchCfg &c, const Ray &r, float &t, float &x, float &y, float &z,
                                            float &dt, uint32_t *lattice_k = nullptr) {
    x = clampf(__fmaf_rn(t, r.dx, r.ox), -c.bound, c.bound);
    y = clampf(__fmaf_rn(t, r.dy, r.oy), -c.bound, c.bound);
    z = clampf(__fmaf_rn(t, r.dz, r.oz), -c.bound, c.bound);
    dt = step_len(c, t);
    int level = 0;
    float mip_bound = fminf(1.0f, c.bound);
    if (c.C > 1) {
        const float cm1 = (float)c.C - 1.0f;
        const int l0 = (int)fminf(cm1, fmaxf(0.0f, (float)frexp_exponent(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))))));
        const int l1 = (int)fminf(cm1, fmaxf(0.0f, (float)frexp_exponent(__fmul_rn(__fmul_rn(dt, c.Hf), 0.5f))));
        level = max(l0, l1);
        mip_bound = fminf(__int_as_float((127 + level) << 23), c.bound);
    }
    const float mip_rbound = __fdiv_rn(1.0f, mip_bound);
    const int nx = cell_coord(c, x, mip_rbound), ny = cell_coord(c, y, mip_rbound), nz = cell_coord(c, z, mip_rbound);
    const uint32_t index = (uint32_t)__fmaf_rn((float)level, c.H3f, (float)morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    if (__ldg(c.grid + (index >> 3)) & (1u << (index & 7))) return true;
    const float tx = exit_dist(c, nx, r.sx, mip_bound, x, r.rdx);
    const float ty = exit_dist(c, ny, r.sy, mip_bound, y, r.rdy);
    const float tz = exit_dist(c, nz, r.sz, mip_bound, z, r.rdz);
    const float tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    do { t = __fadd_rn(t, step_len(c, t)); if (COUNT_LATTICE) (*lattice_k)++; } while (t < tt);
    return false;
}

__device__ __forceinline__ float perturbed_start(const MarchCfg &c, float t0, float noise) {
    return __fmaf_rn(step_len(c, t0), noise, t0);  // raymarching.cu:351
}

// pass 1: count occupied steps per ray -> rays[n] = (n, -, count)
__global__ void __launch_bounds__(128)
k_march_count(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
              float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
              const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
              int *__restrict__ rays, uint32_t *__restrict__ block_sums, uint32_t *__restrict__ bitmap, const int *__restrict__ occ_mm) {
    __shared__ uint32_t s_warp[4];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t num = 0;
    if (n < N) {
        const MarchCfg c = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
        const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        float far = fars[n];
        float t = perturbed_start(c, nears[n], noises ? noises[n] : 0.0f);
        float x, y, z, dt;
        uint32_t k = 0;
        // single cascade + constant step: walk only the part of the ray inside the (widened) bounding box of the occupied cells
        const bool hit = (occ_mm && c.C == 1 && c.dt_gamma == 0.0f) ? clip_to_occupied(c, r, occ_mm, t, far, k) : true;
        if (!hit) far = t;       // no voxel walk at all
        if (bitmap) {
            // also record WHICH points of the ray's step lattice t_{k+1} = t_k + dt(t_k) are samples: the write pass then
            // replays the lattice instead of repeating the voxel walk
            uint32_t *bm = bitmap + (size_t)n * kBitmapWords;
            uint32_t word = 0, widx = 0;
            bool overflow = false;
            while (t < far && num < max_steps) {
                const uint32_t k0 = k;
                if (march_visit<true>(c, r, t, x, y, z, dt, &k)) {
                    if (k0 < kBitmapBits) {
                        const uint32_t wi = k0 >> 5;
                        while (widx < wi) { bm[widx++] = word; word = 0; }
                        word |= 1u << (k0 & 31);
                    } else overflow = true;
                    num++; t = __fadd_rn(t, dt); k++;
                }
            }
            while (widx < kBitmapWords - 1) { bm[widx++] = word; word = 0; }
            bm[kBitmapWords - 1] = overflow ? 0xFFFFFFFFu : 0u;   // last word = "lattice too long, re-walk this ray"
        } else {
            while (t < far && num < max_steps) {
                if (march_visit(c, r, t, x, y, z, dt)) { num++; t = __fadd_rn(t, dt); }
            }
        }
        rays[(size_t)n * 3] = (int)n;
        rays[(size_t)n * 3 + 2] = (int)num;
    }
    // per-CTA sample count for the two-level exclusive scan
    uint32_t v = num;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
}

// ---- pass 1 in compacted form (single cascade, constant step) -----------------------------------------------------------
// Most rays of a batch miss the (widened) bounding box of the occupied cells -- about two thirds on the Lego-shaped scene -- and
// inside k_march_count their lanes idle while the rest of the warp walks voxels (the pass is instruction-bound, so idle lanes
// are lost issue slots).  Here a first kernel clips every ray (cheap, no divergence) and COMPACTS the rays that enter the box
// into a work list with a warp ballot + one atomic per warp; the voxel walk then runs on the list, densely packed.  Results
// are written by ray id, so the (non-deterministic) list order changes nothing; block sums for the scan come from a third
// tiny kernel.
struct WalkItem { uint32_t n; float t, far; uint32_t k; };

__global__ void __launch_bounds__(128)
k_march_clip(const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
             const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises, const int *__restrict__ occ_mm,
             int *__restrict__ rays, WalkItem *__restrict__ items, uint32_t *__restrict__ n_items) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    bool hit = false;
    WalkItem it{n, 0.0f, 0.0f, 0u};
    if (n < N) {
        const MarchCfg c = make_cfg(nullptr, bound, 0.0f, max_steps, C, H);
        const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        it.far = fars[n];
        it.t = perturbed_start(c, nears[n], noises ? noises[n] : 0.0f);
        hit = clip_to_occupied(c, r, occ_mm, it.t, it.far, it.k) && it.t < it.far;
        rays[(size_t)n * 3] = (int)n;
        rays[(size_t)n * 3 + 2] = 0;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(n_items, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit) items[base + __popc(m & ((1u << lane) - 1u))] = it;
}

__global__ void __launch_bounds__(128)
k_march_walk(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid, float bound, uint32_t max_steps,
             uint32_t C, uint32_t H, const WalkItem *__restrict__ items, const uint32_t *__restrict__ n_items, int *__restrict__ rays,
             uint32_t *__restrict__ bitmap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_items) return;
    const WalkItem it = items[i];
    const uint32_t n = it.n;
    const MarchCfg c = make_cfg(grid, bound, 0.0f, max_steps, C, H);
    const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
    float t = it.t, x, y, z, dt;
    const float far = it.far;
    uint32_t k = it.k, num = 0;
    if (bitmap) {
        uint32_t *bm = bitmap + (size_t)n * kBitmapWords;
        uint32_t word = 0, widx = 0;
        bool overflow = false;
        while (t < far && num < max_steps) {
            const uint32_t k0 = k;
            if (march_visit<true>(c, r, t, x, y, z, dt, &k)) {
                if (k0 < kBitmapBits) {
                    const uint32_t wi = k0 >> 5;
                    while (widx < wi) { bm[widx++] = word; word = 0; }
                    word |= 1u << (k0 & 31);
                } else overflow = true;
                num++; t = __fadd_rn(t, dt); k++;
            }
        }
        while (widx < kBitmapWords - 1) { bm[widx++] = word; word = 0; }
        bm[kBitmapWords - 1] = overflow ? 0xFFFFFFFFu : 0u;
    } else {
        while (t < far && num < max_steps) {
            if (march_visit(c, r, t, x, y, z, dt)) { num++; t = __fadd_rn(t, dt); }
        }
    }
    rays[(size_t)n * 3 + 2] = (int)num;
}

__global__ void __launch_bounds__(128) k_march_block_sums(const int *__restrict__ rays, uint32_t N, uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t s_warp[4];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = n < N ? (uint32_t)rays[(size_t)n * 3 + 2] : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
}

// single-CTA exclusive scan of the per-CTA sums (N/128 values) in place; counter += (sum, N)
__global__ void __launch_bounds__(1024) k_march_scan(uint32_t *__restrict__ block_sums, uint32_t nb, uint32_t N, int *__restrict__ counter) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_base;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t chunk = div_up(nb, nthr);
    const uint32_t lo = min(tid * chunk, nb), hi = min(lo + chunk, nb);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += block_sums[i];
    uint32_t incl = sum;
    const uint32_t lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    if (tid == 0) s_base = (uint32_t)counter[0];
    __syncthreads();
    if (wid == 0) {
        uint32_t w = (lane < (nthr >> 5)) ? warp_tot[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += v;
        }
        warp_tot[lane] = wi - w;  // exclusive
        if (lane == 31) {
            counter[0] = (int)(s_base + wi);
            counter[1] = counter[1] + (int)N;
        }
    }
    __syncthreads();
    uint32_t run = s_base + warp_tot[wid] + (incl - sum);
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t c = block_sums[i];
        block_sums[i] = run;
        run += c;
    }
}

// rays[:,1] = CTA offset + exclusive scan of the CTA's 128 counts
__global__ void __launch_bounds__(128) k_march_offsets(int *__restrict__ rays, uint32_t N, const uint32_t *__restrict__ block_offs) {
    __shared__ uint32_t s_warp[4];
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t c = n < N ? (uint32_t)rays[(size_t)n * 3 + 2] : 0;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = block_offs[blockIdx.x];
    for (uint32_t w = 0; w < wid; w++) base += s_warp[w];
    if (n < N) rays[(size_t)n * 3 + 1] = (int)(base + incl - c);
}

// pass 2: re-march and write the samples of every ray that fits (raymarching.cu:415-479)
__global__ void __launch_bounds__(128)
k_march_write(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
              float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
              const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
              const int *__restrict__ rays, float *__restrict__ xyzs, float *__restrict__ dirs,
              float *__restrict__ deltas) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[(size_t)n * 3 + 1], num = (uint32_t)rays[(size_t)n * 3 + 2];
    if (num == 0 || off + num > M) return;
    const MarchCfg c = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
    const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
    const float far = fars[n];
    float t = perturbed_start(c, nears[n], noises ? noises[n] : 0.0f);
    float last_t = t, x, y, z, dt;
    float *px = xyzs + (size_t)off * 3, *pd = dirs + (size_t)off * 3, *pl = deltas + (size_t)off * 2;
    uint32_t step = 0;
    while (t < far && step < num) {
        if (march_visit(c, r, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, dt);
            *reinterpret_cast<float2 *>(pl) = make_float2(dt, __fsub_rn(t, last_t));
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

// pass 2 (bitmap variant): replay the step lattice and emit the recorded samples -- no voxel lookups, no DDA
__global__ void __launch_bounds__(128)
k_march_write_bitmap(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
                     float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                     const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
                     const int *__restrict__ rays, const uint32_t *__restrict__ bitmap, float *__restrict__ xyzs,
                     float *__restrict__ dirs, float *__restrict__ deltas) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[(size_t)n * 3 + 1], num = (uint32_t)rays[(size_t)n * 3 + 2];
    if (num == 0 || off + num > M) return;
    const MarchCfg c = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
    const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
    const float far = fars[n];
    float t = perturbed_start(c, nears[n], noises ? noises[n] : 0.0f);
    float last_t = t, x, y, z, dt;
    float *px = xyzs + (size_t)off * 3, *pd = dirs + (size_t)off * 3, *pl = deltas + (size_t)off * 2;
    const uint32_t *bm = bitmap + (size_t)n * kBitmapWords;
    uint32_t step = 0;
    if (bm[kBitmapWords - 1] != 0u) {   // lattice longer than the bitmap: walk the voxels again for this ray
        while (t < far && step < num) {
            if (march_visit(c, r, t, x, y, z, dt)) {
                px[0] = x; px[1] = y; px[2] = z; pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                t = __fadd_rn(t, dt);
                *reinterpret_cast<float2 *>(pl) = make_float2(dt, __fsub_rn(t, last_t));
                last_t = t; px += 3; pd += 3; pl += 2; step++;
            }
        }
        return;
    }
    const bool const_step = c.dt_gamma == 0.0f;
    const float d0 = step_len(c, 0.0f);
    for (uint32_t w = 0; w < kBitmapWords - 1 && step < num; w++) {
        const uint32_t bits = __ldg(bm + w);
        if (bits == 0u && const_step) { lattice_advance(d0, t, 32u); continue; }   // 32 empty lattice points in closed form
#pragma unroll 4
        for (uint32_t b = 0; b < 32; b++) {
            dt = step_len(c, t);
            if ((bits >> b) & 1u) {
                x = clampf(__fmaf_rn(t, r.dx, r.ox), -c.bound, c.bound);
                y = clampf(__fmaf_rn(t, r.dy, r.oy), -c.bound, c.bound);
                z = clampf(__fmaf_rn(t, r.dz, r.oz), -c.bound, c.bound);
                px[0] = x; px[1] = y; px[2] = z; pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
                const float tn = __fadd_rn(t, dt);
                *reinterpret_cast<float2 *>(pl) = make_float2(dt, __fsub_rn(tn, last_t));
                last_t = tn; px += 3; pd += 3; pl += 2; step++;
            }
            t = __fadd_rn(t, dt);
        }
    }
}

// pass 2, constant step (dt_gamma = 0): ONE THREAD PER SAMPLE.  The ray-per-thread replay above writes each ray's samples with
// 4-byte stores at a stride of one ray per lane (8 store instructions per sample, a sector per lane each); here consecutive
// threads own consecutive samples, so xyzs / dirs / deltas are written fully coalesced.  A thread finds its ray by binary
// search over the (monotone) sample offsets, its lattice index as the (j+1)-th set bit of the ray's bitmap, and its t -- the
// value the reference's serial chain of additions reaches after k steps -- in closed form (lattice_advance).  Rays whose
// lattice outgrew the bitmap are left to k_march_write_overflow.
__global__ void __launch_bounds__(256)
k_march_write_samples(const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, uint32_t max_steps, uint32_t N,
                      uint32_t C, uint32_t H, uint32_t M, const float *__restrict__ nears, const float *__restrict__ noises,
                      const int *__restrict__ rays, const int *__restrict__ counter_total, const uint32_t *__restrict__ bitmap,
                      float *__restrict__ xyzs, float *__restrict__ dirs, float *__restrict__ deltas) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    // counter[0] after the scan = end of this call's samples; they start at the first ray's offset (the counter's value
    // before the call, 0 when the caller zeroed it like nerf/renderer.py:267 does)
    const uint32_t end = (uint32_t)*counter_total;
    if (s >= M || s >= end || s < (uint32_t)__ldg(rays + 1)) return;
    // last ray with offset <= s
    uint32_t lo = 0, hi = N;            // invariant: off[lo] <= s, off[hi] > s or hi = N
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if ((uint32_t)__ldg(rays + (size_t)mid * 3 + 1) <= s) lo = mid; else hi = mid;
    }
    const uint32_t n = lo, off = (uint32_t)rays[(size_t)n * 3 + 1], num = (uint32_t)rays[(size_t)n * 3 + 2];
    if (off + num > M) return;                                   // dropped ray (budget overflow): its slots stay zero
    const uint32_t *bm = bitmap + (size_t)n * kBitmapWords;      // (a ray with samples was walked: its bitmap is initialised)
    if (__ldg(bm + kBitmapWords - 1) != 0u) return;               // re-walked by k_march_write_overflow
    const uint32_t j = s - off;
    // lattice indices of sample j (k) and of the sample before it (kp)
    uint32_t cum = 0, k = 0, kp = 0;
    bool have_prev = false;
    for (uint32_t w = 0; w < kBitmapWords - 1; w++) {
        const uint32_t bits = __ldg(bm + w), c = __popc(bits);
        if (j > 0 && !have_prev && cum + c >= j) { kp = w * 32 + __fns(bits, 0, j - cum); have_prev = true; }   // the j-th set bit
        if (cum + c > j) { k = w * 32 + __fns(bits, 0, j - cum + 1); break; }
        cum += c;
    }
    const MarchCfg c = make_cfg(nullptr, bound, 0.0f, max_steps, C, H);
    const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
    const float d0 = step_len(c, 0.0f);
    const float t0 = perturbed_start(c, nears[n], noises ? noises[n] : 0.0f);
    float t = t0;
    lattice_advance(d0, t, k);
    float last_t = t0;
    if (j > 0) lattice_advance(d0, last_t, kp + 1);
    const float x = clampf(__fmaf_rn(t, r.dx, r.ox), -c.bound, c.bound);
    const float y = clampf(__fmaf_rn(t, r.dy, r.oy), -c.bound, c.bound);
    const float z = clampf(__fmaf_rn(t, r.dz, r.oz), -c.bound, c.bound);
    const float tn = __fadd_rn(t, d0);
    xyzs[(size_t)s * 3] = x; xyzs[(size_t)s * 3 + 1] = y; xyzs[(size_t)s * 3 + 2] = z;
    dirs[(size_t)s * 3] = r.dx; dirs[(size_t)s * 3 + 1] = r.dy; dirs[(size_t)s * 3 + 2] = r.dz;
    reinterpret_cast<float2 *>(deltas)[s] = make_float2(d0, __fsub_rn(tn, last_t));
}

// rays whose step lattice is longer than the bitmap (bound > 1 scenes): walk the voxels again, one thread per ray
__global__ void __launch_bounds__(128)
k_march_write_overflow(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const uint8_t *__restrict__ grid,
                       float bound, float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                       const float *__restrict__ nears, const float *__restrict__ fars, const float *__restrict__ noises,
                       const int *__restrict__ rays, const uint32_t *__restrict__ bitmap, float *__restrict__ xyzs,
                       float *__restrict__ dirs, float *__restrict__ deltas) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t off = (uint32_t)rays[(size_t)n * 3 + 1], num = (uint32_t)rays[(size_t)n * 3 + 2];
    if (num == 0 || off + num > M) return;        // rays without samples may never have been walked: their bitmap is not read
    if (bitmap[(size_t)n * kBitmapWords + kBitmapWords - 1] == 0u) return;
    const MarchCfg c = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
    const Ray r = load_ray(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
    const float far = fars[n];
    float t = perturbed_start(c, nears[n], noises ? noises[n] : 0.0f);
    float last_t = t, x, y, z, dt;
    float *px = xyzs + (size_t)off * 3, *pd = dirs + (size_t)off * 3, *pl = deltas + (size_t)off * 2;
    uint32_t step = 0;
    while (t < far && step < num) {
        if (march_visit(c, r, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z; pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, dt);
            *reinterpret_cast<float2 *>(pl) = make_float2(dt, __fsub_rn(t, last_t));
            last_t = t; px += 3; pd += 3; pl += 2; step++;
        }
    }
}

// inference marcher: n_step samples per alive ray at a fixed stride (raymarching.cu:701-805)
__global__ void __launch_bounds__(128)
k_march_rays(uint32_t n_alive, uint32_t n_step, const int *__restrict__ rays_alive, const float *__restrict__ rays_t,
             const float *__restrict__ rays_o, const float *__restrict__ rays_d, float bound, float dt_gamma,
             uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *__restrict__ grid,
             const float *__restrict__ nears, const float *__restrict__ fars, float *__restrict__ xyzs,
             float *__restrict__ dirs, float *__restrict__ deltas, const float *__restrict__ noises) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const MarchCfg c = make_cfg(grid, bound, dt_gamma, max_steps, C, H);
    const Ray r = load_ray(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
    const float far = fars[index];
    float t = perturbed_start(c, rays_t[index], noises ? noises[n] : 0.0f);
    float last_t = t, x, y, z, dt;
    float *px = xyzs + (size_t)n * n_step * 3, *pd = dirs + (size_t)n * n_step * 3, *pl = deltas + (size_t)n * n_step * 2;
    uint32_t step = 0;
    while (t < far && step < n_step) {
        if (march_visit(c, r, t, x, y, z, dt)) {
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            t = __fadd_rn(t, dt);
            pl[0] = dt; pl[1] = __fsub_rn(t, last_t);
            last_t = t;
            px += 3; pd += 3; pl += 2; step++;
        }
    }
}

// ---------------------------------------------------------------------------------------
// compositing
// ---------------------------------------------------------------------------------------

// Training compositor: ONE WARP PER RAY.  Lanes take consecutive samples (coalesced 4/12/8-byte streams); the
// transmittance T_before(i) = prod_{j<i} (1 - alpha_j) is a multiplicative warp scan carried across 32-sample chunks.
// The reference is one thread per ray walking its samples serially (raymarching.cu:501-577): same arithmetic per
// sample, products associated as a tree instead of a chain.  Early stop: a sample is accumulated iff the
// transmittance before it is still >= T_thresh (the reference breaks AFTER the sample that drives T below it).
__device__ __forceinline__ float warp_excl_prod(float v, uint32_t lane, float &total) {
    float incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl *= o;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    const float ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 1.0f : ex;
}
__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ float warp_incl_sum(float v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += o;
    }
    return v;
}

__global__ void __launch_bounds__(256)
k_composite_train_fwd(const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
                      const int *__restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                      float *__restrict__ weights_sum, float *__restrict__ depth, float *__restrict__ image) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[(size_t)n * 3], offset = (uint32_t)rays[(size_t)n * 3 + 1],
                   num = (uint32_t)rays[(size_t)n * 3 + 2];
    float r = 0, g = 0, b = 0, ws = 0, d = 0, T = 1.0f, t0 = 0.0f;
    if (num != 0 && offset + num <= M) {
        const float *s = sigmas + offset, *c = rgbs + (size_t)offset * 3;
        const float2 *dl = reinterpret_cast<const float2 *>(deltas) + offset;
        for (uint32_t base = 0; base < num; base += 32) {
            const uint32_t i = base + lane;
            const bool in = i < num;
            float2 de = make_float2(0.f, 0.f);
            float sg = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
            if (in) { de = __ldg(dl + i); sg = __ldg(s + i); c0 = __ldg(c + i * 3); c1 = __ldg(c + i * 3 + 1); c2 = __ldg(c + i * 3 + 2); }
            const float alpha = in ? 1.0f - __expf(-sg * de.x) : 0.0f;
            float tot;
            const float Tb = T * warp_excl_prod(1.0f - alpha, lane, tot);   // transmittance before this sample
            const float tt = t0 + warp_incl_sum(de.y, lane);               // accumulated "real delta" up to this sample
            const bool take = in && (Tb >= T_thresh || i == 0);
            const float w = take ? alpha * Tb : 0.0f;
            r += w * c0; g += w * c1; b += w * c2; ws += w; d += w * tt;
            T *= tot;
            t0 = __shfl_sync(0xffffffffu, tt, 31);
            if (T < T_thresh) break;   // warp-uniform: every later sample has T_before < T_thresh
        }
        r = warp_sum_all(r); g = warp_sum_all(g); b = warp_sum_all(b); ws = warp_sum_all(ws); d = warp_sum_all(d);
    }
    if (lane == 0) {
        weights_sum[index] = ws;
        depth[index] = d;
        image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
    }
}

// Backward (raymarching.cu:602-682): grad_rgb_i = g * w_i ; grad_sigma_i = delta_i * (sum_c g_c (T_after_i c_i - (C_final - C_incl_i))
// + g_ws (1 - ws_final)), written for the accumulated prefix only (later samples keep the caller's zeros).
__global__ void __launch_bounds__(256)
k_composite_train_bwd(const float *__restrict__ grad_ws, const float *__restrict__ grad_image,
                      const float *__restrict__ sigmas, const float *__restrict__ rgbs, const float *__restrict__ deltas,
                      const int *__restrict__ rays, const float *__restrict__ weights_sum,
                      const float *__restrict__ image, uint32_t M, uint32_t N, float T_thresh,
                      float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const uint32_t index = (uint32_t)rays[(size_t)n * 3], offset = (uint32_t)rays[(size_t)n * 3 + 1],
                   num = (uint32_t)rays[(size_t)n * 3 + 2];
    if (num == 0 || offset + num > M) return;
    const float gws = grad_ws[index];
    const float g0 = grad_image[(size_t)index * 3], g1 = grad_image[(size_t)index * 3 + 1], g2 = grad_image[(size_t)index * 3 + 2];
    const float rf = image[(size_t)index * 3], gf = image[(size_t)index * 3 + 1], bf = image[(size_t)index * 3 + 2];
    const float tail = gws * (1 - weights_sum[index]);
    const float *s = sigmas + offset, *c = rgbs + (size_t)offset * 3;
    const float2 *dl = reinterpret_cast<const float2 *>(deltas) + offset;
    float *gs = grad_sigmas + offset, *gc = grad_rgbs + (size_t)offset * 3;
    float T = 1.0f, r0 = 0, gr0 = 0, b0 = 0;
    for (uint32_t base = 0; base < num; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < num;
        float2 de = make_float2(0.f, 0.f);
        float sg = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (in) { de = __ldg(dl + i); sg = __ldg(s + i); c0 = __ldg(c + i * 3); c1 = __ldg(c + i * 3 + 1); c2 = __ldg(c + i * 3 + 2); }
        const float alpha = in ? 1.0f - __expf(-sg * de.x) : 0.0f;
        float tot;
        const float Tb = T * warp_excl_prod(1.0f - alpha, lane, tot);
        const bool take = in && (Tb >= T_thresh || i == 0);
        const float w = take ? alpha * Tb : 0.0f;
        const float Ta = Tb * (1.0f - alpha);
        const float r = r0 + warp_incl_sum(w * c0, lane), g = gr0 + warp_incl_sum(w * c1, lane), b = b0 + warp_incl_sum(w * c2, lane);
        if (take) {
            gc[i * 3] = g0 * w; gc[i * 3 + 1] = g1 * w; gc[i * 3 + 2] = g2 * w;
            gs[i] = de.x * (g0 * (Ta * c0 - (rf - r)) + g1 * (Ta * c1 - (gf - g)) + g2 * (Ta * c2 - (bf - b)) + tail);
        }
        T *= tot;
        r0 = __shfl_sync(0xffffffffu, r, 31); gr0 = __shfl_sync(0xffffffffu, g, 31); b0 = __shfl_sync(0xffffffffu, b, 31);
        if (T < T_thresh) break;
    }
}

// One warp per ray, the whole per-ray part of a distillation / fine-tuning step in ONE launch:
//   teacher forward scan (optional: sig_t / rgb_t on the student's samples -> target image + bg and target depth; otherwise the
//   targets are read from image_t / depth_t), student forward scan, the photometric loss of nerf/utils.py:484-489, 530
//   (mean_rays mean_c (rgb - gt)^2 + mean |depth - gt|) with its gradient, and the student backward scan
//   (raymarching.cu:602-682) -- what the separate path does with composite_fwd x 2, a background add, finetune_loss, two
//   gradient scalings and composite_bwd.  Arithmetic per sample is that of k_composite_train_fwd / _bwd and k_finetune_loss.
//   grad_sigmas / grad_rgbs must be zero-initialised by the caller (samples past a ray's accumulated prefix keep the zeros).
struct RayScan { float r, g, b, ws, d; };

// forward scan of one (S only) or two (S and T on the same samples) fields over the samples of a ray; the deltas are read once
template <bool TWO>
__device__ __forceinline__ void ray_forward(const float *__restrict__ s, const float *__restrict__ c, const float *__restrict__ s2, const float *__restrict__ c2,
                                            const float2 *__restrict__ dl, uint32_t num, uint32_t lane, float T_thresh, RayScan &A, RayScan &B) {
    float r = 0, g = 0, b = 0, ws = 0, d = 0, T = 1.0f, t0 = 0.0f;
    float r2 = 0, g2 = 0, b2 = 0, ws2 = 0, d2 = 0, T2 = 1.0f;
    for (uint32_t base = 0; base < num; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < num;
        float2 de = make_float2(0.f, 0.f);
        float sg = 0.f, c0 = 0.f, c1 = 0.f, cc2 = 0.f, sh = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f;
        if (in) {
            de = __ldg(dl + i); sg = __ldg(s + i); c0 = __ldg(c + i * 3); c1 = __ldg(c + i * 3 + 1); cc2 = __ldg(c + i * 3 + 2);
            if (TWO) { sh = __ldg(s2 + i); e0 = __ldg(c2 + i * 3); e1 = __ldg(c2 + i * 3 + 1); e2 = __ldg(c2 + i * 3 + 2); }
        }
        const float tt = t0 + warp_incl_sum(de.y, lane);
        t0 = __shfl_sync(0xffffffffu, tt, 31);
        const bool liveA = !(T < T_thresh), liveB = TWO && !(T2 < T_thresh);      // warp-uniform
        if (liveA) {
            const float alpha = in ? 1.0f - __expf(-sg * de.x) : 0.0f;
            float tot;
            const float Tb = T * warp_excl_prod(1.0f - alpha, lane, tot);
            const bool take = in && (Tb >= T_thresh || i == 0);
            const float w = take ? alpha * Tb : 0.0f;
            r += w * c0; g += w * c1; b += w * cc2; ws += w; d += w * tt;
            T *= tot;
        }
        if (liveB) {
            const float alpha = in ? 1.0f - __expf(-sh * de.x) : 0.0f;
            float tot;
            const float Tb = T2 * warp_excl_prod(1.0f - alpha, lane, tot);
            const bool take = in && (Tb >= T_thresh || i == 0);
            const float w = take ? alpha * Tb : 0.0f;
            r2 += w * e0; g2 += w * e1; b2 += w * e2; ws2 += w; d2 += w * tt;
            T2 *= tot;
        }
        if (T < T_thresh && (!TWO || T2 < T_thresh)) break;
    }
    A.r = warp_sum_all(r); A.g = warp_sum_all(g); A.b = warp_sum_all(b); A.ws = warp_sum_all(ws); A.d = warp_sum_all(d);
    if (TWO) { B.r = warp_sum_all(r2); B.g = warp_sum_all(g2); B.b = warp_sum_all(b2); B.ws = warp_sum_all(ws2); B.d = warp_sum_all(d2); }
}

__global__ void __launch_bounds__(256)
k_distill_rays(const float *__restrict__ sig_t, const float *__restrict__ rgb_t, const float *__restrict__ image_t, const float *__restrict__ depth_t,
               const float *__restrict__ sig_s, const float *__restrict__ rgb_s, const float *__restrict__ deltas, const int *__restrict__ rays,
               uint32_t M, uint32_t N, float T_thresh, float bg, float scale, const float *__restrict__ scale_dev, float *__restrict__ loss,
               float *__restrict__ grad_sigmas, float *__restrict__ grad_rgbs) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float a0 = 0.0f, a1 = 0.0f;
    if (n < N) {
        const uint32_t index = (uint32_t)rays[(size_t)n * 3], offset = (uint32_t)rays[(size_t)n * 3 + 1], num = (uint32_t)rays[(size_t)n * 3 + 2];
        const bool live = num != 0 && offset + num <= M;
        const float2 *dl = reinterpret_cast<const float2 *>(deltas) + offset;
        RayScan S{0, 0, 0, 0, 0}, T{0, 0, 0, 0, 0};
        if (live) {
            if (sig_t) ray_forward<true>(sig_s + offset, rgb_s + (size_t)offset * 3, sig_t + offset, rgb_t + (size_t)offset * 3, dl, num, lane, T_thresh, S, T);
            else ray_forward<false>(sig_s + offset, rgb_s + (size_t)offset * 3, nullptr, nullptr, dl, num, lane, T_thresh, S, T);
        }
        float t0, t1, t2, td = 0.0f;
        bool has_depth = true;
        if (sig_t) {
            const float back_t = (1.0f - T.ws) * bg;
            t0 = T.r + back_t; t1 = T.g + back_t; t2 = T.b + back_t; td = T.d;
        } else {
            t0 = image_t[(size_t)index * 3]; t1 = image_t[(size_t)index * 3 + 1]; t2 = image_t[(size_t)index * 3 + 2];
            has_depth = depth_t != nullptr;
            if (has_depth) td = depth_t[index];
        }
        const float sc = scale_dev ? *scale_dev : scale;
        const float inv_n = 1.0f / (float)N, k = 2.0f / (3.0f * (float)N);
        const float back = (1.0f - S.ws) * bg;
        const float d0 = S.r + back - t0, d1 = S.g + back - t1, d2 = S.b + back - t2;
        a0 = (d0 * d0 + d1 * d1 + d2 * d2) * (inv_n / 3.0f);
        if (has_depth) a1 = fabsf(S.d - td) * inv_n;
        const float g0 = k * d0 * sc, g1 = k * d1 * sc, g2 = k * d2 * sc;
        const float gws = -bg * (g0 + g1 + g2);
        if (live) {
            const float tail = gws * (1 - S.ws);
            const float *s = sig_s + offset, *c = rgb_s + (size_t)offset * 3;
            float *gs = grad_sigmas + offset, *gc = grad_rgbs + (size_t)offset * 3;
            float Tr = 1.0f, r0 = 0, gr0 = 0, b0 = 0;
            for (uint32_t base = 0; base < num; base += 32) {
                const uint32_t i = base + lane;
                const bool in = i < num;
                float2 de = make_float2(0.f, 0.f);
                float sg = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
                if (in) { de = __ldg(dl + i); sg = __ldg(s + i); c0 = __ldg(c + i * 3); c1 = __ldg(c + i * 3 + 1); c2 = __ldg(c + i * 3 + 2); }
                const float alpha = in ? 1.0f - __expf(-sg * de.x) : 0.0f;
                float tot;
                const float Tb = Tr * warp_excl_prod(1.0f - alpha, lane, tot);
                const bool take = in && (Tb >= T_thresh || i == 0);
                const float w = take ? alpha * Tb : 0.0f;
                const float Ta = Tb * (1.0f - alpha);
                const float r = r0 + warp_incl_sum(w * c0, lane), g = gr0 + warp_incl_sum(w * c1, lane), b = b0 + warp_incl_sum(w * c2, lane);
                if (take) {
                    gc[i * 3] = g0 * w; gc[i * 3 + 1] = g1 * w; gc[i * 3 + 2] = g2 * w;
                    gs[i] = de.x * (g0 * (Ta * c0 - (S.r - r)) + g1 * (Ta * c1 - (S.g - g)) + g2 * (Ta * c2 - (S.b - b)) + tail);
                }
                Tr *= tot;
                r0 = __shfl_sync(0xffffffffu, r, 31); gr0 = __shfl_sync(0xffffffffu, g, 31); b0 = __shfl_sync(0xffffffffu, b, 31);
                if (Tr < T_thresh) break;
            }
        }
    }
    // loss terms: one value per ray (lane 0 of its warp) -> CTA sum -> one atomic pair per CTA
    __shared__ float s_a0[8], s_a1[8];
    if (lane == 0) { s_a0[threadIdx.x >> 5] = a0; s_a1[threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float x = 0, y = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { x += s_a0[w]; y += s_a1[w]; }
        atomicAdd(loss, x); atomicAdd(loss + 1, y);
    }
}

__global__ void __launch_bounds__(128)
k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int *__restrict__ rays_alive,
                 float *__restrict__ rays_t, const float *__restrict__ sigmas, const float *__restrict__ rgbs,
                 const float *__restrict__ deltas, float *__restrict__ weights_sum, float *__restrict__ depth,
                 float *__restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int index = rays_alive[n];
    const float *s = sigmas + (size_t)n * n_step, *c = rgbs + (size_t)n * n_step * 3, *dl = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index], ws = weights_sum[index], d = depth[index];
    float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];
    uint32_t step = 0;
    while (step < n_step) {
        if (dl[0] == 0) break;  // zero-filled rows terminate the ray (raymarching.cu:858)
        const float alpha = 1.0f - __expf(-s[0] * dl[0]);
        const float T = 1 - ws;
        const float w = alpha * T;
        ws += w;
        t += dl[1];
        d = fmaf(w, t, d);
        r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
        if (T < T_thresh) break;
        s++; c += 3; dl += 2; step++;
    }
    if (step < n_step) rays_alive[n] = -1; else rays_t[index] = t;
    weights_sum[index] = ws; depth[index] = d;
    image[(size_t)index * 3] = r; image[(size_t)index * 3 + 1] = g; image[(size_t)index * 3 + 2] = b;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------

S3D_API int s3d_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                                   float min_near, float *nears, float *fars, void *stream) {
    if (N == 0) return 0;
    k_near_far<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    S3D_RETURN_LAST();
}

S3D_API int s3d_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords,
                             void *stream) {
    if (N == 0) return 0;
    k_sph_from_ray<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(rays_o, rays_d, radius, N, coords);
    S3D_RETURN_LAST();
}

S3D_API int s3d_morton3D(const int *coords, uint32_t N, int *indices, void *stream) {
    if (N == 0) return 0;
    k_morton3D<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(coords, N, indices);
    S3D_RETURN_LAST();
}

S3D_API int s3d_morton3D_invert(const int *indices, uint32_t N, int *coords, void *stream) {
    if (N == 0) return 0;
    k_morton3D_invert<<<div_up(N, 256u), 256, 0, as_stream(stream)>>>(indices, N, coords);
    S3D_RETURN_LAST();
}

S3D_API int s3d_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream) {
    if (N == 0) return 0;
    const uint32_t words = div_up(N, 4u);
    k_packbits<<<div_up(words, 256u), 256, 0, as_stream(stream)>>>(grid, N, density_thresh, bitfield);
    S3D_RETURN_LAST();
}

namespace {
int g_march_clip = 1;   // s3d_march_set_clip: 0 walks every ray from its near point like the reference (tests compare the two)
int march_count_scan(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t N, uint32_t C, uint32_t H, const float *nears, const float *fars, const float *noises, int *rays,
                     int *counter, cudaStream_t st, uint32_t *bitmap = nullptr) {
    const uint32_t nb = div_up(N, 128u);
    uint32_t *block_sums = nullptr;
    cudaError_t e = scratch_alloc((void **)&block_sums, ((size_t)nb + 8) * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (g_march_clip && C == 1 && dt_gamma == 0.0f && H <= 1024) {
        // occupied bounding box of the bitfield (stateless: recomputed per call, 262 KB read), clip + compact, walk the list
        int *occ_mm = (int *)(block_sums + nb);               // 6 ints + the list counter in the 8 spare words
        uint32_t *n_items = (uint32_t *)(occ_mm + 6);
        WalkItem *items = nullptr;
        e = scratch_alloc((void **)&items, (size_t)N * sizeof(WalkItem), st);
        if (e != cudaSuccess) { cudaFreeAsync(block_sums, st); return (int)e; }
        k_occ_bbox_init<<<1, 32, 0, st>>>(occ_mm);
        const uint32_t n_bytes = H * H * H / 8;
        k_occ_bbox<<<div_up(n_bytes, 256u), 256, 0, st>>>(grid, n_bytes, occ_mm);
        k_march_clip<<<nb, 128, 0, st>>>(rays_o, rays_d, bound, max_steps, N, C, H, nears, fars, noises, occ_mm, rays, items, n_items);
        k_march_walk<<<nb, 128, 0, st>>>(rays_o, rays_d, grid, bound, max_steps, C, H, items, n_items, rays, bitmap);
        k_march_block_sums<<<nb, 128, 0, st>>>(rays, N, block_sums);
        cudaFreeAsync(items, st);
    } else {
        k_march_count<<<nb, 128, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, block_sums, bitmap, nullptr);
    }
    k_march_scan<<<1, 1024, 0, st>>>(block_sums, nb, N, counter);
    k_march_offsets<<<nb, 128, 0, st>>>(rays, N, block_sums);
    e = cudaPeekAtLastError();
    cudaFreeAsync(block_sums, st);
    return (int)e;
}
}  // namespace

// development / test switch: 1 (default) = rays are walked only inside the widened bounding box of the occupied cells
// (single cascade, dt_gamma = 0), 0 = every ray is walked from its near point.  The samples are the same either way.
S3D_API int s3d_march_set_clip(int enable) {
    g_march_clip = enable ? 1 : 0;
    return 0;
}

S3D_API int s3d_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                 const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                                 int *rays, int *counter, const float *noises, void *stream) {
    if (N == 0) return 0;
    if (C == 0 || H == 0 || max_steps == 0) return S3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    // the count pass records the sample positions on the ray's step lattice; the write pass replays the lattice
    uint32_t *bitmap = nullptr;
    cudaError_t e = scratch_alloc((void **)&bitmap, (size_t)N * kBitmapWords * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    int rc = march_count_scan(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, st, bitmap);
    if (rc == 0) {
        if (dt_gamma == 0.0f && g_march_clip) {
            // constant step: one thread per sample (coalesced stores, t in closed form)
            k_march_write_samples<<<div_up(M, 256u), 256, 0, st>>>(rays_o, rays_d, bound, max_steps, N, C, H, M, nears, noises, rays, counter, bitmap,
                                                                   xyzs, dirs, deltas);
            k_march_write_overflow<<<div_up(N, 128u), 128, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, noises, rays,
                                                                    bitmap, xyzs, dirs, deltas);
        } else {
            k_march_write_bitmap<<<div_up(N, 128u), 128, 0, st>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears,
                                                                   fars, noises, rays, bitmap, xyzs, dirs, deltas);
        }
        rc = (int)cudaPeekAtLastError();
    }
    cudaFreeAsync(bitmap, st);
    return rc;
}

// Two-phase variant of the same op for callers that want an exactly-sized sample buffer: phase 1 counts and
// scans (rays[:,0..2] and counter are final afterwards), the host reads counter[0], allocates, phase 2 writes.
S3D_API int s3d_march_rays_train_count(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                       float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                       const float *nears, const float *fars, int *rays, int *counter,
                                       const float *noises, void *stream) {
    if (N == 0) return 0;
    if (C == 0 || H == 0 || max_steps == 0) return S3D_EINVAL;
    return march_count_scan(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, noises, rays, counter, as_stream(stream));
}

S3D_API int s3d_march_rays_train_write(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                       float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                       const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                                       const int *rays, const float *noises, void *stream) {
    if (N == 0) return 0;
    k_march_write<<<div_up(N, 128u), 128, 0, as_stream(stream)>>>(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H,
                                                                  M, nears, fars, noises, rays, xyzs, dirs, deltas);
    S3D_RETURN_LAST();
}

S3D_API int s3d_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                             const int *rays, uint32_t M, uint32_t N, float T_thresh,
                                             float *weights_sum, float *depth, float *image, void *stream) {
    if (N == 0) return 0;
    k_composite_train_fwd<<<div_up(N, 8u), 256, 0, as_stream(stream)>>>(sigmas, rgbs, deltas, rays, M, N, T_thresh,
                                                                          weights_sum, depth, image);
    S3D_RETURN_LAST();
}

S3D_API int s3d_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                              const float *sigmas, const float *rgbs, const float *deltas,
                                              const int *rays, const float *weights_sum, const float *image, uint32_t M,
                                              uint32_t N, float T_thresh, float *grad_sigmas, float *grad_rgbs,
                                              void *stream) {
    if (N == 0) return 0;
    k_composite_train_bwd<<<div_up(N, 8u), 256, 0, as_stream(stream)>>>(
        grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N, T_thresh, grad_sigmas, grad_rgbs);
    S3D_RETURN_LAST();
}

// The per-ray part of a distillation (sig_t / rgb_t given: teacher composited on the same samples) or fine-tuning (image_t
// [N,3] / depth_t [N] or NULL given) step in one launch: both forward scans, loss[0] += MSE term, loss[1] += L1 depth term,
// and the student's compositor backward with the loss gradient times `scale` (*scale_dev when not NULL: device-side
// GradScaler).  grad_sigmas [M] / grad_rgbs [M,3] must be zeroed by the caller.
S3D_API int s3d_distill_rays(const float *sig_t, const float *rgb_t, const float *image_t, const float *depth_t, const float *sig_s,
                             const float *rgb_s, const float *deltas, const int *rays, uint32_t M, uint32_t N, float T_thresh, float bg_color,
                             float scale, const float *scale_dev, float *loss, float *grad_sigmas, float *grad_rgbs, void *stream) {
    if (N == 0) return 0;
    if (!sig_t && !image_t) return S3D_EINVAL;
    k_distill_rays<<<div_up(N, 8u), 256, 0, as_stream(stream)>>>(sig_t, rgb_t, image_t, depth_t, sig_s, rgb_s, deltas, rays, M, N, T_thresh, bg_color,
                                                                   scale, scale_dev, loss, grad_sigmas, grad_rgbs);
    S3D_RETURN_LAST();
}

S3D_API int s3d_march_rays(uint32_t n_alive, uint32_t n_step, const int *rays_alive, const float *rays_t,
                           const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                           uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                           float *xyzs, float *dirs, float *deltas, const float *noises, void *stream) {
    if (n_alive == 0) return 0;
    k_march_rays<<<div_up(n_alive, 128u), 128, 0, as_stream(stream)>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d,
                                                                       bound, dt_gamma, max_steps, C, H, grid, nears,
                                                                       fars, xyzs, dirs, deltas, noises);
    S3D_RETURN_LAST();
}

S3D_API int s3d_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive, float *rays_t,
                               const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum,
                               float *depth, float *image, void *stream) {
    if (n_alive == 0) return 0;
    k_composite_rays<<<div_up(n_alive, 128u), 128, 0, as_stream(stream)>>>(n_alive, n_step, T_thresh, rays_alive, rays_t,
                                                                           sigmas, rgbs, deltas, weights_sum, depth, image);
    S3D_RETURN_LAST();
}
