/*
 * seal_oracle.c -- CPU restatement of the Seal-3D / torch-ngp hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (seal-3d_b200/) fails loudly when its CUDA library is missing; it
 * never falls back to this code.
 *
 * Every function restates the arithmetic of one reference kernel, cited file:line relative
 * to /root/reference.  Where the reference is float32 the restatement is float32 with the
 * same operation order; nvcc's default FMA contraction of the reference expressions is made
 * explicit with fmaf() and this file must be compiled with -ffp-contract=off so the host
 * compiler adds no contraction of its own.  Integer / index work is bit-exact by
 * construction.  Accumulations the reference does with float atomics in a nondeterministic
 * order (grid backward) are done here in double and rounded once.
 *
 * Pinning: the reference holds no golden vectors for this path.  The
 * restatement is pinned against (1) the closed-form SH of testing/test_shencoder.py and the
 * pure-torch proxy/colour functions of SealNeRF/seal_utils.py + color_utils.py, imported in
 * the build container (tests/golden/make_cpu_golden.py), and (2) outputs of the unmodified
 * reference extensions (oracle/_ref) run on the B200 (tests/golden/make_gpu_golden.py -> tests/golden/gpu_ref.npz,
 * checked on CPU by tests/test_oracle_golden.py).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC seal_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float signf_(float x) { return copysignf(1.0f, x); }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------ */
/* raymarching/src/raymarching.cu                                                         */
/* ------------------------------------------------------------------------------------ */

/* raymarching.cu:56-71 (10-bit interleave) */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
/* raymarching.cu:73-81 */
static inline uint32_t morton3_inv(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

/* raymarching.cu:92-145 kernel_near_far_from_aabb */
ORC_API void orc_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb,
                                    uint32_t N, float min_near, float *nears, float *fars) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float rdx = 1.0f / dx, rdy = 1.0f / dy, rdz = 1.0f / dz;
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, tmp;
        if (near > far) { tmp = near; near = far; far = tmp; }
        float ny = (aabb[1] - oy) * rdy, fy = (aabb[4] - oy) * rdy;
        if (ny > fy) { tmp = ny; ny = fy; fy = tmp; }
        if (near > fy || ny > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (ny > near) near = ny;
        if (fy < far) far = fy;
        float nz = (aabb[2] - oz) * rdz, fz = (aabb[5] - oz) * rdz;
        if (nz > fz) { tmp = nz; nz = fz; fz = tmp; }
        if (near > fz || nz > far) { nears[n] = fars[n] = FLT_MAX; continue; }
        if (nz > near) near = nz;
        if (fz < far) far = fz;
        if (near < min_near) near = min_near;
        nears[n] = near;
        fars[n] = far;
    }
}

/* raymarching.cu:163-198 kernel_sph_from_ray (tolerance-level: atan2/sqrt) */
ORC_API void orc_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords) {
    const float RPI = 0.3183098861837907f;
    for (uint32_t n = 0; n < N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = dx * dx + dy * dy + dz * dz;
        const float B = ox * dx + oy * dy + oz * dz;
        const float C = ox * ox + oy * oy + oz * oz - radius * radius;
        const float t = (-B + sqrtf(B * B - A * C)) / A;
        const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
        const float theta = atan2f(sqrtf(x * x + z * z), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = 2 * theta * RPI - 1;
        coords[n * 2 + 1] = phi * RPI;
    }
}

/* raymarching.cu:214-226 / 237-254 */
ORC_API void orc_morton3D(const int *coords, uint32_t N, int *indices) {
    for (uint32_t n = 0; n < N; n++)
        indices[n] = (int)morton3((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
ORC_API void orc_morton3D_invert(const int *indices, uint32_t N, int *coords) {
    for (uint32_t n = 0; n < N; n++) {
        const int ind = indices[n];
        coords[n * 3] = (int)morton3_inv((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int)morton3_inv((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int)morton3_inv((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:268-289 kernel_packbits: N bytes, bit i = grid[8n+i] > thresh */
ORC_API void orc_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:42-54 */
static inline int mip_from_pos(float x, float y, float z, float max_cascade) {
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) {
    const float mx = (float)((double)(dt * H) * 0.5);
    int e;
    frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, rH, H3, far, dt_min, dt_max, bound, dt_gamma;
    uint32_t C, H;
    const uint8_t *grid;
} march_ctx;

/* One visit of the marching loop body (raymarching.cu:359-400 == 427-479 == 750-804).
 * Returns 1 and fills xyz/dt when the cell is occupied (caller advances t by dt),
 * returns 0 after advancing *t past the empty voxel. */
static inline int march_visit(const march_ctx *c, float *t, float *x, float *y, float *z, float *dt_out) {
    const float tt0 = *t;
    *x = clampf(fmaf(tt0, c->dx, c->ox), -c->bound, c->bound);
    *y = clampf(fmaf(tt0, c->dy, c->oy), -c->bound, c->bound);
    *z = clampf(fmaf(tt0, c->dz, c->oz), -c->bound, c->bound);
    const float dt = clampf(tt0 * c->dt_gamma, c->dt_min, c->dt_max);
    const int l0 = mip_from_pos(*x, *y, *z, (float)c->C), l1 = mip_from_dt(dt, (float)c->H, (float)c->C);
    const int level = l0 > l1 ? l0 : l1;
    const float mip_bound = fminf(scalbnf(1.0f, level), c->bound);
    const float mip_rbound = 1.0f / mip_bound;
    /* 0.5 * (x * r + 1) * H evaluated in double from a float FMA (raymarching.cu:374-376) */
    const float Hm1 = (float)(c->H - 1);
    const int nx = (int)clampf((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const int ny = (int)clampf((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const int nz = (int)clampf((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const uint32_t index = (uint32_t)((float)level * c->H3 + (float)morton3((uint32_t)nx, (uint32_t)ny, (uint32_t)nz));
    const int occ = c->grid[index / 8] & (1 << (index % 8));
    *dt_out = dt;
    if (occ) return 1;
    /* distance to the exit face of this voxel (raymarching.cu:390-398) */
    const float tx = fmaf(fmaf(fmaf(0.5f, signf_(c->dx), (float)nx + 0.5f) * c->rH, 2.0f, -1.0f), mip_bound, -*x) * c->rdx;
    const float ty = fmaf(fmaf(fmaf(0.5f, signf_(c->dy), (float)ny + 0.5f) * c->rH, 2.0f, -1.0f), mip_bound, -*y) * c->rdy;
    const float tz = fmaf(fmaf(fmaf(0.5f, signf_(c->dz), (float)nz + 0.5f) * c->rH, 2.0f, -1.0f), mip_bound, -*z) * c->rdz;
    const float tt = tt0 + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    float tc = tt0;
    do { tc += clampf(tc * c->dt_gamma, c->dt_min, c->dt_max); } while (tc < tt);
    *t = tc;
    return 0;
}

static inline void march_ctx_init(march_ctx *c, const float *o, const float *d, const uint8_t *grid, float bound,
                                  float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, float far) {
    c->ox = o[0]; c->oy = o[1]; c->oz = o[2];
    c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
    c->rdx = 1.0f / c->dx; c->rdy = 1.0f / c->dy; c->rdz = 1.0f / c->dz;
    c->rH = 1.0f / (float)H;
    c->H3 = (float)(H * H * H);
    c->far = far;
    const float SQRT3 = 1.7320508075688772f;
    c->dt_min = 2 * SQRT3 / (float)max_steps;
    c->dt_max = 2 * SQRT3 * (float)(1 << (C - 1)) / (float)H;
    c->bound = bound; c->dt_gamma = dt_gamma; c->C = C; c->H = H; c->grid = grid;
}

/* raymarching.cu:312-480 kernel_march_rays_train.
 * Canonical (deterministic) slot order: ray n gets ray_index n and point_index = sum of the
 * counts of rays 0..n-1.  The reference assigns both by atomicAdd (order nondeterministic);
 * per-ray content (count, xyz/dir/delta sequence) and counter = (sum n, N) are identical. */
ORC_API void orc_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                                  float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                  const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                                  int *rays, int *counter, const float *noises) {
    uint32_t *counts = (uint32_t *)malloc(sizeof(uint32_t) * (N ? N : 1));
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        march_ctx c;
        march_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H, fars[n]);
        float t = nears[n];
        t = fmaf(clampf(t * dt_gamma, c.dt_min, c.dt_max), noises[n], t);
        uint32_t num = 0;
        float x, y, z, dt;
        while (t < c.far && num < max_steps) {
            if (march_visit(&c, &t, &x, &y, &z, &dt)) { num++; t += dt; }
        }
        counts[n] = num;
    }
    uint32_t base = (uint32_t)counter[0];
    uint32_t rbase = (uint32_t)counter[1];
    uint32_t *offs = (uint32_t *)malloc(sizeof(uint32_t) * (N ? N : 1));
    for (uint32_t n = 0; n < N; n++) { offs[n] = base; base += counts[n]; }
    counter[0] = (int)base;
    counter[1] = (int)(rbase + N);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t num = counts[n], off = offs[n];
        int *r = rays + ((int64_t)rbase + n) * 3;
        r[0] = (int)n; r[1] = (int)off; r[2] = (int)num;
        if (num == 0 || off + num > M) continue;
        march_ctx c;
        march_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, grid, bound, dt_gamma, max_steps, C, H, fars[n]);
        float t = nears[n];
        t = fmaf(clampf(t * dt_gamma, c.dt_min, c.dt_max), noises[n], t);
        float last_t = t, x, y, z, dt;
        uint32_t step = 0;
        float *px = xyzs + (int64_t)off * 3, *pd = dirs + (int64_t)off * 3, *pl = deltas + (int64_t)off * 2;
        while (t < c.far && step < num) {
            if (march_visit(&c, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
    free(counts);
    free(offs);
}

/* raymarching.cu:501-577 kernel_composite_rays_train_forward */
ORC_API void orc_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                              const int *rays, uint32_t M, uint32_t N, float T_thresh,
                                              float *weights_sum, float *depth, float *image) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num = (uint32_t)rays[n * 3 + 2];
        if (num == 0 || offset + num > M) {
            weights_sum[index] = 0; depth[index] = 0;
            image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        const float *s = sigmas + offset, *c = rgbs + (int64_t)offset * 3, *dl = deltas + (int64_t)offset * 2;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t step = 0; step < num; step++) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float w = alpha * T;
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            t += dl[1];
            d = fmaf(w, t, d);
            ws += w;
            T *= 1.0f - alpha;
            if (T < T_thresh) break;
            s++; c += 3; dl += 2;
        }
        weights_sum[index] = ws; depth[index] = d;
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* raymarching.cu:602-682 kernel_composite_rays_train_backward (grad_depth is not propagated) */
ORC_API void orc_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                               const float *sigmas, const float *rgbs, const float *deltas,
                                               const int *rays, const float *weights_sum, const float *image,
                                               uint32_t M, uint32_t N, float T_thresh, float *grad_sigmas,
                                               float *grad_rgbs) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num = (uint32_t)rays[n * 3 + 2];
        if (num == 0 || offset + num > M) continue;
        const float gws = grad_weights_sum[index];
        const float *gi = grad_image + (int64_t)index * 3;
        const float rf = image[index * 3], gf = image[index * 3 + 1], bf = image[index * 3 + 2], wsf = weights_sum[index];
        const float *s = sigmas + offset, *c = rgbs + (int64_t)offset * 3, *dl = deltas + (int64_t)offset * 2;
        float *gs = grad_sigmas + offset, *gc = grad_rgbs + (int64_t)offset * 3;
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        for (uint32_t step = 0; step < num; step++) {
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
            const float w = alpha * T;
            r = fmaf(w, c[0], r); g = fmaf(w, c[1], g); b = fmaf(w, c[2], b);
            ws += w;
            T *= 1.0f - alpha;
            gc[0] = gi[0] * w; gc[1] = gi[1] * w; gc[2] = gi[2] * w;
            gs[0] = dl[0] * (gi[0] * (T * c[0] - (rf - r)) + gi[1] * (T * c[1] - (gf - g)) +
                             gi[2] * (T * c[2] - (bf - b)) + gws * (1 - wsf));
            if (T < T_thresh) break;
            s++; c += 3; dl += 2; gs++; gc += 3;
        }
    }
}

/* raymarching.cu:701-805 kernel_march_rays (inference, fixed n_step per alive ray) */
ORC_API void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int *rays_alive, const float *rays_t,
                            const float *rays_o, const float *rays_d, float bound, float dt_gamma, uint32_t max_steps,
                            uint32_t C, uint32_t H, const uint8_t *grid, const float *nears, const float *fars,
                            float *xyzs, float *dirs, float *deltas, const float *noises) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int index = rays_alive[n];
        march_ctx c;
        march_ctx_init(&c, rays_o + (int64_t)index * 3, rays_d + (int64_t)index * 3, grid, bound, dt_gamma, max_steps, C, H,
                       fars[index]);
        float t = rays_t[index];
        t = fmaf(clampf(t * dt_gamma, c.dt_min, c.dt_max), noises[n], t);
        float last_t = t, x, y, z, dt;
        uint32_t step = 0;
        float *px = xyzs + n * n_step * 3, *pd = dirs + n * n_step * 3, *pl = deltas + n * n_step * 2;
        while (t < c.far && step < n_step) {
            if (march_visit(&c, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z;
                pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt;
                pl[0] = dt; pl[1] = t - last_t;
                last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* raymarching.cu:819-905 kernel_composite_rays (in-place accumulate, kill ray with -1) */
ORC_API void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive, float *rays_t,
                                const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum,
                                float *depth, float *image) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int index = rays_alive[n];
        const float *s = sigmas + n * n_step, *c = rgbs + n * n_step * 3, *dl = deltas + n * n_step * 2;
        float t = rays_t[index], ws = weights_sum[index], d = depth[index];
        float r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            if (dl[0] == 0) break;
            const float alpha = 1.0f - expf(-s[0] * dl[0]);
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
        image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* ------------------------------------------------------------------------------------ */
/* gridencoder/src/gridencoder.cu                                                         */
/* ------------------------------------------------------------------------------------ */

/* gridencoder.cu:51-84 fast_hash + get_grid_index (ch = 0) */
static inline uint32_t grid_index(uint32_t gridtype, int align_corners, uint32_t D, uint32_t C, uint32_t hashmap_size,
                                  uint32_t resolution, const uint32_t *pg) {
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pg[d] * stride;
        stride *= align_corners ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) {
        index = 0;
        for (uint32_t d = 0; d < D; d++) index ^= pg[d] * primes[d];
    }
    return (index % hashmap_size) * C;
}

typedef struct {
    uint32_t hashmap_size, resolution;
    float scale;
} level_info;

/* Optional precomputed per-level scales.  exp2f is the one value of this path that is not reproducible across math
 * libraries (CUDA's exp2f is not correctly rounded, glibc's is): a 1-ulp difference in exp2f(7) moves
 * pos = x*scale+0.5 of the finest level by ~1e-4.  GPU parity tests therefore read the device's scales back
 * (s3d_grid_level_scales) and install them here; everything else is restated independently. */
static const float *volatile g_level_scales = NULL;
ORC_API void orc_set_level_scales(const float *scales) { g_level_scales = scales; }
static inline level_info level_setup(const int *offsets, uint32_t level, float S, uint32_t H) {
    const float *ls = g_level_scales;
    level_info li;
    li.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    li.scale = ls ? ls[level] : fmaf(exp2f((float)level * S), (float)H, -1.0f); /* gridencoder.cu:138 */
    li.resolution = (uint32_t)ceilf(li.scale) + 1;              /* gridencoder.cu:139 */
    return li;
}

/* gridencoder.cu:88-242 kernel_grid; embeddings float32; outputs [L,B,C]; dy_dx [B,L,D,C] or NULL.
 * emb_round: 0 = float32 accumulate; 1 = round every partial sum to fp16 like the
 * scalar_t=half instantiation does (table values are expected to be fp16-representable). */
static inline float round_half(float v);

ORC_API void orc_grid_encode_forward(const float *inputs, const float *emb, const int *offsets, float *outputs,
                                     uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, float *dy_dx,
                                     uint32_t gridtype, int align_corners, uint32_t interp, int half_accum) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t level = 0; level < (int64_t)L; level++) {
        for (int64_t b = 0; b < (int64_t)B; b++) {
            const float *in = inputs + b * D;
            const float *grid = emb + (int64_t)(uint32_t)offsets[level] * C;
            float *out = outputs + (level * B + b) * C;
            float *dd = dy_dx ? dy_dx + (b * L + level) * D * C : NULL;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) {
                for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
                if (dd) for (uint32_t i = 0; i < D * C; i++) dd[i] = 0;
                continue;
            }
            const level_info li = level_setup(offsets, (uint32_t)level, S, H);
            float pos[5], pos_deriv[5] = {1.0f, 0, 0, 0, 0}; /* reference initialises only [0] (gridencoder.cu:143) */
            uint32_t pg[5];
            for (uint32_t d = 0; d < D; d++) {
                pos[d] = fmaf(in[d], li.scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pg[d];
                if (interp == 1) {
                    pos_deriv[d] = 6 * pos[d] * (1.0f - pos[d]);
                    pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
                }
            }
            float res[8] = {0};
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pl[5];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                const uint32_t index = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pl);
                for (uint32_t ch = 0; ch < C; ch++) {
                    if (half_accum) res[ch] = round_half(res[ch] + round_half(w * grid[index + ch]));
                    else res[ch] = fmaf(w, grid[index + ch], res[ch]);
                }
            }
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = res[ch];
            if (dd) {
                for (uint32_t gd = 0; gd < D; gd++) {
                    float rg[8] = {0};
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = li.scale;
                        uint32_t pl[5];
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                            else { w *= pos[d]; pl[d] = pg[d] + 1; }
                        }
                        pl[gd] = pg[gd];
                        const uint32_t il = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pl);
                        pl[gd] = pg[gd] + 1;
                        const uint32_t ir = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pl);
                        for (uint32_t ch = 0; ch < C; ch++)
                            rg[ch] += w * (grid[ir + ch] - grid[il + ch]) * pos_deriv[gd];   /* linear: {1,0,0,..}: only d/dx0 is non-zero, like the reference kernel's output (tests/golden/gpu_ref.npz) */
                    }
                    for (uint32_t ch = 0; ch < C; ch++) dd[gd * C + ch] = rg[ch];
                }
            }
        }
    }
}

/* gridencoder.cu:246-337 kernel_grid_backward (+ :341-366 kernel_input_backward).
 * grad [L,B,C]; grad_emb is ADDED into (caller zero-fills, like grid.py:77).  Per-entry sums
 * are accumulated in double and rounded once (the reference's float atomics are
 * order-nondeterministic). */
ORC_API void orc_grid_encode_backward(const float *grad, const float *inputs, const int *offsets, float *grad_emb,
                                      uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                      const float *dy_dx, float *grad_inputs, uint32_t gridtype, int align_corners,
                                      uint32_t interp) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t level = 0; level < (int64_t)L; level++) {
        const level_info li = level_setup(offsets, (uint32_t)level, S, H);
        double *acc = (double *)calloc((size_t)li.hashmap_size * C, sizeof(double));
        for (int64_t b = 0; b < (int64_t)B; b++) {
            const float *in = inputs + b * D;
            const float *g = grad + (level * B + b) * C;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            float pos[5];
            uint32_t pg[5];
            for (uint32_t d = 0; d < D; d++) {
                pos[d] = fmaf(in[d], li.scale, align_corners ? 0.0f : 0.5f);
                pg[d] = (uint32_t)floorf(pos[d]);
                pos[d] -= (float)pg[d];
                if (interp == 1) pos[d] = pos[d] * pos[d] * (3.0f - 2.0f * pos[d]);
            }
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                uint32_t pl[5];
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) { w *= 1 - pos[d]; pl[d] = pg[d]; }
                    else { w *= pos[d]; pl[d] = pg[d] + 1; }
                }
                const uint32_t index = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pl);
                for (uint32_t ch = 0; ch < C; ch++) acc[index + ch] += (double)(w * g[ch]);
            }
        }
        float *ge = grad_emb + (int64_t)(uint32_t)offsets[level] * C;
        for (size_t i = 0; i < (size_t)li.hashmap_size * C; i++) ge[i] += (float)acc[i];
        free(acc);
    }
    if (dy_dx && grad_inputs) {
        for (int64_t t = 0; t < (int64_t)B * D; t++) {
            const int64_t b = t / D, d = t - b * D;
            float r = 0;
            for (uint32_t l = 0; l < L; l++)
                for (uint32_t ch = 0; ch < C; ch++)
                    r += grad[(l * (int64_t)B + b) * C + ch] * dy_dx[((b * L + l) * D + d) * C + ch];
            grad_inputs[t] = r;
        }
    }
}

/* gridencoder.cu:504-607 kernel_grad_tv (adds into grad) */
ORC_API void orc_grad_total_variation(const float *inputs, const float *emb, float *grad, const int *offsets,
                                      float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                      uint32_t gridtype, int align_corners) {
    for (uint32_t level = 0; level < L; level++) {
        const level_info li = level_setup(offsets, level, S, H);
        const float *grid = emb + (int64_t)(uint32_t)offsets[level] * C;
        float *gg = grad + (int64_t)(uint32_t)offsets[level] * C;
        for (uint32_t b = 0; b < B; b++) {
            const float *in = inputs + (int64_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++) if (in[d] < 0 || in[d] > 1) oob = 1;
            if (oob) continue;
            uint32_t pg[5];
            for (uint32_t d = 0; d < D; d++) pg[d] = (uint32_t)floorf(fmaf(in[d], li.scale, align_corners ? 0.0f : 0.5f));
            float res[8] = {0}, idelta[8] = {0};
            const uint32_t index = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pg);
            const float w = weight / (2 * D);
            for (uint32_t d = 0; d < D; d++) {
                const uint32_t cur = pg[d];
                if (cur < li.resolution) {
                    pg[d] = cur + 1;
                    const uint32_t ir = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pg);
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float gv = grid[index + ch] - grid[ir + ch];
                        res[ch] += gv; idelta[ch] += gv * gv;
                    }
                }
                if (cur > 0) {
                    pg[d] = cur - 1;
                    const uint32_t il = grid_index(gridtype, align_corners, D, C, li.hashmap_size, li.resolution, pg);
                    for (uint32_t ch = 0; ch < C; ch++) {
                        const float gv = grid[index + ch] - grid[il + ch];
                        res[ch] += gv; idelta[ch] += gv * gv;
                    }
                }
                pg[d] = cur;
            }
            for (uint32_t ch = 0; ch < C; ch++) gg[index + ch] += w * res[ch] * (1.0f / sqrtf(idelta[ch] + 1e-9f));
        }
    }
}

/* IEEE binary16 round-to-nearest-even of a float, returned as float */
static inline float round_half(float v) {
    _Float16 h = (_Float16)v;
    return (float)h;
}
ORC_API void orc_round_to_half(const float *in, float *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = round_half(in[i]);
}

/* ------------------------------------------------------------------------------------ */
/* shencoder/src/shencoder.cu                                                             */
/* ------------------------------------------------------------------------------------ */
/* The reference hard-codes, for degree<=8, the real SH basis (Condon-Shortley phase) as
 * polynomials c_m(x,y)*Q_l^m(z), s_m(x,y)*Q_l^m(z) (shencoder.cu:49-121) and their analytic
 * partial derivatives (:130-350).  The same polynomials are generated here in double by the
 * associated-Legendre recurrence with the (1-z^2)^{m/2} factor carried by
 * c_m + i s_m = (x + i y)^m, which reproduces the hard-coded forms as functions of (x,y,z)
 * (also off the unit sphere).  index = l*l + l + m. */
static double sh_K(int l, int m) { /* sqrt((2l+1)/(4pi) (l-m)!/(l+m)!) */
    double r = (2.0 * l + 1.0) / (4.0 * M_PI);
    for (int k = l - m + 1; k <= l + m; k++) r /= (double)k;
    return sqrt(r);
}

static void sh_eval(double x, double y, double z, int deg, double *Y, double *dYx, double *dYy, double *dYz) {
    double c[9], s[9];
    c[0] = 1; s[0] = 0;
    for (int m = 1; m < deg; m++) { c[m] = x * c[m - 1] - y * s[m - 1]; s[m] = x * s[m - 1] + y * c[m - 1]; }
    for (int m = 0; m < deg; m++) {
        /* Q_m^m = (-1)^m (2m-1)!! ; Q_{m+1}^m = (2m+1) z Q_m^m ; (l-m) Q_l^m = (2l-1) z Q_{l-1}^m - (l+m-1) Q_{l-2}^m */
        double qmm = 1;
        for (int k = 1; k <= m; k++) qmm *= -(2.0 * k - 1.0);
        double q2 = 0, q1 = 0, dq2 = 0, dq1 = 0; /* Q_{l-2}, Q_{l-1} and their d/dz */
        for (int l = m; l < deg; l++) {
            double q, dq;
            if (l == m) { q = qmm; dq = 0; }
            else {
                q = ((2.0 * l - 1.0) * z * q1 - (l + m - 1.0) * q2) / (double)(l - m);
                dq = ((2.0 * l - 1.0) * (q1 + z * dq1) - (l + m - 1.0) * dq2) / (double)(l - m);
            }
            const double K = sh_K(l, m);
            const int base = l * l + l;
            if (m == 0) {
                Y[base] = K * q;
                if (dYx) { dYx[base] = 0; dYy[base] = 0; dYz[base] = K * dq; }
            } else {
                const double k2 = M_SQRT2 * K;
                Y[base + m] = k2 * q * c[m];
                Y[base - m] = k2 * q * s[m];
                if (dYx) {
                    dYx[base + m] = k2 * q * m * c[m - 1];
                    dYy[base + m] = -k2 * q * m * s[m - 1];
                    dYz[base + m] = k2 * dq * c[m];
                    dYx[base - m] = k2 * q * m * s[m - 1];
                    dYy[base - m] = k2 * q * m * c[m - 1];
                    dYz[base - m] = k2 * dq * s[m];
                }
            }
            q2 = q1; q1 = q; dq2 = dq1; dq1 = dq;
        }
    }
}

/* shencoder.cu:28-355 kernel_sh; outputs [B, C*C]; dy_dx [B, 3, C*C] or NULL */
ORC_API void orc_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C, float *dy_dx) {
    const uint32_t C2 = C * C;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)B; b++) {
        double Y[64], gx[64], gy[64], gz[64];
        sh_eval(inputs[b * D], inputs[b * D + 1], inputs[b * D + 2], (int)C, Y, dy_dx ? gx : NULL, gy, gz);
        for (uint32_t i = 0; i < C2; i++) outputs[b * C2 + i] = (float)Y[i];
        if (dy_dx) {
            float *o = dy_dx + b * D * C2;
            for (uint32_t i = 0; i < C2; i++) { o[i] = (float)gx[i]; o[C2 + i] = (float)gy[i]; o[2 * C2 + i] = (float)gz[i]; }
        }
    }
}

/* shencoder.cu:359-382 kernel_sh_backward: grad_inputs[b,d] += sum_ch grad * dy_dx */
ORC_API void orc_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C,
                                    const float *dy_dx, float *grad_inputs) {
    (void)inputs;
    const uint32_t C2 = C * C;
    for (int64_t t = 0; t < (int64_t)B * D; t++) {
        const int64_t b = t / D, d = t - b * D;
        float acc = grad_inputs[t];
        for (uint32_t ch = 0; ch < C2; ch++) acc += grad[b * C2 + ch] * dy_dx[(b * D + d) * C2 + ch];
        grad_inputs[t] = acc;
    }
}

/* ------------------------------------------------------------------------------------ */
/* freqencoder/src/freqencoder.cu                                                         */
/* ------------------------------------------------------------------------------------ */
/* freqencoder.cu:30-60 kernel_freq: out[b,c] = x (c<D) else sin(2^f x + (col%2) pi/2) */
ORC_API void orc_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *outputs) {
    (void)deg;
    const float PI = 3.141592653589793f;
    for (int64_t t = 0; t < (int64_t)B * C; t++) {
        const int64_t b = t / C, c = t - b * C;
        if (c < D) outputs[t] = inputs[b * D + c];
        else {
            const uint32_t col = (uint32_t)(c / D - 1), d = (uint32_t)(c % D), freq = col / 2;
            const float phase = (float)(col % 2) * (PI / 2);
            outputs[t] = sinf(scalbnf(inputs[b * D + d], (int)freq) + phase);
        }
    }
}
/* freqencoder.cu:63-94 kernel_freq_backward */
ORC_API void orc_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg,
                                      uint32_t C, float *grad_inputs) {
    for (int64_t t = 0; t < (int64_t)B * D; t++) {
        const int64_t b = t / D, d = t - b * D;
        const float *g = grad + b * C, *o = outputs + b * C;
        float r = g[d];
        g += D; o += D;
        for (uint32_t f = 0; f < deg; f++) {
            r += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
            g += 2 * D; o += 2 * D;
        }
        grad_inputs[t] = r;
    }
}

/* ------------------------------------------------------------------------------------ */
/* ffmlp/src/ffmlp.cu : bias-free MLP, weights row-major [out,in] blocks concatenated     */
/* ------------------------------------------------------------------------------------ */
/* activation ids ffmlp.cu:22-33 / ffmlp.py:89-96 / utils.h:424-582 */
static inline double act_fwd(uint32_t a, double x) {
    switch (a) {
        case 0: return x > 0 ? x : 0;
        case 1: return exp(x);
        case 2: return sin(x);
        case 3: return 1.0 / (1.0 + exp(-x));
        case 4: return 0.5 * (x + sqrt(x * x + 4.0));
        case 5: return log(exp(x) + 1.0);
        default: return x;
    }
}
/* derivative expressed through the saved forward activation y (what the reference's
 * warp_activation_backward receives, utils.h:537-582) */
static inline double act_bwd_from_out(uint32_t a, double y) {
    switch (a) {
        case 0: return y > 0 ? 1.0 : 0.0;
        case 1: return y;
        case 3: return y * (1.0 - y);
        case 4: { const double y2 = y * y; return y2 / (y2 + 1.0); }
        case 5: return 1.0 - exp(-y);
        case 6: return 1.0;
        default: return 0.0; /* sine: not invertible from the output; unsupported in the reference backward */
    }
}

/* ffmlp.cu:332-407 + :635-671.  inputs [B,in]; forward_buffer [num_layers,B,hidden] (NULL ok);
 * outputs [B,out].  num_layers hidden activations => num_layers+1 matmuls.  When
 * round_half_act != 0 every stored activation is rounded to fp16 (what the fp16 kernels keep). */
ORC_API void orc_ffmlp_forward(const float *inputs, const float *weights, uint32_t B, uint32_t in_dim, uint32_t out_dim,
                               uint32_t hidden, uint32_t num_layers, uint32_t act, uint32_t out_act,
                               float *forward_buffer, float *outputs, int round_half_act) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)B; b++) {
        double cur[256], nxt[256];
        const float *w = weights;
        for (uint32_t j = 0; j < hidden; j++) {
            double a = 0;
            for (uint32_t k = 0; k < in_dim; k++) a += (double)inputs[b * in_dim + k] * (double)w[j * in_dim + k];
            a = act_fwd(act, a);
            cur[j] = round_half_act ? (double)round_half((float)a) : a;
        }
        if (forward_buffer) for (uint32_t j = 0; j < hidden; j++) forward_buffer[(0 * (int64_t)B + b) * hidden + j] = (float)cur[j];
        w += hidden * in_dim;
        for (uint32_t l = 1; l < num_layers; l++) {
            for (uint32_t j = 0; j < hidden; j++) {
                double a = 0;
                for (uint32_t k = 0; k < hidden; k++) a += cur[k] * (double)w[j * hidden + k];
                a = act_fwd(act, a);
                nxt[j] = round_half_act ? (double)round_half((float)a) : a;
            }
            memcpy(cur, nxt, sizeof(double) * hidden);
            if (forward_buffer) for (uint32_t j = 0; j < hidden; j++) forward_buffer[(l * (int64_t)B + b) * hidden + j] = (float)cur[j];
            w += hidden * hidden;
        }
        for (uint32_t j = 0; j < out_dim; j++) {
            double a = 0;
            for (uint32_t k = 0; k < hidden; k++) a += cur[k] * (double)w[j * hidden + k];
            outputs[b * out_dim + j] = (float)act_fwd(out_act, a);
        }
    }
}

/* ffmlp.cu:411-518 + :749-894.  grad [B,out]; backward_buffer [num_layers,B,hidden] (NULL ok);
 * grad_inputs [B,in] or NULL; grad_weights flat (overwritten).  Output activation is ignored in
 * the backward exactly like the reference ("I gonna discard output_activation", ffmlp.cu:781). */
ORC_API void orc_ffmlp_backward(const float *grad, const float *inputs, const float *weights, const float *forward_buffer,
                                uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers,
                                uint32_t act, float *backward_buffer, float *grad_inputs, float *grad_weights) {
    const size_t nW = (size_t)hidden * in_dim + (size_t)hidden * hidden * (num_layers - 1) + (size_t)out_dim * hidden;
    double *gw = (double *)calloc(nW, sizeof(double));
    const float *w_out = weights + (size_t)hidden * in_dim + (size_t)hidden * hidden * (num_layers - 1);
    double *gw_out = gw + (size_t)hidden * in_dim + (size_t)hidden * hidden * (num_layers - 1);
    for (int64_t b = 0; b < (int64_t)B; b++) {
        double dcur[256], dnext[256];
        const float *hl = forward_buffer + ((int64_t)(num_layers - 1) * B + b) * hidden;
        for (uint32_t j = 0; j < out_dim; j++)
            for (uint32_t k = 0; k < hidden; k++) gw_out[j * hidden + k] += (double)grad[b * out_dim + j] * (double)hl[k];
        for (uint32_t k = 0; k < hidden; k++) {
            double a = 0;
            for (uint32_t j = 0; j < out_dim; j++) a += (double)grad[b * out_dim + j] * (double)w_out[j * hidden + k];
            dcur[k] = a * act_bwd_from_out(act, hl[k]);
        }
        if (backward_buffer) for (uint32_t k = 0; k < hidden; k++) backward_buffer[(0 * (int64_t)B + b) * hidden + k] = (float)dcur[k];
        for (uint32_t i = 0; i + 1 < num_layers; i++) {
            const uint32_t mi = num_layers - 2 - i; /* hidden matrix index, maps h_mi -> h_{mi+1} */
            const float *w = weights + (size_t)hidden * in_dim + (size_t)hidden * hidden * mi;
            double *g = gw + (size_t)hidden * in_dim + (size_t)hidden * hidden * mi;
            const float *hp = forward_buffer + ((int64_t)mi * B + b) * hidden;
            for (uint32_t j = 0; j < hidden; j++)
                for (uint32_t k = 0; k < hidden; k++) g[j * hidden + k] += dcur[j] * (double)hp[k];
            for (uint32_t k = 0; k < hidden; k++) {
                double a = 0;
                for (uint32_t j = 0; j < hidden; j++) a += dcur[j] * (double)w[j * hidden + k];
                dnext[k] = a * act_bwd_from_out(act, hp[k]);
            }
            memcpy(dcur, dnext, sizeof(double) * hidden);
            if (backward_buffer) for (uint32_t k = 0; k < hidden; k++) backward_buffer[((i + 1) * (int64_t)B + b) * hidden + k] = (float)dcur[k];
        }
        for (uint32_t j = 0; j < hidden; j++)
            for (uint32_t k = 0; k < in_dim; k++) gw[j * in_dim + k] += dcur[j] * (double)inputs[b * in_dim + k];
        if (grad_inputs)
            for (uint32_t k = 0; k < in_dim; k++) {
                double a = 0;
                for (uint32_t j = 0; j < hidden; j++) a += dcur[j] * (double)weights[j * in_dim + k];
                grad_inputs[b * in_dim + k] = (float)a;
            }
    }
    for (size_t i = 0; i < nW; i++) grad_weights[i] = (float)gw[i];
    free(gw);
}

/* ------------------------------------------------------------------------------------ */
/* SealNeRF/seal_utils.py : bbox proxy mapping and colour edits                           */
/* ------------------------------------------------------------------------------------ */

/* seal_utils.py:630-664 moller_trumbore, one ray vs F triangles, "any hit" */
static int mt_any_hit(const float *o, const float *d, const float *tris, uint32_t F) {
    for (uint32_t f = 0; f < F; f++) {
        const float *v0 = tris + f * 9, *v1 = v0 + 3, *v2 = v0 + 6;
        const float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
        const float e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
        const float nrm[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const float invdet = 1.0f / -((d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2]) + 1e-8f);
        const float a0[3] = {o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]};
        const float da0[3] = {a0[1] * d[2] - a0[2] * d[1], a0[2] * d[0] - a0[0] * d[2], a0[0] * d[1] - a0[1] * d[0]};
        const float u = (da0[0] * e2[0] + da0[1] * e2[1] + da0[2] * e2[2]) * invdet;
        const float v = -(da0[0] * e1[0] + da0[1] * e1[1] + da0[2] * e1[2]) * invdet;
        const float t = (a0[0] * nrm[0] + a0[1] * nrm[1] + a0[2] * nrm[2]) * invdet;
        if (t >= 0.0f && u >= 0.0f && v >= 0.0f && (u + v) <= 1.0f) return 1;
    }
    return 0;
}

/* seal_utils.py:132-153 map_mask + :667-685 points_in_mesh */
ORC_API void orc_seal_map_mask(const float *points, int64_t P, const float *bounds /*[nb,2,3]*/, uint32_t nb,
                               const float *tris /*[F,3,3]*/, uint32_t F, const float *test_dir /*[3] or NULL*/,
                               uint8_t *mask) {
    const float d0[3] = {0.4395064455f, 0.617598629942f, 0.652231566745f};
    const float *dir = test_dir ? test_dir : d0;
    const float ndir[3] = {-dir[0], -dir[1], -dir[2]};
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; p++) {
        const float *x = points + p * 3;
        int m = 0;
        if (x[0] != 0 && x[1] != 0 && x[2] != 0) {
            for (uint32_t i = 0; i < nb && !m; i++) {
                const float *lo = bounds + i * 6, *hi = lo + 3;
                if (hi[0] > x[0] && x[0] > lo[0] && hi[1] > x[1] && x[1] > lo[1] && hi[2] > x[2] && x[2] > lo[2]) m = 1;
            }
        }
        if (m) m = mt_any_hit(x, dir, tris, F) && mt_any_hit(x, ndir, tris, F);
        mask[p] = (uint8_t)m;
    }
}

/* seal_utils.py:237-279 SealBBoxMapper.map_to_origin.  transform = inverse 4x4 (row-major),
 * rotation = inverse 3x3, scale = 1/scale, center = from_center; optional map_source teleport. */
ORC_API void orc_seal_bbox_map_to_origin(const float *points, const float *dirs, int64_t P, const float *transform,
                                         const float *rotation, const float *scale, const float *center,
                                         const float *bounds, uint32_t nb, const float *tris, uint32_t F,
                                         const float *test_dir, const float *src_bound /*[2,3] or NULL*/,
                                         const float *map_source /*[3] or NULL*/, float *out_points, float *out_dirs,
                                         uint8_t *mask) {
    orc_seal_map_mask(points, P, bounds, nb, tris, F, test_dir, mask);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; p++) {
        const float *x = points + p * 3;
        float *ox = out_points + p * 3;
        ox[0] = x[0]; ox[1] = x[1]; ox[2] = x[2];
        if (dirs) { out_dirs[p * 3] = dirs[p * 3]; out_dirs[p * 3 + 1] = dirs[p * 3 + 1]; out_dirs[p * 3 + 2] = dirs[p * 3 + 2]; }
        if (src_bound && map_source) {
            const float *lo = src_bound, *hi = src_bound + 3;
            if (hi[0] > x[0] && x[0] > lo[0] && hi[1] > x[1] && x[1] > lo[1] && hi[2] > x[2] && x[2] > lo[2]) {
                ox[0] = map_source[0]; ox[1] = map_source[1]; ox[2] = map_source[2];
            }
        }
        if (!mask[p]) continue;
        for (int i = 0; i < 3; i++) {
            const float *r = transform + i * 4;
            const float tp = r[0] * x[0] + r[1] * x[1] + r[2] * x[2] + r[3];
            ox[i] = (tp - center[i]) * scale[i] + center[i];
        }
        if (dirs) {
            const float *d = dirs + p * 3;
            for (int i = 0; i < 3; i++) out_dirs[p * 3 + i] = rotation[i * 3] * d[0] + rotation[i * 3 + 1] * d[1] + rotation[i * 3 + 2] * d[2];
        }
    }
}

/* color_utils.py:31-43 rgb2hsv_torch */
static void rgb2hsv(const float *rgb, float *hsv) {
    const float r = rgb[0], g = rgb[1], b = rgb[2];
    float cmax = r; int idx = 0;
    if (g > cmax) { cmax = g; idx = 1; }
    if (b > cmax) { cmax = b; idx = 2; }
    const float cmin = fminf(r, fminf(g, b));
    const float delta = cmax - cmin;
    float h;
    if (delta == 0) h = 0;
    else if (idx == 0) { h = fmodf((g - b) / delta, 6.0f); if (h < 0) h += 6.0f; }
    else if (idx == 1) h = (b - r) / delta + 2;
    else h = (r - g) / delta + 4;
    hsv[0] = h / 6.0f;
    hsv[1] = (cmax == 0) ? 0.0f : delta / cmax;
    hsv[2] = cmax;
}
/* color_utils.py:46-63 hsv2rgb_torch */
static void hsv2rgb(const float *hsv, float *rgb) {
    const float h = hsv[0], s = hsv[1], v = hsv[2];
    const float c = v * s;
    float hm = fmodf(h * 6.0f, 2.0f); if (hm < 0) hm += 2.0f; /* torch % is floor-mod */
    const float x = c * (-fabsf(hm - 1) + 1.0f);
    const float m = v - c;
    const int idx = ((int)(uint8_t)(int)(h * 6.0f)) % 6;
    float r, g, b;
    switch (idx) {
        case 0: r = c; g = x; b = 0; break;
        case 1: r = x; g = c; b = 0; break;
        case 2: r = 0; g = c; b = x; break;
        case 3: r = 0; g = x; b = c; break;
        case 4: r = x; g = 0; b = c; break;
        default: r = c; g = 0; b = x; break;
    }
    rgb[0] = r + m; rgb[1] = g + m; rgb[2] = b + m;
}

/* seal_utils.py:739-751 modify_hsv */
ORC_API void orc_seal_modify_hsv(const float *rgb, int64_t P, const float *mod, float *out) {
    for (int64_t p = 0; p < P; p++) {
        float hsv[3];
        rgb2hsv(rgb + p * 3, hsv);
        hsv[0] += mod[0]; hsv[1] += mod[1]; hsv[2] += mod[2];
        hsv2rgb(hsv, out + p * 3);
    }
}
/* seal_utils.py:754-769 modify_rgb: replace H,S; V = clamp(V_target + (V - mean_batch V) + light, 0, 1) */
ORC_API void orc_seal_modify_rgb(const float *rgb, int64_t P, const float *target_rgb, float light_offset, float *out) {
    if (P == 0) return;
    float thsv[3];
    rgb2hsv(target_rgb, thsv);
    double sum = 0;
    float *vs = (float *)malloc(sizeof(float) * P);
    for (int64_t p = 0; p < P; p++) { float hsv[3]; rgb2hsv(rgb + p * 3, hsv); vs[p] = hsv[2]; sum += hsv[2]; }
    const float mean = (float)(sum / (double)P);
    for (int64_t p = 0; p < P; p++) {
        float hsv[3] = {thsv[0], thsv[1], fminf(1.0f, fmaxf(0.0f, thsv[2] + (vs[p] - mean) + light_offset))};
        hsv2rgb(hsv, out + p * 3);
    }
    free(vs);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Brush / Anchor mappers and the texture colour map
 * ------------------------------------------------------------------------------------------------------------- */
/* seal_utils.py:728-736 project_points */
static void project_point(const float *n, const float *o, const float *p, float *out) {
    const float v[3] = {p[0] - o[0], p[1] - o[1], p[2] - o[2]};
    const float s = (v[0] * n[0] + v[1] * n[1] + v[2] * n[2]) / (n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    out[0] = p[0] - s * n[0]; out[1] = p[1] - s * n[1]; out[2] = p[2] - s * n[2];
}

/* seal_utils.py:408-453 SealBrushMapper.map_to_origin.  mode 0 = 'linear', 1 = 'dry' (no space mapping).
 * border [K,3]; distance = min_k ||proj - border_k|| (torch.cdist(...).min(1)), evaluated directly. */
ORC_API void orc_seal_brush_map_to_origin(const float *points, int64_t P, const float *bounds, uint32_t nb, const float *tris,
                                          uint32_t F, const float *test_dir, const float *normal_expand, const float *center,
                                          const float *border, uint32_t K, float att_dist, int mode, float *out_points,
                                          uint8_t *mask) {
    orc_seal_map_mask(points, P, bounds, nb, tris, F, test_dir, mask);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; p++) {
        const float *x = points + p * 3;
        float *ox = out_points + p * 3;
        ox[0] = x[0]; ox[1] = x[1]; ox[2] = x[2];
        if (!mask[p] || mode == 1) continue;
        float pr[3];
        project_point(normal_expand, center, x, pr);
        float best = INFINITY;
        for (uint32_t k = 0; k < K; k++) {
            const float *b = border + k * 3;
            const float dx = pr[0] - b[0], dy = pr[1] - b[1], dz = pr[2] - b[2];
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            if (d < best) best = d;
        }
        for (int i = 0; i < 3; i++) ox[i] = x[i] - normal_expand[i];
        if (att_dist > best) {
            const float c = fabsf(att_dist - best) / att_dist;
            for (int i = 0; i < 3; i++) ox[i] += c * normal_expand[i];
        }
    }
}

/* seal_utils.py:514-570 SealAnchorMapper.map_to_origin.  Reference quirk kept: map_mask only gates the early exit
 * (`if not map_mask.any(): return`); once any sample is inside the map region the cone test runs on ALL samples and
 * the returned mask is the cone mask, not ANDed with map_mask. */
ORC_API void orc_seal_anchor_map_to_origin(const float *points, int64_t P, const float *bounds, uint32_t nb, const float *tris,
                                           uint32_t F, const float *test_dir, const float *v_anchor, const float *v_offset,
                                           const float *v_h, float len_h, float radius, const float *scale, float *out_points,
                                           uint8_t *mask) {
    orc_seal_map_mask(points, P, bounds, nb, tris, F, test_dir, mask);
    int any = 0;
    for (int64_t p = 0; p < P; p++) any |= mask[p];
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < P; p++) {
        const float *x = points + p * 3;
        float *ox = out_points + p * 3;
        ox[0] = x[0]; ox[1] = x[1]; ox[2] = x[2];
        if (!any) { mask[p] = 0; continue; }
        float pr[3];
        project_point(v_h, v_anchor, x, pr);
        const float vp[3] = {pr[0] - x[0], pr[1] - x[1], pr[2] - x[2]};
        const float dist = sqrtf(vp[0] * vp[0] + vp[1] * vp[1] + vp[2] * vp[2]);
        const float os = dist / len_h;
        const float po[3] = {pr[0] - os * v_offset[0], pr[1] - os * v_offset[1], pr[2] - os * v_offset[2]};
        const float q[3] = {po[0] - v_anchor[0], po[1] - v_anchor[1], po[2] - v_anchor[2]};
        const float pad = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
        const int cone = (pad <= radius) && (dist / (radius - pad) < len_h / radius * 1.1f);
        const int side = (vp[0] * v_h[0] + vp[1] * v_h[1] + vp[2] * v_h[2]) > 0.0f;
        const int valid = cone && side;
        mask[p] = (uint8_t)valid;
        if (!valid) continue;
        const float f = -((len_h - dist) / 10.0f);
        for (int i = 0; i < 3; i++) {
            const float vm = f * v_h[i] / len_h;
            const float mp = po[i] - vm;
            ox[i] = (mp - v_anchor[i]) * scale[i] + v_anchor[i];
        }
    }
}

/* seal_utils.py:58-79 texture branch of SealMapper.map_color: pixel = floor(<proj - o, w - o> / |w - o|^2 * W) clamped,
 * colour = modify_rgb(colors, image[pixel]) (per-sample target; V re-lit around the mean V of the batch),
 * blended with image_mask[pixel].  image [H,W,3], image_mask [H,W]. */
ORC_API void orc_seal_map_color_image(const float *points, const float *rgb, int64_t P, const float *image, const float *image_mask,
                                      uint32_t H, uint32_t W, const float *v_norm, const float *v_o, const float *v_w,
                                      const float *v_hh, float light_offset, float *out) {
    if (P == 0) return;
    const float ow[3] = {v_w[0] - v_o[0], v_w[1] - v_o[1], v_w[2] - v_o[2]};
    const float oh[3] = {v_hh[0] - v_o[0], v_hh[1] - v_o[1], v_hh[2] - v_o[2]};
    const float low = sqrtf(ow[0] * ow[0] + ow[1] * ow[1] + ow[2] * ow[2]), loh = sqrtf(oh[0] * oh[0] + oh[1] * oh[1] + oh[2] * oh[2]);
    double sum = 0;
    for (int64_t p = 0; p < P; p++) { float hsv[3]; rgb2hsv(rgb + p * 3, hsv); sum += hsv[2]; }
    const float mean = (float)(sum / (double)P);
    for (int64_t p = 0; p < P; p++) {
        float pr[3];
        project_point(v_norm, v_o, points + p * 3, pr);
        const float op[3] = {pr[0] - v_o[0], pr[1] - v_o[1], pr[2] - v_o[2]};
        float fw = floorf((op[0] * ow[0] + op[1] * ow[1] + op[2] * ow[2]) / (low * low) * (float)W);
        float fh = floorf((op[0] * oh[0] + op[1] * oh[1] + op[2] * oh[2]) / (loh * loh) * (float)H);
        fw = fminf(fmaxf(0.0f, fw), (float)(W - 1));
        fh = fminf(fmaxf(0.0f, fh), (float)(H - 1));
        const size_t pix = (size_t)fh * W + (size_t)fw;
        float hsv[3], thsv[3], mod[3];
        rgb2hsv(rgb + p * 3, hsv);
        rgb2hsv(image + pix * 3, thsv);
        const float nh[3] = {thsv[0], thsv[1], fminf(1.0f, fmaxf(0.0f, thsv[2] + (hsv[2] - mean) + light_offset))};
        hsv2rgb(nh, mod);
        const float a = image_mask[pix];
        for (int i = 0; i < 3; i++) out[p * 3 + i] = a * mod[i] + (1.0f - a) * rgb[p * 3 + i];
    }
}

/* ------------------------------------------------------------------------------------ */
/* tensoRF/network.py -- vector-matrix (VM) decomposition lookups                         */
/* ------------------------------------------------------------------------------------ */
/*
 * The reference has no kernel of its own here: get_sigma_feat / get_color_feat (tensoRF/network.py:99-151)
 * call torch.nn.functional.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) twelve
 * times per query batch.  The arithmetic is ATen's grid_sampler_2d (third-party: torch, unpinned in the
 * reference's requirements.txt:6; restated from its published algorithm, aten/src/ATen/native/GridSampler.h
 * grid_sampler_unnormalize + GridSampler.cu grid_sampler_2d_kernel):
 *     ix = ((x + 1) / 2) * (W - 1);  iy likewise with H
 *     nw = (floor ix, floor iy), ne = nw + (1,0), sw = nw + (0,1), se = nw + (1,1)
 *     w_nw = (ix_se - ix)(iy_se - iy), w_ne = (ix - ix_sw)(iy_sw - iy), w_sw = (ix_ne - ix)(iy - iy_ne), w_se = (ix - ix_nw)(iy - iy_nw)
 *     out = sum over the in-bounds taps, in the order nw, ne, sw, se
 * A "line" is a [R, D, 1] image sampled at x = 0 (network.py:106-107): ix = 0, so only the nw / sw taps of column 0
 * are in bounds and the weights reduce to (iy_se - iy) and (iy - iy_ne).
 * Pinned by tests/golden/cpu_tensorf.npz = the reference's own NeRFNetwork methods run on CPU torch.
 *
 * Layout here is the reference's: planes [R, H, W], lines [R, D].  Plane i uses coordinates
 * (x[mat_ids[i][0]] -> W, x[mat_ids[i][1]] -> H), mat_ids = {0,1},{0,2},{1,2}; line i uses x[vec_ids[i]], vec_ids = {2,1,0}
 * (network.py:37-38).  dims[i*3 + {0,1,2}] = H_i, W_i, D_i.
 */
static const int kMatId0[3] = {0, 0, 1}, kMatId1[3] = {1, 2, 2}, kVecId[3] = {2, 1, 0};

typedef struct { int x0, y0; float w[4]; int ok[4]; } vm_taps2;   /* nw, ne, sw, se */

static inline void vm_plane_taps(float gx, float gy, int H, int W, vm_taps2 *t) {
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1), iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
    const float fx = floorf(ix), fy = floorf(iy);
    const float ix_nw = fx, iy_nw = fy, ix_se = fx + 1.0f, iy_se = fy + 1.0f;
    t->x0 = (int)fx; t->y0 = (int)fy;
    t->w[0] = (ix_se - ix) * (iy_se - iy);
    t->w[1] = (ix - ix_nw) * (iy_se - iy);
    t->w[2] = (ix_se - ix) * (iy - iy_nw);
    t->w[3] = (ix - ix_nw) * (iy - iy_nw);
    for (int k = 0; k < 4; k++) {
        const int xx = t->x0 + (k & 1), yy = t->y0 + (k >> 1);
        t->ok[k] = xx >= 0 && xx < W && yy >= 0 && yy < H;
    }
}

/* aabb normalisation of tensoRF/network.py:158: x = 2 * (x - lo) / (hi - lo) - 1, elementwise float32 ops */
static inline float vm_normalise(float x, float lo, float hi) { return (2.0f * (x - lo)) / (hi - lo) - 1.0f; }

/* out_feat: [M, 3R] = mat_feat * vec_feat (plane-major, channel-minor: network.py:141-146) when reduce == 0,
 * or [M] = sum over planes of sum over channels (network.py:118-121) when reduce != 0.
 * xyz are world coordinates when aabb != NULL (normalised here), already-normalised coordinates otherwise. */
ORC_API void orc_vm_forward(const float *xyz, uint32_t M, const float *aabb, const float *const *mats, const float *const *vecs,
                            const int *dims, uint32_t R, int reduce, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < (int64_t)M; m++) {
        float x[3];
        for (int d = 0; d < 3; d++) x[d] = aabb ? vm_normalise(xyz[m * 3 + d], aabb[d], aabb[3 + d]) : xyz[m * 3 + d];
        float total = 0.0f;
        for (int i = 0; i < 3; i++) {
            const int H = dims[i * 3], W = dims[i * 3 + 1], D = dims[i * 3 + 2];
            vm_taps2 tp, tl;
            vm_plane_taps(x[kMatId0[i]], x[kMatId1[i]], H, W, &tp);
            vm_plane_taps(0.0f, x[kVecId[i]], D, 1, &tl);
            float plane_sum = 0.0f;
            for (uint32_t r = 0; r < R; r++) {
                const float *pm = mats[i] + (size_t)r * H * W, *pv = vecs[i] + (size_t)r * D;
                float a = 0.0f, b = 0.0f;
                for (int k = 0; k < 4; k++)
                    if (tp.ok[k]) a = fmaf(pm[(size_t)(tp.y0 + (k >> 1)) * W + tp.x0 + (k & 1)], tp.w[k], a);
                for (int k = 0; k < 4; k++)
                    if (tl.ok[k]) b = fmaf(pv[tl.y0 + (k >> 1)], tl.w[k], b);
                if (reduce) plane_sum += a * b;
                else out[(size_t)m * 3 * R + (size_t)i * R + r] = a * b;
            }
            total += plane_sum;
        }
        if (reduce) out[m] = total;
    }
}

/* gradients of the planes / lines for d(out) = g: g [M] (reduce) or [M, 3R].  Accumulated in double (the reference's
 * grid_sampler backward uses float atomics in a nondeterministic order), rounded once.  g_mats / g_vecs are overwritten. */
ORC_API void orc_vm_backward(const float *xyz, uint32_t M, const float *aabb, const float *const *mats, const float *const *vecs,
                             const int *dims, uint32_t R, int reduce, const float *g, float *const *g_mats, float *const *g_vecs) {
    for (int i = 0; i < 3; i++) {
        const int H = dims[i * 3], W = dims[i * 3 + 1], D = dims[i * 3 + 2];
        double *am = (double *)calloc((size_t)R * H * W, sizeof(double)), *av = (double *)calloc((size_t)R * D, sizeof(double));
        for (int64_t m = 0; m < (int64_t)M; m++) {
            float x[3];
            for (int d = 0; d < 3; d++) x[d] = aabb ? vm_normalise(xyz[m * 3 + d], aabb[d], aabb[3 + d]) : xyz[m * 3 + d];
            vm_taps2 tp, tl;
            vm_plane_taps(x[kMatId0[i]], x[kMatId1[i]], H, W, &tp);
            vm_plane_taps(0.0f, x[kVecId[i]], D, 1, &tl);
            for (uint32_t r = 0; r < R; r++) {
                const float *pm = mats[i] + (size_t)r * H * W, *pv = vecs[i] + (size_t)r * D;
                double a = 0.0, b = 0.0;
                for (int k = 0; k < 4; k++)
                    if (tp.ok[k]) a += (double)pm[(size_t)(tp.y0 + (k >> 1)) * W + tp.x0 + (k & 1)] * tp.w[k];
                for (int k = 0; k < 4; k++)
                    if (tl.ok[k]) b += (double)pv[tl.y0 + (k >> 1)] * tl.w[k];
                const double go = reduce ? (double)g[m] : (double)g[(size_t)m * 3 * R + (size_t)i * R + r];
                for (int k = 0; k < 4; k++)
                    if (tp.ok[k]) am[(size_t)r * H * W + (size_t)(tp.y0 + (k >> 1)) * W + tp.x0 + (k & 1)] += go * b * tp.w[k];
                for (int k = 0; k < 4; k++)
                    if (tl.ok[k]) av[(size_t)r * D + tl.y0 + (k >> 1)] += go * a * tl.w[k];
            }
        }
        for (size_t j = 0; j < (size_t)R * H * W; j++) g_mats[i][j] = (float)am[j];
        for (size_t j = 0; j < (size_t)R * D; j++) g_vecs[i][j] = (float)av[j];
        free(am); free(av);
    }
}

/* ------------------------------------------------------------------------------------ */
/* nerf/renderer.py:379-443 mark_untrained_grid                                           */
/* ------------------------------------------------------------------------------------ */
/* count[cas, morton(cell)] = number of cameras that see the cell centre: cam = (x - t) . R (poses cam2world, row vector
 * times R), in front (z > 0) and |cam.x| < cx/fx * z + 2*hgs, |cam.y| < cy/fy * z + 2*hgs; centre = (2c/(H-1) - 1)(bound - hgs)
 * in float32 like the reference's tensor expressions.  The caller sets density_grid = -1 where count == 0.
 * Pinned by tests/golden/cpu_untrained.npz (the reference method run on CPU torch). */
ORC_API void orc_mark_untrained_count(const float *poses, uint32_t B, float kx, float ky, uint32_t C, uint32_t H, float bound, int32_t *count) {
    const int64_t H3 = (int64_t)H * H * H;
    for (uint32_t cas = 0; cas < C; cas++) {
        const float bound_cas = fminf((float)(1u << cas), bound);
        const float hgs = bound_cas / (float)H, sc = bound_cas - hgs, margin = hgs * 2.0f;
#pragma omp parallel for schedule(static)
        for (int64_t cell = 0; cell < H3; cell++) {
            const uint32_t c3[3] = {(uint32_t)(cell / ((int64_t)H * H)), (uint32_t)((cell / H) % H), (uint32_t)(cell % H)};
            float w[3];
            for (int d = 0; d < 3; d++) w[d] = ((2.0f * (float)c3[d]) / (float)(H - 1) - 1.0f) * sc;
            int n = 0;
            for (uint32_t b = 0; b < B; b++) {
                const float *P = poses + (size_t)b * 16;
                const float dx = w[0] - P[3], dy = w[1] - P[7], dz = w[2] - P[11];
                const float cx_ = (dx * P[0] + dy * P[4]) + dz * P[8];
                const float cy_ = (dx * P[1] + dy * P[5]) + dz * P[9];
                const float cz_ = (dx * P[2] + dy * P[6]) + dz * P[10];
                n += (cz_ > 0.0f && fabsf(cx_) < kx * cz_ + margin && fabsf(cy_) < ky * cz_ + margin) ? 1 : 0;
            }
            count[(int64_t)cas * H3 + morton3(c3[0], c3[1], c3[2])] = n;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* nerf/utils.py:53-140 get_rays                                                          */
/* ------------------------------------------------------------------------------------ */
/* pixel id = row * W + col -> i = col + 0.5, j = row + 0.5 (:64-66), d = ((i-cx)/fx, (j-cy)/fy, 1) / |.| (:123-128),
 * rays_d = d @ R^T, rays_o = t (:129-132).  inds: int64 [inds_rows, N] (inds_rows 1 or B) or NULL = all pixels.
 * Pinned by tests/golden/cpu_get_rays.npz (the reference function run on CPU torch). */
ORC_API void orc_get_rays(const float *poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t W, const int64_t *inds,
                          uint32_t inds_rows, uint32_t N, float *rays_o, float *rays_d) {
    for (uint32_t b = 0; b < B; b++) {
        const float *P = poses + (size_t)b * 16;
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < (int64_t)N; n++) {
            const int64_t pix = inds ? inds[(size_t)(inds_rows == 1 ? 0 : b) * N + n] : n;
            const float i = (float)(pix % W) + 0.5f, j = (float)(pix / W) + 0.5f;
            const float x = (i - cx) / fx, y = (j - cy) / fy;
            const float norm = sqrtf((x * x + y * y) + 1.0f);
            const float d[3] = {x / norm, y / norm, 1.0f / norm};
            float *o = rays_o + ((size_t)b * N + n) * 3, *r = rays_d + ((size_t)b * N + n) * 3;
            for (int k = 0; k < 3; k++) {
                r[k] = (d[0] * P[k * 4] + d[1] * P[k * 4 + 1]) + d[2] * P[k * 4 + 2];
                o[k] = P[k * 4 + 3];
            }
        }
    }
}
