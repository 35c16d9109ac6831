// Fused NGP field for the distillation trainer (sm_100a): samples -> (sigma, rgb) and back.
//
// The reference evaluates nerf/network.py:99-128 as ~20 launches per call (2 grid encodes with a full
// table cast + permute copies, SH, a concat, 5 cuBLAS GEMMs, ~8 elementwise kernels) and writes every
// intermediate to HBM.  Here the field is four kernels with one 128-byte row per sample between them:
//
//   k_ngp_encode    xyz -> feats[M,64] fp16 = [sigma-grid 32 | colour-grid 32].  Both hash tables live in ONE
//                   interleaved fp16 table (entry = {s0,s1,c0,c1}, 8 bytes): a corner costs one 64-bit load for
//                   both encoders instead of two 32-bit loads from two tables (the cell geometry is shared).
//   k_ngp_mlp_fwd   feats + dirs -> sigma, rgb.  All five GEMMs on tcgen05 (M=128 samples per tile, K=16 per MMA,
//                   fp32 accumulators in TMEM), SH(4) evaluated in registers and written straight into the operand
//                   tile; activations never leave the SM.  The colour layer reads its 63-wide input from two
//                   tiles through per-MMA descriptors ([colour feats] from the feature tile, [SH | geo] from the
//                   scratch tile), so nothing is concatenated.
//   k_ngp_mlp_bwd   recomputes the forward from feats (cheaper than storing 3x64 activations per sample), then
//                   data gradients (weights read MN-major = transposed for free) and weight gradients (both
//                   operands MN-major, K = the 128 samples of the tile) as tcgen05 MMAs; weight gradients
//                   accumulate in TMEM across all tiles of the persistent CTA and are flushed once.  Two chains are
//                   in flight per SM (backward of tile t, forward recompute of tile t+1), each with its own
//                   epilogue warps, to cover the per-round tensor-core round trip.
//   k_ngp_scatter   xyz + dfeats -> interleaved fp32 gradient table with the in-warp segmented pre-reduction
//                   (ray-major samples share cells at coarse levels) and one 128-bit RED per corner.
//
// Level geometry follows gridencoder.cu:137-156 of the reference (D=3, linear interpolation,
// align_corners=false); MLP semantics follow nerf/network.py:99-128 and activation.py:5-17.
#include "tc05.cuh"
#include <cstdlib>

namespace {

using namespace tc05;

constexpr uint32_t kMaxLevels = 16;
constexpr uint32_t kRows = 128;
constexpr uint32_t kTileBytes = kRows * 128;  // [128 x 64] fp16
constexpr uint32_t kWTile = 64 * 128;         // [64 x 64] fp16
constexpr uint32_t kOTile = 16 * 128;         // [16 x 64] fp16

struct Geo {
    float scale[kMaxLevels];
    uint32_t res1[kMaxLevels];    // resolution + 1 (dense stride)
    uint32_t size[kMaxLevels];    // hashmap_size
    uint32_t offset[kMaxLevels];  // first entry of the level
    uint32_t hashed;              // bit l: level l uses the xor-prime hash
    uint32_t pow2;                // bit l: size is a power of two
    uint32_t pair_ok;             // bit l: offset and size are even, so aligned entry pairs stay inside the level
};

__device__ void geo_init(Geo &g, const int *__restrict__ offsets, uint32_t L, float S, uint32_t H) {
    const uint32_t l = threadIdx.x;
    if (l == 0) { g.hashed = 0; g.pow2 = 0; g.pair_ok = 0; }
    __syncthreads();
    if (l < L) {
        const uint32_t size = (uint32_t)(offsets[l + 1] - offsets[l]);
        const float scale = __fmaf_rn(exp2f((float)l * S), (float)H, -1.0f);
        const uint32_t res1 = (uint32_t)ceilf(scale) + 2;  // resolution + 1
        g.scale[l] = scale; g.res1[l] = res1; g.size[l] = size; g.offset[l] = (uint32_t)offsets[l];
        uint32_t stride = 1;
#pragma unroll
        for (int d = 0; d < 3; d++) if (stride <= size) stride *= res1;
        if (stride > size) atomicOr(&g.hashed, 1u << l);
        if ((size & (size - 1)) == 0) atomicOr(&g.pow2, 1u << l);
        if ((((uint32_t)offsets[l] | size) & 1u) == 0) atomicOr(&g.pair_ok, 1u << l);
    }
    __syncthreads();
}

struct Cell {
    uint32_t idx[8];
    float w[8];
};

// corner indices (entry units, level offset included) and trilinear weights of u in level l
__device__ __forceinline__ void locate(const Geo &g, uint32_t l, float ux, float uy, float uz, Cell &c, unsigned long long *key) {
    const float s = g.scale[l];
    const float px = __fmaf_rn(ux, s, 0.5f), py = __fmaf_rn(uy, s, 0.5f), pz = __fmaf_rn(uz, s, 0.5f);
    const float fx0 = floorf(px), fy0 = floorf(py), fz0 = floorf(pz);
    const uint32_t gx = (uint32_t)fx0, gy = (uint32_t)fy0, gz = (uint32_t)fz0;
    const float fx = __fsub_rn(px, fx0), fy = __fsub_rn(py, fy0), fz = __fsub_rn(pz, fz0);
    if (key) *key = (unsigned long long)gx | ((unsigned long long)gy << 21) | ((unsigned long long)gz << 42);
    const uint32_t size = g.size[l], off = g.offset[l];
    const bool hashed = (g.hashed >> l) & 1u, p2 = (g.pow2 >> l) & 1u;
    const uint32_t r1 = g.res1[l];
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) {
        const uint32_t x = gx + (i & 1u), y = gy + ((i >> 1) & 1u), z = gz + (i >> 2);
        uint32_t id;
        if (hashed) {
            id = x ^ (y * 2654435761u) ^ (z * 805459861u);
            id = p2 ? (id & (size - 1)) : (id % size);
        } else {
            id = x + y * r1;
            if (r1 * r1 <= size) id += z * r1 * r1;   // get_grid_index stops adding once the stride exceeds the table
            id = id % size;
        }
        c.idx[i] = off + id;
        // same product order as the reference: ((1 * wx) * wy) * wz
        float w = (i & 1u) ? fx : __fsub_rn(1.0f, fx);
        w = __fmul_rn(w, ((i >> 1) & 1u) ? fy : __fsub_rn(1.0f, fy));
        w = __fmul_rn(w, (i >> 2) ? fz : __fsub_rn(1.0f, fz));
        c.w[i] = w;
    }
}

__device__ __forceinline__ bool to_unit(float x, float y, float z, float bound, float &ux, float &uy, float &uz) {
    const float inv = 2.0f * bound;  // GridEncoder.forward: (inputs + bound) / (2 * bound)
    ux = __fdiv_rn(__fadd_rn(x, bound), inv); uy = __fdiv_rn(__fadd_rn(y, bound), inv); uz = __fdiv_rn(__fadd_rn(z, bound), inv);
    return !(ux < 0.0f || ux > 1.0f || uy < 0.0f || uy > 1.0f || uz < 0.0f || uz > 1.0f);
}

// ------------------------------------------------------------------------------------------------
// encode: one thread per sample, both tables, all levels
// ------------------------------------------------------------------------------------------------
// table entries are {s0,s1,c0,c1} fp16 (8 bytes) at  table + idx * stride + off : stride 8 / off 0 for a stand-alone
// model, stride 16 / off 0|8 for the teacher | student half of a paired table (see k_ngp_encode_pair)
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

__device__ __forceinline__ uint2 ld_entry(const uint8_t *__restrict__ table, uint32_t idx, uint32_t stride) {
    return __ldg(reinterpret_cast<const uint2 *>(table + (size_t)idx * stride));
}
__device__ __forceinline__ bool load_unit(const float *__restrict__ xyz, uint32_t i, float bound, float &ux, float &uy, float &uz) {
    return to_unit(xyz[(size_t)i * 3], xyz[(size_t)i * 3 + 1], xyz[(size_t)i * 3 + 2], bound, ux, uy, uz);
}

// The two x-corners of a cell, idx[2j] and idx[2j+1], are the halves of one ALIGNED entry pair whenever they differ in
// bit 0 only (even x on hashed levels, even dense index otherwise): one double-width load then serves both.  The L1 tag
// stage pays per (lane, sector), so this removes a quarter of the gather cost on average.  All loads are issued before any
// use (the unmerged second load is predicated, not branched) so the 8 gathers of a level stay in flight together.
__device__ __forceinline__ void gather_cell_e8(const uint8_t *__restrict__ table, const Cell &c, uint2 (&v)[8]) {
    uint4 a[4];
    bool merged[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        merged[j] = (c.idx[2 * j] ^ c.idx[2 * j + 1]) == 1u;
        a[j] = __ldg(reinterpret_cast<const uint4 *>(table) + (c.idx[2 * j] >> 1));
        if (!merged[j]) v[2 * j + 1] = __ldg(reinterpret_cast<const uint2 *>(table) + c.idx[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool odd = c.idx[2 * j] & 1u;
        const uint2 lo = make_uint2(a[j].x, a[j].y), hi = make_uint2(a[j].z, a[j].w);
        v[2 * j] = odd ? hi : lo;
        if (merged[j]) v[2 * j + 1] = odd ? lo : hi;
    }
}

__global__ void __launch_bounds__(256)
k_ngp_encode(const float *__restrict__ xyz, uint32_t M, float bound, const uint8_t *__restrict__ table, uint32_t stride,
             const int *__restrict__ offsets, uint32_t L, float S, uint32_t H, __half *__restrict__ feats, int sigma_only) {
    __shared__ Geo g;
    geo_init(g, offsets, L, S, H);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float ux, uy, uz;
    const bool ok = load_unit(xyz, i, bound, ux, uy, uz);
    uint4 *row = reinterpret_cast<uint4 *>(feats + (size_t)i * 64);
    for (uint32_t grp = 0; grp < 4; grp++) {      // 4 levels -> one 16-byte chunk per table
        uint32_t fs[4], fc[4];   // one packed half2 per level: 4 levels -> one 16-byte chunk per table
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) {
            const uint32_t l = grp * 4 + q;
            float s0 = 0.f, s1 = 0.f, c0 = 0.f, c1 = 0.f;
            if (ok && l < L) {
                Cell c;
                locate(g, l, ux, uy, uz, c, nullptr);
                uint2 v[8];
                if (stride == 8 && ((g.pair_ok >> l) & 1u)) gather_cell_e8(table, c, v);
                else {
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = ld_entry(table, c.idx[k], stride);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v[k].x));
                    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v[k].y));
                    s0 = __fmaf_rn(c.w[k], a.x, s0); s1 = __fmaf_rn(c.w[k], a.y, s1);
                    c0 = __fmaf_rn(c.w[k], b.x, c0); c1 = __fmaf_rn(c.w[k], b.y, c1);
                }
            }
            fs[q] = pack2(s0, s1); fc[q] = pack2(c0, c1);
        }
        row[grp] = make_uint4(fs[0], fs[1], fs[2], fs[3]);
        if (!sigma_only) row[4 + grp] = make_uint4(fc[0], fc[1], fc[2], fc[3]);
    }
}

// Teacher and student on the same sample: ONE 128-bit load per corner from a paired table
// {teacher s0,s1,c0,c1 | student s0,s1,c0,c1} feeds both feature rows, because both models share the level geometry and
// (except for the ~1 % of samples the proxy mapping moved) the query point.  Moved samples (mask != 0) take a second
// gather for the teacher at its mapped position.
__device__ __forceinline__ void acc4(const uint2 v, float w, float &a0, float &a1, float &a2, float &a3) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
    a0 = __fmaf_rn(w, a.x, a0); a1 = __fmaf_rn(w, a.y, a1); a2 = __fmaf_rn(w, b.x, a2); a3 = __fmaf_rn(w, b.y, a3);
}

__global__ void __launch_bounds__(256, 4)
k_ngp_encode_pair(const float *__restrict__ xyz, const float *__restrict__ xyz_teacher, const uint8_t *__restrict__ mask, uint32_t M,
                  float bound, const uint4 *__restrict__ table8, const int *__restrict__ offsets, uint32_t L, float S, uint32_t H,
                  __half *__restrict__ feats_teacher, __half *__restrict__ feats_student) {
    __shared__ Geo g;
    geo_init(g, offsets, L, S, H);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    float ux, uy, uz, tx = 0, ty = 0, tz = 0;
    const bool ok = load_unit(xyz, i, bound, ux, uy, uz);
    const bool moved = mask && mask[i];
    bool tok = ok;
    if (moved) tok = load_unit(xyz_teacher, i, bound, tx, ty, tz);
    uint4 *row_t = reinterpret_cast<uint4 *>(feats_teacher + (size_t)i * 64), *row_s = reinterpret_cast<uint4 *>(feats_student + (size_t)i * 64);
    for (uint32_t grp = 0; grp < 4; grp++) {
        uint32_t ts[4], tc[4], ss[4], sc[4];
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) {
            const uint32_t l = grp * 4 + q;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            if (ok && l < L) {
                Cell c;
                locate(g, l, ux, uy, uz, c, nullptr);
                uint4 v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = __ldg(table8 + c.idx[k]);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    acc4(make_uint2(v[k].z, v[k].w), c.w[k], s0, s1, s2, s3);
                    if (!moved) acc4(make_uint2(v[k].x, v[k].y), c.w[k], t0, t1, t2, t3);
                }
            }
            if (moved && tok && l < L) {
                Cell c;
                locate(g, l, tx, ty, tz, c, nullptr);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint4 v = __ldg(table8 + c.idx[k]);
                    acc4(make_uint2(v.x, v.y), c.w[k], t0, t1, t2, t3);
                }
            }
            ts[q] = pack2(t0, t1); tc[q] = pack2(t2, t3); ss[q] = pack2(s0, s1); sc[q] = pack2(s2, s3);
        }
        row_t[grp] = make_uint4(ts[0], ts[1], ts[2], ts[3]); row_t[4 + grp] = make_uint4(tc[0], tc[1], tc[2], tc[3]);
        row_s[grp] = make_uint4(ss[0], ss[1], ss[2], ss[3]); row_s[4 + grp] = make_uint4(sc[0], sc[1], sc[2], sc[3]);
    }
}

__global__ void k_pair_tables(const uint2 *__restrict__ teacher4, const uint2 *__restrict__ student4, uint4 *__restrict__ table8, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 t = teacher4[i], s = student4[i];
    table8[i] = make_uint4(t.x, t.y, s.x, s.y);
}

// ------------------------------------------------------------------------------------------------
// scatter: dfeats -> interleaved fp32 gradient table
// ------------------------------------------------------------------------------------------------
// one level of one sample's gradient: in-warp segmented pre-reduction over runs of lanes that sit in the same cell (ray-major
// samples share cells at the coarse levels), then one 128-bit RED per corner of every run.  Called by whole warps.
// Fixed-point accumulation (the deterministic mode): a contribution is rounded once to a multiple of 2^-32 and added as a 64-bit
// integer -- integer addition is associative, so the table no longer depends on the order in which the SMs' reductions reach L2
// and two runs of the same step are bit-identical.  Range: |sum| < 2^31 per value; non-finite contributions (an fp16 overflow the
// GradScaler must see) raise a flag word that k_fixed_to_float turns back into a NaN.
constexpr float kFixedOne = 4294967296.0f;   // 2^32
__device__ __forceinline__ void fixed_add(long long *__restrict__ p, float v, uint32_t *__restrict__ nonfinite) {
    if (!(fabsf(v) < 2147483648.0f)) { atomicOr(nonfinite, 1u); return; }   // inf, NaN or out of range
    const long long q = __float2ll_rn(v * kFixedOne);
    if (q != 0) atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)q);
}

template <bool FIXED>
__device__ __forceinline__ void red_entry(float4 *__restrict__ grad4, uint32_t idx, float a0, float a1, float a2, float a3, uint32_t *__restrict__ nonfinite) {
    if constexpr (FIXED) {
        long long *e = reinterpret_cast<long long *>(grad4) + (size_t)idx * 4;
        fixed_add(e, a0, nonfinite); fixed_add(e + 1, a1, nonfinite); fixed_add(e + 2, a2, nonfinite); fixed_add(e + 3, a3, nonfinite);
    } else {
        atomicAdd(grad4 + idx, make_float4(a0, a1, a2, a3));
    }
}

template <bool FIXED = false>
__device__ __forceinline__ void scatter_level(const Geo &g, uint32_t l, bool ok, float ux, float uy, float uz, float g0, float g1, float g2, float g3,
                                              uint32_t lane, float4 *__restrict__ grad4, uint32_t *__restrict__ nonfinite = nullptr) {
    Cell c;
    unsigned long long key = ~0ull;
    if (ok) locate(g, l, ux, uy, uz, c, &key);
    else {
#pragma unroll
        for (int k = 0; k < 8; k++) { c.idx[k] = 0; c.w[k] = 0.f; }
    }
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (__popc(heads) <= 16) {
        // runs of equal cells: fold every run into its head lane (segmented suffix sums), then one RED per corner
        const uint32_t after = heads >> 1 >> lane;
        const uint32_t seg_left = after ? (uint32_t)(__ffs(after) - 1) : (31u - lane);
        // number of doubling steps the longest run of this warp needs (a run of n lanes needs ceil(log2 n))
        const uint32_t cont = ~heads;                 // bit i: lane i continues the run of lane i-1
        const uint32_t c2 = cont & (cont >> 1);       // some run longer than 2
        const uint32_t c4 = c2 & (c2 >> 2);           // >= 4 consecutive continuation bits: longer than 4
        const uint32_t c8 = c4 & (c4 >> 4);           // >= 8 consecutive: longer than 8
        const uint32_t c16 = c8 & (c8 >> 8);          // longer than 16
        const uint32_t nsteps = c16 ? 5 : (c8 ? 4 : (c4 ? 3 : (c2 ? 2 : 1)));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float a0 = c.w[k] * g0, a1 = c.w[k] * g1, a2 = c.w[k] * g2, a3 = c.w[k] * g3;
#pragma unroll
            for (uint32_t st = 0; st < 5; st++) {
                if (st < nsteps) {
                    const uint32_t d = 1u << st;
                    const float o0 = __shfl_down_sync(0xffffffffu, a0, d), o1 = __shfl_down_sync(0xffffffffu, a1, d);
                    const float o2 = __shfl_down_sync(0xffffffffu, a2, d), o3 = __shfl_down_sync(0xffffffffu, a3, d);
                    if (d <= seg_left) { a0 += o0; a1 += o1; a2 += o2; a3 += o3; }
                }
            }
            if (ok && head) red_entry<FIXED>(grad4, c.idx[k], a0, a1, a2, a3, nonfinite);
        }
    } else if (ok) {
#pragma unroll
        for (int k = 0; k < 8; k++) red_entry<FIXED>(grad4, c.idx[k], c.w[k] * g0, c.w[k] * g1, c.w[k] * g2, c.w[k] * g3, nonfinite);
    }
}

template <bool FIXED>
__global__ void __launch_bounds__(256)
k_ngp_scatter(const float *__restrict__ xyz, const __half *__restrict__ dfeats, uint32_t M, float bound,
              float4 *__restrict__ grad4, const int *__restrict__ offsets, uint32_t L, float S, uint32_t H, float grad_scale,
              uint32_t grp_begin, uint32_t grp_end, uint32_t *__restrict__ nonfinite) {
    __shared__ Geo g;
    geo_init(g, offsets, L, S, H);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // whole warps stay alive
    float ux = 0, uy = 0, uz = 0;
    bool ok = false;
    if (i < M) ok = load_unit(xyz, i, bound, ux, uy, uz);
    const uint32_t lane = lane_id();
    const uint4 *row = reinterpret_cast<const uint4 *>(dfeats + (size_t)(i < M ? i : 0) * 64);
    const bool wide = aligned32(row);
    // gradient rows of an aligned PAIR of groups: one 256-bit load per table (row-per-thread accesses cost 32 L1 wavefronts per
    // instruction whatever their width)
    uint4 rs_lo = make_uint4(0, 0, 0, 0), rs_hi = rs_lo, rc_lo = rs_lo, rc_hi = rs_lo;
    for (uint32_t grp = grp_begin; grp < grp_end; grp++) {   // 4 levels per group; a launch may cover a sub-range (s3d_ngp_scatter_levels)
        // gradient rows of the group's 4 levels: one half2 per (level, table), kept packed -- the level loop is NOT unrolled:
        // unrolled, the kernel is ~3900 instructions (63 KB) and ncu shows 16 % of its stall samples on instruction fetch
        // (2.24 ms per 5.35 M samples); with one level per iteration it is ~1250 instructions and takes 2.03 ms
        const bool odd = grp & 1u;
        if (ok) {
            if (wide && !odd && grp + 1 < grp_end) { ldg256(row + grp, rs_lo, rs_hi); ldg256(row + 4 + grp, rc_lo, rc_hi); }
            else if (!wide || !odd || grp == grp_begin) {
                const uint4 a = __ldg(row + grp), b = __ldg(row + 4 + grp);
                if (odd) { rs_hi = a; rc_hi = b; } else { rs_lo = a; rc_lo = b; }
            }
        }
        const uint4 rs = odd ? rs_hi : rs_lo, rc = odd ? rc_hi : rc_lo;
#pragma unroll 1
        for (uint32_t q = 0; q < 4; q++) {
            const uint32_t l = grp * 4 + q;
            if (l >= L) break;
            const uint32_t ps = q == 0 ? rs.x : (q == 1 ? rs.y : (q == 2 ? rs.z : rs.w));
            const uint32_t pc = q == 0 ? rc.x : (q == 1 ? rc.y : (q == 2 ? rc.z : rc.w));
            const float2 fs = __half22float2(*reinterpret_cast<const __half2 *>(&ps)), fc = __half22float2(*reinterpret_cast<const __half2 *>(&pc));
            const float g0 = fs.x * grad_scale, g1 = fs.y * grad_scale, g2 = fc.x * grad_scale, g3 = fc.y * grad_scale;
            scatter_level<FIXED>(g, l, ok, ux, uy, uz, g0, g1, g2, g3, lane, grad4, nonfinite);
        }
    }
}

// How many reductions k_ngp_scatter issues for a batch (same run detection, nothing written): the kernel's real unit of work.
// B200 retires ~149 G random global reductions per second whatever their width (scripts/r2/red_micro.cu), so
// reductions / time against that rate is the roofline that actually bounds the scatter.
__global__ void __launch_bounds__(256)
k_ngp_scatter_count(const float *__restrict__ xyz, uint32_t M, float bound, const int *__restrict__ offsets, uint32_t L, float S, uint32_t H,
                    unsigned long long *__restrict__ count) {
    __shared__ Geo g;
    geo_init(g, offsets, L, S, H);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float ux = 0, uy = 0, uz = 0;
    bool ok = false;
    if (i < M) ok = load_unit(xyz, i, bound, ux, uy, uz);
    const uint32_t lane = lane_id();
    uint32_t n = 0, nh = 0;      // all levels / hashed levels only (count[0], count[1])
    for (uint32_t l = 0; l < L; l++) {
        Cell c;
        unsigned long long key = ~0ull;
        if (ok) locate(g, l, ux, uy, uz, c, &key);
        const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (prev != key);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const uint32_t k = 8u * __popc(__ballot_sync(0xffffffffu, ok && (head || __popc(heads) > 16)));
        n += k;
        if ((g.hashed >> l) & 1u) nh += k;
    }
    if (lane == 0 && n) { atomicAdd(count, (unsigned long long)n); atomicAdd(count + 1, (unsigned long long)nh); }
}

// fixed-point arena -> fp32 gradient arena (accumulated into), arena cleared; *nonfinite -> NaN in grad[0]
__global__ void k_fixed_to_float(long long *__restrict__ fx, float *__restrict__ grad, size_t n, uint32_t *__restrict__ nonfinite) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long q = fx[i];
    if (q != 0) {
        grad[i] = __fadd_rn(grad[i], __double2float_rn(__dmul_rn(__ll2double_rn(q), 1.0 / 4294967296.0)));
        fx[i] = 0;
    }
    if (i == 0 && *nonfinite) { grad[0] = __int_as_float(0x7fc00000); *nonfinite = 0u; }
}

// ------------------------------------------------------------------------------------------------
// MLP: shared pieces
// ------------------------------------------------------------------------------------------------
struct MlpWeights {
    const __half *ws0, *ws1, *wc0, *wc1, *wc2;  // row-major [64,32] [16,64] [64,63] [64,64] [3,64] (nn.Linear layout)
};

// weight tiles in shared memory (SW128, K = 64 columns):
//   Ws0 [64 x 64]  cols 0-31 = W_s0, cols 32-63 = 0
//   Ws1 [16 x 64]
//   Wc0 [64 x 64]  cols 0-31 = colour-grid inputs (orig 31..62), 32-47 = SH (orig 0..15), 48-62 = geo (orig 16..30), 63 = 0
//   Wc1 [64 x 64]
//   Wc2 [16 x 64]  rows 0-2 = W_c2, rows 3-15 = 0
__device__ void load_weights(uint8_t *tWs0, uint8_t *tWs1, uint8_t *tWc0, uint8_t *tWc1, uint8_t *tWc2, const MlpWeights &w) {
    const __half zero = __float2half_rn(0.0f);
    for (uint32_t i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const uint32_t r = i >> 6, c = i & 63;
        const uint32_t off = sw128_off(r, c >> 3) + (c & 7) * 2;
        *reinterpret_cast<__half *>(tWs0 + off) = c < 32 ? w.ws0[r * 32 + c] : zero;
        const uint32_t oc = c < 32 ? 31 + c : (c < 48 ? c - 32 : (c < 63 ? c - 48 + 16 : 0xffffffffu));
        *reinterpret_cast<__half *>(tWc0 + off) = oc != 0xffffffffu ? w.wc0[r * 63 + oc] : zero;
        *reinterpret_cast<__half *>(tWc1 + off) = w.wc1[r * 64 + c];
        if (r < 16) {
            *reinterpret_cast<__half *>(tWs1 + off) = w.ws1[r * 64 + c];
            *reinterpret_cast<__half *>(tWc2 + off) = r < 3 ? w.wc2[r * 64 + c] : zero;
        }
    }
}

// real SH, degree 4 (16 values), same polynomials as shencoder.cu:49-68 of the reference, written through the
// (x + iy)^m factorisation: Y = N * Q_l^m(z) * {c_m | s_m}(x, y)
__device__ __forceinline__ void sh4(float x, float y, float z, float *o) {
    const float z2 = z * z;
    const float c1 = x, s1 = y;
    const float c2 = x * x - y * y, s2 = 2.0f * x * y;
    const float c3 = x * c2 - y * s2, s3 = x * s2 + y * c2;
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291992f * s1;
    o[2] = 0.48860251190291992f * z;
    o[3] = -0.48860251190291992f * c1;
    o[4] = 0.54627421529603959f * s2;
    o[5] = -1.0925484305920792f * z * s1;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * z * c1;
    o[8] = 0.54627421529603959f * c2;
    o[9] = -0.59004358992664352f * s3;
    o[10] = 1.4453057213202769f * z * s2;
    o[11] = 0.45704579946446572f * (1.0f - 5.0f * z2) * s1;
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * (1.0f - 5.0f * z2) * c1;
    o[14] = 1.4453057213202769f * z * c2;
    o[15] = -0.59004358992664352f * c3;
}

// Thread layout of the MLP kernels: 256 threads = 2 warpgroups; thread t owns sample row (t & 127) and column half
// hf = t >> 7 (columns 32*hf .. 32*hf+31) of every 64-wide tile / accumulator.  A warp may only read the 32 TMEM lanes
// of its sub-partition (warp % 4), which is exactly rows 32*(warp%4) .. +31 for both warpgroups.
// Inputs of one tile held in registers: the loads for tile t+1 are issued at the top of tile t, so their DRAM latency
// runs under the 5 / 11 tensor-core rounds of tile t instead of sitting at the head of every round that needs them.
struct FeatPre {
    uint4 f[4];
    __device__ __forceinline__ void load(const __half *__restrict__ feats, uint32_t row_g, uint32_t hf, bool in_range) {
        const uint4 *src = reinterpret_cast<const uint4 *>(feats + (size_t)row_g * 64) + hf * 4;
        // a thread's half row is 64 contiguous bytes: two 256-bit loads (a row-per-thread access costs one L1 wavefront per lane
        // whatever its width)
        if (in_range && aligned32(src)) { ldg256(src, f[0], f[1]); ldg256(src + 2, f[2], f[3]); return; }
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) f[q] = in_range ? __ldg(src + q) : make_uint4(0, 0, 0, 0);
    }
    __device__ __forceinline__ void store(uint8_t *tile, uint32_t r, uint32_t hf) const {
#pragma unroll
        for (uint32_t q = 0; q < 4; q++) *reinterpret_cast<uint4 *>(tile + sw128_off(r, hf * 4 + q)) = f[q];
    }
};

__device__ __forceinline__ void sync_tiles() {  // writers -> tensor core, tensor core results -> readers
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}

struct Issue {
    uint32_t mbar, parity;
    __device__ __forceinline__ void wait() {
        mbar_wait(mbar, parity);
        parity ^= 1;
        fence_after_sync();
    }
};

// write 32 fp32 values (one half of a 64-wide row) as fp16 into chunks [4*half, 4*half+4) of row r
__device__ __forceinline__ void store_half_row(uint8_t *tile, uint32_t r, uint32_t half, const float *v) {
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) *reinterpret_cast<uint4 *>(tile + sw128_off(r, half * 4 + q)) = pack8(v + q * 8);
}
__device__ __forceinline__ void zero_half_row(uint8_t *tile, uint32_t r, uint32_t half) {
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) *reinterpret_cast<uint4 *>(tile + sw128_off(r, half * 4 + q)) = make_uint4(0, 0, 0, 0);
}
// accumulator half -> ReLU -> operand tile
__device__ __forceinline__ void relu_to_tile(uint32_t t_row, uint32_t hf, uint8_t *tile, uint32_t r) {
    float v[32];
    tmem_ld32(t_row + hf * 32, v);
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = fmaxf(v[i], 0.0f);
    store_half_row(tile, r, hf, v);
}
// gradient half: D (.) relu'(act) -> gradient tile, in two phases: the masked, packed half row is built in registers first (while MMAs may still read the destination
// tile) and stored once the caller knows the tile is free
__device__ __forceinline__ void masked_half_row(uint32_t t_row, uint32_t hf, const uint8_t *act_tile, uint32_t r, uint4 (&out)[4]) {
    float v[32];
    tmem_ld32(t_row + hf * 32, v);
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) {
        float act[8];
        unpack8(*reinterpret_cast<const uint4 *>(act_tile + sw128_off(r, hf * 4 + q)), act);
#pragma unroll
        for (int i = 0; i < 8; i++) v[q * 8 + i] = act[i] > 0.0f ? v[q * 8 + i] : 0.0f;
        out[q] = pack8(v + q * 8);
    }
}
__device__ __forceinline__ void store_packed_half_row(uint8_t *tile, uint32_t r, uint32_t hf, const uint4 (&v)[4]) {
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) *reinterpret_cast<uint4 *>(tile + sw128_off(r, hf * 4 + q)) = v[q];
}

// ------------------------------------------------------------------------------------------------
// MLP kernels: 9 warps.  Warps 0-7 (two warpgroups) own the rows / column halves of the tiles and run the epilogues;
// warp 8 is the MMA issuer: it allocates TMEM, builds the descriptors and issues every tcgen05.mma, so that descriptor
// arithmetic and the (asynchronous) weight-gradient MMAs never sit on the critical path of the epilogue warps.
// Every round:   epilogue warps write the operand tile -> fence -> __syncthreads -> issuer issues + commits ->
//                epilogue warps wait on the mbarrier -> read the accumulator.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kMlpThreads = 288;
constexpr uint32_t kIssuerWarp = 8;

__device__ __forceinline__ void epi_publish() {   // epilogue warps: my tile writes are done
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
}
__device__ __forceinline__ void iss_acquire() {   // issuer warp: all tile writes of this round are visible
    __syncthreads();
    fence_after_sync();
}
__device__ __forceinline__ void tile_end_sync() {  // both roles: accumulator reads done, tiles may be overwritten
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}

struct FwdArgs {
    const __half *feats;
    const float *dirs;
    MlpWeights w;
    float *sigma, *rgb, *geo;  // geo [M,15] optional (density() callers)
    uint32_t M, n_tiles;
    float density_scale;
    int sigma_only;
};

__global__ void __launch_bounds__(kMlpThreads)
k_ngp_mlp_fwd(const FwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *tF = smem, *tT = tF + kTileBytes, *tWs0 = tT + kTileBytes, *tWc0 = tWs0 + kWTile, *tWc1 = tWc0 + kWTile,
            *tWs1 = tWc1 + kWTile, *tWc2 = tWs1 + kOTile;
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync(), r = tid & 127, hf = (tid >> 7) & 1u;
    const bool is_issuer = (warp == kIssuerWarp);
    if (is_issuer) tmem_alloc(smem_u32(&s_tmem), 64);
    if (tid == 0) mbar_init(smem_u32(&s_mbar), 1);
    load_weights(tWs0, tWs1, tWc0, tWc1, tWc2, a.w);
    sync_tiles();
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    Issue is{smem_u32(&s_mbar), 0};
    const uint32_t aF = smem_u32(tF), aT = smem_u32(tT);
    const uint32_t id64 = make_idesc(128, 64, false, false), id16 = make_idesc(128, 16, false, false);

    if (is_issuer) {
        const bool lead = elect_one();
        const uint32_t aWs0 = smem_u32(tWs0), aWs1 = smem_u32(tWs1), aWc0 = smem_u32(tWc0), aWc1 = smem_u32(tWc1), aWc2 = smem_u32(tWc2);
        for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            iss_acquire();   // features loaded
            { for (uint32_t k = 0; k < 2; k++) mma_f16_if(lead, tmem, desc_kmajor(aF, k), desc_kmajor(aWs0, k), id64, k > 0); mma_commit_if(lead, is.mbar); }
            iss_acquire();   // H1 written
            { for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, tmem, desc_kmajor(aT, k), desc_kmajor(aWs1, k), id16, k > 0); mma_commit_if(lead, is.mbar); }
            if (a.sigma_only) { tile_end_sync(); continue; }
            iss_acquire();   // [SH | geo] written
            {
                for (uint32_t k = 0; k < 2; k++) mma_f16_if(lead, tmem, desc_kmajor(aF, 2 + k), desc_kmajor(aWc0, k), id64, k > 0);
                for (uint32_t k = 0; k < 2; k++) mma_f16_if(lead, tmem, desc_kmajor(aT, k), desc_kmajor(aWc0, 2 + k), id64, true);
                mma_commit_if(lead, is.mbar);
            }
            iss_acquire();   // C1 written
            { for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, tmem, desc_kmajor(aT, k), desc_kmajor(aWc1, k), id64, k > 0); mma_commit_if(lead, is.mbar); }
            iss_acquire();   // C2 written
            { for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, tmem, desc_kmajor(aT, k), desc_kmajor(aWc2, k), id16, k > 0); mma_commit_if(lead, is.mbar); }
            tile_end_sync();
        }
    } else {
        const bool want_dirs = (hf == 0) && !a.sigma_only;
        FeatPre nf;
        float ndx = 0.f, ndy = 0.f, ndz = 0.f;
        {
            const uint32_t row0 = blockIdx.x * kRows + r;
            nf.load(a.feats, row0, hf, blockIdx.x < a.n_tiles && row0 < a.M);
            if (want_dirs && blockIdx.x < a.n_tiles && row0 < a.M) { ndx = a.dirs[(size_t)row0 * 3]; ndy = a.dirs[(size_t)row0 * 3 + 1]; ndz = a.dirs[(size_t)row0 * 3 + 2]; }
        }
        for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            const uint32_t row = tile * kRows + r;
            const bool in_range = row < a.M;
            nf.store(tF, r, hf);
            const float dx = ndx, dy = ndy, dz = ndz;
            {   // next tile's inputs
                const uint32_t nt = tile + gridDim.x, nrow = nt * kRows + r;
                const bool nin = nt < a.n_tiles && nrow < a.M;
                nf.load(a.feats, nrow, hf, nin);
                ndx = ndy = ndz = 0.f;
                if (want_dirs && nin) { ndx = a.dirs[(size_t)nrow * 3]; ndy = a.dirs[(size_t)nrow * 3 + 1]; ndz = a.dirs[(size_t)nrow * 3 + 2]; }
            }
            epi_publish();
            is.wait();       // sigma layer 0 done
            relu_to_tile(t_row, hf, tT, r);
            epi_publish();
            is.wait();       // sigma layer 1 done (16 outputs)
            if (hf == 0) {
                float h2[16];
                tmem_ld16(t_row, h2);
                if (in_range) {
                    a.sigma[row] = a.density_scale * __expf(h2[0]);
                    if (a.geo) for (int i = 0; i < 15; i++) a.geo[(size_t)row * 15 + i] = h2[1 + i];
                }
                if (!a.sigma_only) {
                    float g[32];
                    sh4(dx, dy, dz, g);
#pragma unroll
                    for (int i = 0; i < 15; i++) g[16 + i] = h2[1 + i];
                    g[31] = 0.0f;
                    store_half_row(tT, r, 0, g);   // colour input, second half: [SH16 | geo15 | 0] -> scratch cols 0..31
                }
            }
            if (a.sigma_only) { tile_end_sync(); continue; }
            epi_publish();
            is.wait();       // colour layer 0 done
            relu_to_tile(t_row, hf, tT, r);
            epi_publish();
            is.wait();       // colour layer 1 done
            relu_to_tile(t_row, hf, tT, r);
            epi_publish();
            is.wait();       // colour layer 2 done (3 outputs)
            if (hf == 0) {
                float o[16];
                tmem_ld16(t_row, o);
                if (in_range) {
#pragma unroll
                    for (int c = 0; c < 3; c++) a.rgb[(size_t)row * 3 + c] = 1.0f / (1.0f + __expf(-o[c]));
                }
            }
            tile_end_sync();
        }
    }
    if (is_issuer) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// MLP forward, activations in TMEM (the shipped forward).
//   The chain feats -> H1 -> h2 -> [colour feats | SH | geo] -> C1 -> C2 -> rgb keeps its A operand in TENSOR MEMORY: the
//   thread that owns a sample row packs its fp16 activations and writes them with tcgen05.st; the next layer's MMA takes A
//   from TMEM (TS form, tc05.cuh) and only the weights from shared memory.  Against the SS-form kernel above (ncu: l1tex 67 %,
//   tensor pipe 13 %: every M128 N64 K16 MMA pulled 4 KB of A + 2 KB of B through the 128 B/clk shared-memory port, 48 cycles
//   for 32 cycles of math, plus the STS traffic of the activation tiles and a fence.proxy.async per round) there are no
//   activation tiles at all: shared memory holds 28 KB of weights, nothing else.
//   TMEM columns of a CTA (128): [0,64) fp32 accumulator | [64,96) H: hidden activations, 64 halfs (the sigma-grid features
//   FS = 32 halfs alias its first 16 columns) | [96,128) CIN: colour-layer input = [colour-grid feats 32 | SH 16 | geo 15 | 0].
//   Thread layout as above: thread (r, hf) owns row r and the column half hf of every 64-wide accumulator; a hidden
//   activation half (32 values) packs into 16 TMEM columns.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTsAcc = 0, kTsH = 64, kTsCin = 96, kTsCols = 128;

__device__ __forceinline__ void relu_to_tmem(uint32_t t_acc_half, uint32_t t_dst) {
    float v[32];
    tmem_ld32(t_acc_half, v);
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; i++) pk[i] = pack_half2(fmaxf(v[2 * i], 0.0f), fmaxf(v[2 * i + 1], 0.0f));
    tmem_st16(t_dst, pk);
}
__device__ __forceinline__ void ts_publish() {   // epilogue warps: my TMEM stores / accumulator reads of this round are done
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
}

__global__ void __launch_bounds__(kMlpThreads)
k_ngp_mlp_fwd_ts(const FwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *tWs0 = smem, *tWc0 = tWs0 + kWTile, *tWc1 = tWc0 + kWTile, *tWs1 = tWc1 + kWTile, *tWc2 = tWs1 + kOTile;
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync(), r = tid & 127, hf = (tid >> 7) & 1u;
    const bool is_issuer = (warp == kIssuerWarp);
    if (is_issuer) tmem_alloc(smem_u32(&s_tmem), kTsCols);
    if (tid == 0) mbar_init(smem_u32(&s_mbar), 1);
    load_weights(tWs0, tWs1, tWc0, tWc1, tWc2, a.w);
    sync_tiles();
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    Issue is{smem_u32(&s_mbar), 0};
    const uint32_t id64 = make_idesc(128, 64, false, false), id16 = make_idesc(128, 16, false, false);

    if (is_issuer) {
        const bool lead = elect_one();
        const uint32_t aWs0 = smem_u32(tWs0), aWs1 = smem_u32(tWs1), aWc0 = smem_u32(tWc0), aWc1 = smem_u32(tWc1), aWc2 = smem_u32(tWc2);
        const uint32_t acc = tmem + kTsAcc, tH = tmem + kTsH, tC = tmem + kTsCin;
        for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            iss_acquire();   // features stored: sigma layer 0, K = 32
            { for (uint32_t k = 0; k < 2; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWs0, k), id64, k > 0); mma_commit_if(lead, is.mbar); }
            iss_acquire();   // H1 stored
            { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWs1, k), id16, k > 0); mma_commit_if(lead, is.mbar); }
            if (a.sigma_only) { tile_end_sync(); continue; }
            iss_acquire();   // [SH | geo] stored next to the colour features
            { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tC + 8 * k, desc_kmajor(aWc0, k), id64, k > 0); mma_commit_if(lead, is.mbar); }
            iss_acquire();   // C1 stored
            { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWc1, k), id64, k > 0); mma_commit_if(lead, is.mbar); }
            iss_acquire();   // C2 stored
            { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWc2, k), id16, k > 0); mma_commit_if(lead, is.mbar); }
            tile_end_sync();
        }
    } else {
        const bool want_dirs = (hf == 0) && !a.sigma_only;
        const uint32_t t_acc = t_row + kTsAcc + hf * 32, t_h = t_row + kTsH + hf * 16;
        // thread (r, hf) stages half of the row's features: hf 0 the sigma-grid half -> FS, hf 1 the colour-grid half -> CIN[0,16)
        const uint32_t t_feat = t_row + (hf == 0 ? kTsH : kTsCin);
        FeatPre nf;
        float ndx = 0.f, ndy = 0.f, ndz = 0.f;
        {
            const uint32_t row0 = blockIdx.x * kRows + r;
            nf.load(a.feats, row0, hf, blockIdx.x < a.n_tiles && row0 < a.M);
            if (want_dirs && blockIdx.x < a.n_tiles && row0 < a.M) { ndx = a.dirs[(size_t)row0 * 3]; ndy = a.dirs[(size_t)row0 * 3 + 1]; ndz = a.dirs[(size_t)row0 * 3 + 2]; }
        }
        for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            const uint32_t row = tile * kRows + r;
            const bool in_range = row < a.M;
            {
                uint32_t pk[16];
#pragma unroll
                for (uint32_t q = 0; q < 4; q++) { pk[4 * q] = nf.f[q].x; pk[4 * q + 1] = nf.f[q].y; pk[4 * q + 2] = nf.f[q].z; pk[4 * q + 3] = nf.f[q].w; }
                if (hf == 0 || !a.sigma_only) tmem_st16(t_feat, pk);
            }
            const float dx = ndx, dy = ndy, dz = ndz;
            {   // next tile's inputs
                const uint32_t nt = tile + gridDim.x, nrow = nt * kRows + r;
                const bool nin = nt < a.n_tiles && nrow < a.M;
                nf.load(a.feats, nrow, hf, nin);
                ndx = ndy = ndz = 0.f;
                if (want_dirs && nin) { ndx = a.dirs[(size_t)nrow * 3]; ndy = a.dirs[(size_t)nrow * 3 + 1]; ndz = a.dirs[(size_t)nrow * 3 + 2]; }
            }
            ts_publish();
            is.wait();       // sigma layer 0 done
            relu_to_tmem(t_acc, t_h);
            ts_publish();
            is.wait();       // sigma layer 1 done (16 outputs)
            if (hf == 0) {
                float h2[16];
                tmem_ld16(t_row + kTsAcc, h2);
                if (in_range) {
                    a.sigma[row] = a.density_scale * __expf(h2[0]);
                    if (a.geo) for (int i = 0; i < 15; i++) a.geo[(size_t)row * 15 + i] = h2[1 + i];
                }
                if (!a.sigma_only) {
                    float g[32];
                    sh4(dx, dy, dz, g);
#pragma unroll
                    for (int i = 0; i < 15; i++) g[16 + i] = h2[1 + i];
                    g[31] = 0.0f;
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) pk[i] = pack_half2(g[2 * i], g[2 * i + 1]);
                    tmem_st16(t_row + kTsCin + 16, pk);   // colour input, second half: [SH16 | geo15 | 0]
                }
            }
            if (a.sigma_only) { tile_end_sync(); continue; }
            ts_publish();
            is.wait();       // colour layer 0 done
            relu_to_tmem(t_acc, t_h);
            ts_publish();
            is.wait();       // colour layer 1 done
            relu_to_tmem(t_acc, t_h);
            ts_publish();
            is.wait();       // colour layer 2 done (3 outputs)
            if (hf == 0) {
                float o[16];
                tmem_ld16(t_row + kTsAcc, o);
                if (in_range) {
#pragma unroll
                    for (int c = 0; c < 3; c++) a.rgb[(size_t)row * 3 + c] = 1.0f / (1.0f + __expf(-o[c]));
                }
            }
            tile_end_sync();
        }
    }
    if (is_issuer) tmem_dealloc(tmem, kTsCols);
}

// ------------------------------------------------------------------------------------------------
// Teacher + student field forward in ONE kernel: hash-grid gather and both MLPs on the same SM at the same time.
//
//   The gather is bound by the SM's load pipe (one L1 wavefront per divergent 16-byte corner), the MLP by the tensor-core /
//   TMEM round trips of its five dependent layers; run as two kernels they take 1.65 + 0.61 ms and each leaves the other's
//   resource idle.  Here a persistent CTA (one per SM) owns two SLOTS of tensor memory (256 columns each: a teacher chain and
//   a student chain laid out like k_ngp_mlp_fwd_ts).  Each slot has 8 generalist warps: they gather the 128 samples of a tile
//   for both models from the paired table (thread (r, hf): sample r, levels 8 hf .. 8 hf + 7) and write the feature halves
//   STRAIGHT INTO TENSOR MEMORY as the first layer's A operand (tcgen05.st; no [M,64] feature rows through HBM for the teacher,
//   no shared-memory tile), then serve the epilogues of the slot's two MLP chains.  The two slots run half a period apart, so
//   while one slot's warps wait on tensor-core round trips the other slot's warps keep the load pipe busy.  One issuer warp
//   per slot issues that slot's MMAs; every hand-off is an mbarrier (bounded waits: a protocol error traps instead of hanging).
//   The student's feature rows are also written to global memory: the backward recomputes the MLP from them.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kPairSlotThreads = 256, kPairSlots = 2;
constexpr uint32_t kPairThreads = kPairSlots * kPairSlotThreads + kPairSlots * 32;   // + one issuer warp per slot

struct PairArgs {
    const float *xyz, *xyz_teacher, *dirs, *dirs_teacher;
    const uint8_t *mask;
    const uint4 *table8;
    const int *offsets;
    MlpWeights wt, ws;
    float *sigma_t, *rgb_t, *sigma_s, *rgb_s;
    __half *feats_s;
    uint32_t M, n_tiles, L, H;
    float bound, S, dscale_t, dscale_s;
};

__global__ void __launch_bounds__(kPairThreads, 1)
k_ngp_pair_fwd(const PairArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Geo g;
    __shared__ __align__(8) uint64_t s_full[kPairSlots][2], s_done[kPairSlots][2];
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t kWBytes = 3 * kWTile + 2 * kOTile;
    uint8_t *wT = smem, *wS = smem + kWBytes;
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();
    const bool is_issuer = warp >= kPairSlots * 8;
    const uint32_t slot = is_issuer ? warp - kPairSlots * 8 : warp >> 3;
    geo_init(g, a.offsets, a.L, a.S, a.H);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
    if (tid == 32) {
        for (uint32_t s = 0; s < kPairSlots; s++)
            for (uint32_t c = 0; c < 2; c++) { mbar_init(smem_u32(&s_full[s][c]), kPairSlotThreads); mbar_init(smem_u32(&s_done[s][c]), 1); }
    }
    load_weights(wT, wT + 3 * kWTile, wT + kWTile, wT + 2 * kWTile, wT + 3 * kWTile + kOTile, a.wt);
    load_weights(wS, wS + 3 * kWTile, wS + kWTile, wS + 2 * kWTile, wS + 3 * kWTile + kOTile, a.ws);
    sync_tiles();
    const uint32_t tmem = s_tmem;
    const uint32_t id64 = make_idesc(128, 64, false, false), id16 = make_idesc(128, 16, false, false);
    const uint32_t first = blockIdx.x * kPairSlots + slot, stride = gridDim.x * kPairSlots;

    if (is_issuer) {
        const bool lead = elect_one();
        uint32_t par = 0;
        for (uint32_t tile = first; tile < a.n_tiles; tile += stride) {
            for (uint32_t round = 0; round < 5; round++) {
#pragma unroll
                for (uint32_t c = 0; c < 2; c++) {
                    const uint32_t base = tmem + slot * 256 + c * 128, acc = base + kTsAcc, tH = base + kTsH, tC = base + kTsCin;
                    const uint32_t w0 = smem_u32(c == 0 ? wT : wS);   // tile order inside a weight block: Ws0 | Wc0 | Wc1 | Ws1 | Wc2
                    const uint32_t aWs0 = w0, aWc0 = w0 + kWTile, aWc1 = w0 + 2 * kWTile, aWs1 = w0 + 3 * kWTile, aWc2 = w0 + 3 * kWTile + kOTile;
                    mbar_wait(smem_u32(&s_full[slot][c]), par);
                    fence_after_sync();
                    if (round == 0) { for (uint32_t k = 0; k < 2; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWs0, k), id64, k > 0); }
                    else if (round == 1) { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWs1, k), id16, k > 0); }
                    else if (round == 2) { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tC + 8 * k, desc_kmajor(aWc0, k), id64, k > 0); }
                    else if (round == 3) { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWc1, k), id64, k > 0); }
                    else { for (uint32_t k = 0; k < 4; k++) mma_f16_ts_if(lead, acc, tH + 8 * k, desc_kmajor(aWc2, k), id16, k > 0); }
                    mma_commit_if(lead, smem_u32(&s_done[slot][c]));
                }
                par ^= 1;
            }
        }
    } else {
        const uint32_t st = tid & (kPairSlotThreads - 1), r = st & 127, hf = st >> 7;
        const uint32_t t_row = tmem + (((warp & 3u) * 32u) << 16) + slot * 256;
        uint32_t par = 0;    // parity of the done barriers (flips once per round)
        auto publish = [&](uint32_t c) {
            tmem_st_wait();
            fence_before_sync();
            mbar_arrive(smem_u32(&s_full[slot][c]));
        };
        auto wait_done = [&](uint32_t c) {
            mbar_wait(smem_u32(&s_done[slot][c]), par);
            fence_after_sync();
        };
        for (uint32_t tile = first; tile < a.n_tiles; tile += stride) {
            const uint32_t row = tile * kRows + r;
            const bool in_range = row < a.M;
            // ---------------- gather: levels 8 hf .. 8 hf + 7 of sample `row`, teacher and student ----------------
            float ux = 0, uy = 0, uz = 0, tx = 0, ty = 0, tz = 0;
            bool ok = false, moved = false, tok = false;
            if (in_range) {
                ok = load_unit(a.xyz, row, a.bound, ux, uy, uz);
                moved = a.mask && a.mask[row];
                tok = ok;
                if (moved) tok = load_unit(a.xyz_teacher, row, a.bound, tx, ty, tz);
            }
            uint32_t pTs[8], pTc[8], pSs[8], pSc[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; j++) {
                const uint32_t l = hf * 8 + j;
                float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                if (ok && l < a.L) {
                    Cell c;
                    locate(g, l, ux, uy, uz, c, nullptr);
                    uint4 v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] = __ldg(a.table8 + c.idx[k]);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        acc4(make_uint2(v[k].z, v[k].w), c.w[k], s0, s1, s2, s3);
                        if (!moved) acc4(make_uint2(v[k].x, v[k].y), c.w[k], t0, t1, t2, t3);
                    }
                }
                if (moved && tok && l < a.L) {
                    Cell c;
                    locate(g, l, tx, ty, tz, c, nullptr);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint4 v = __ldg(a.table8 + c.idx[k]);
                        acc4(make_uint2(v.x, v.y), c.w[k], t0, t1, t2, t3);
                    }
                }
                pTs[j] = pack2(t0, t1); pTc[j] = pack2(t2, t3); pSs[j] = pack2(s0, s1); pSc[j] = pack2(s2, s3);
            }
            // feature halves -> tensor memory (A operands of sigma layer 0 and colour layer 0), student rows -> global
            tmem_st8(t_row + 0 * 128 + kTsH + 8 * hf, pTs);
            tmem_st8(t_row + 0 * 128 + kTsCin + 8 * hf, pTc);
            tmem_st8(t_row + 1 * 128 + kTsH + 8 * hf, pSs);
            tmem_st8(t_row + 1 * 128 + kTsCin + 8 * hf, pSc);
            if (in_range) {
                uint4 *dst = reinterpret_cast<uint4 *>(a.feats_s + (size_t)row * 64);
                dst[2 * hf] = make_uint4(pSs[0], pSs[1], pSs[2], pSs[3]); dst[2 * hf + 1] = make_uint4(pSs[4], pSs[5], pSs[6], pSs[7]);
                dst[4 + 2 * hf] = make_uint4(pSc[0], pSc[1], pSc[2], pSc[3]); dst[4 + 2 * hf + 1] = make_uint4(pSc[4], pSc[5], pSc[6], pSc[7]);
            }
            float dT[3] = {0.f, 0.f, 0.f}, dS[3] = {0.f, 0.f, 0.f};
            if (hf == 0 && in_range) {
#pragma unroll
                for (int i = 0; i < 3; i++) { dS[i] = a.dirs[(size_t)row * 3 + i]; dT[i] = a.dirs_teacher[(size_t)row * 3 + i]; }
            }
            publish(0); publish(1);
            // ---------------- the two MLP chains, epilogues interleaved ----------------
            // round 0: sigma layer 0 -> H1
#pragma unroll
            for (uint32_t c = 0; c < 2; c++) { wait_done(c); relu_to_tmem(t_row + c * 128 + kTsAcc + hf * 32, t_row + c * 128 + kTsH + hf * 16); publish(c); }
            par ^= 1;
            // round 1: sigma layer 1 -> sigma, [SH | geo]
#pragma unroll
            for (uint32_t c = 0; c < 2; c++) {
                wait_done(c);
                if (hf == 0) {
                    float h2[16], gg[32];
                    tmem_ld16(t_row + c * 128 + kTsAcc, h2);
                    if (in_range) (c == 0 ? a.sigma_t : a.sigma_s)[row] = (c == 0 ? a.dscale_t : a.dscale_s) * __expf(h2[0]);
                    const float *dd = c == 0 ? dT : dS;
                    sh4(dd[0], dd[1], dd[2], gg);
#pragma unroll
                    for (int i = 0; i < 15; i++) gg[16 + i] = h2[1 + i];
                    gg[31] = 0.0f;
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) pk[i] = pack_half2(gg[2 * i], gg[2 * i + 1]);
                    tmem_st16(t_row + c * 128 + kTsCin + 16, pk);
                }
                publish(c);
            }
            par ^= 1;
            // rounds 2, 3: colour layers 0, 1
            for (uint32_t rr = 0; rr < 2; rr++) {
#pragma unroll
                for (uint32_t c = 0; c < 2; c++) { wait_done(c); relu_to_tmem(t_row + c * 128 + kTsAcc + hf * 32, t_row + c * 128 + kTsH + hf * 16); publish(c); }
                par ^= 1;
            }
            // round 4: colour layer 2 -> rgb
#pragma unroll
            for (uint32_t c = 0; c < 2; c++) {
                wait_done(c);
                if (hf == 0) {
                    float o[16];
                    tmem_ld16(t_row + c * 128 + kTsAcc, o);
                    if (in_range) {
                        float *dst = (c == 0 ? a.rgb_t : a.rgb_s) + (size_t)row * 3;
#pragma unroll
                        for (int i = 0; i < 3; i++) dst[i] = 1.0f / (1.0f + __expf(-o[i]));
                    }
                }
            }
            par ^= 1;
        }
    }
    tile_end_sync();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// MLP backward (persistent, 1 CTA / SM)
// ------------------------------------------------------------------------------------------------
struct BwdArgs {
    const __half *feats;
    const float *dirs;
    MlpWeights w;
    const float *g_sigma, *g_rgb;   // dL/dsigma [M] (w.r.t. density_scale * exp(h0)), dL/drgb [M,3]
    __half *dfeats;                 // [M,64]
    float *gw_s0, *gw_s1, *gw_c0, *gw_c1, *gw_c2;  // fp32, nn.Linear shapes, accumulated into (atomics)
    uint32_t M, n_tiles;
    float density_scale, out_scale;  // out_scale multiplies dfeats (keeps fp16 gradients in range)
    int train_mlp;
    uint32_t *nonfinite;             // non-NULL: deterministic mode, gw_* point at 64-bit fixed-point values (see fixed_add)
};

// ------------------------------------------------------------------------------------------------
// MLP backward: two dependency chains in flight, one epilogue warp set per chain.
//   * The backward of tile t (5 tensor-core rounds) and the forward recompute of tile t+1 (5 rounds) are independent
//     chains.  Threads 0-255 (8 warps) run the forward epilogues, threads 256-511 the backward epilogues, warp 16 issues
//     every tcgen05.mma.  ncu on the one-chain predecessor showed 41 % of the epilogue warps' samples on the
//     MMA-completion mbarrier and a 20 % busy tensor pipe: the kernel is bound by the publish -> issue -> MMA -> commit ->
//     wake round trip of each round, so the second chain fills that time.
//   * Two sets of six [128 x 64] fp16 tiles {F, H1, G, C1, C2, X} alternate between consecutive tiles (192 KB + 28 KB of
//     weights).  The single gradient tile X of a set is reused by every backward round: a round's commit follows its
//     weight-gradient MMAs, so the wait that guards the accumulator read also guarantees nothing still reads X.
//   * TMEM columns: [0,64) backward accumulator | [64,128) dWc2 | [128,192) dWc1 | [192,256) dC1^T.F | [256,320) dC1^T.G |
//     [320,384) dWs1 | [384,448) dH1^T.F | [448,512) forward accumulator.  Weight gradients accumulate in TMEM over all
//     tiles of the CTA (every block a full M = 64, N = 64 MMA) and are flushed once with fp32 atomics.
//   * Hand-off: an epilogue set publishes with bar.arrive on its named barrier (1 = forward, 2 = backward) and goes
//     straight to its mbarrier; the issuer bar.syncs on the two barriers alternately and always commits (an empty commit
//     in the first / last iteration keeps the sets in lock step).  The sets meet once per iteration (named barrier 3)
//     before the tile sets swap roles; the sigma logit travels forward set -> backward set through shared memory.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kSetTiles = 6;
enum : uint32_t { kTF = 0, kTH1 = 1, kTG = 2, kTC1 = 3, kTC2 = 4, kTX = 5 };

#ifdef S3D_TRACE
__device__ long long g_trace[1 << 15];
#define S3D_TR(role, ev, it) do { if (blockIdx.x == 0 && (it) >= 2 && (it) < 8) { g_trace[(((it) - 2) * 3 + (role)) * 64 + (ev)] = clock64(); } } while (0)
#else
#define S3D_TR(role, ev, it) do { } while (0)
#endif
constexpr uint32_t kBwdThreads = 544;
constexpr uint32_t kBwdIssuer = 16;

__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync_n(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void set_publish(uint32_t id) {   // epilogue set: tile writes / accumulator reads of this round are done
    fence_async_smem();
    fence_before_sync();
    bar_arrive(id, 288);
}
__device__ __forceinline__ void iss_acquire_n(uint32_t id) {
    bar_sync_n(id, 288);
    fence_after_sync();
}

template <bool FIXED>   // FIXED: the weight gradients are flushed as 64-bit fixed point (deterministic mode)
__global__ void __launch_bounds__(kBwdThreads, 1)
k_ngp_mlp_bwd(const BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar[3];
    __shared__ uint32_t s_tmem;
    __shared__ float s_h0[2][kRows];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *tWs0 = smem + 2 * kSetTiles * kTileBytes, *tWc0 = tWs0 + kWTile, *tWc1 = tWc0 + kWTile, *tWs1 = tWc1 + kWTile, *tWc2 = tWs1 + kOTile;
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync(), r = tid & 127, hf = (tid >> 7) & 1u;
    const bool is_issuer = (warp == kBwdIssuer), is_fwd = warp < 8;
    if (is_issuer) tmem_alloc(smem_u32(&s_tmem), 512);
    if (tid == 0) { mbar_init(smem_u32(&s_mbar[0]), 1); mbar_init(smem_u32(&s_mbar[1]), 1); mbar_init(smem_u32(&s_mbar[2]), 1); }
    load_weights(tWs0, tWs1, tWc0, tWc1, tWc2, a.w);
    sync_tiles();
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t n_my = blockIdx.x < a.n_tiles ? (a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t mbarF = smem_u32(&s_mbar[0]), mbarB = smem_u32(&s_mbar[1]), mbarW = smem_u32(&s_mbar[2]);
    constexpr uint32_t kAccF = 448, kAccB = 0;

    if (is_issuer) {
        const bool lead = elect_one();
        const uint32_t base = smem_u32(smem);
        const uint32_t aWs0 = smem_u32(tWs0), aWs1 = smem_u32(tWs1), aWc0 = smem_u32(tWc0), aWc1 = smem_u32(tWc1), aWc2 = smem_u32(tWc2);
        const uint32_t id64 = make_idesc(128, 64, false, false), id16 = make_idesc(128, 16, false, false);
        const uint32_t id64t = make_idesc(128, 64, false, true);        // B read MN-major (= W^T)
        const uint32_t idw64 = make_idesc(64, 64, true, true);
        const uint32_t accF = tmem + kAccF, accB = tmem + kAccB;
        const uint32_t accC2 = tmem + 64, accC1 = tmem + 128, accC0f = tmem + 192, accC0g = tmem + 256, accS1 = tmem + 320, accS0 = tmem + 384;
        const bool tw = a.train_mlp != 0;
        for (uint32_t it = 0; it <= n_my; it++) {
            const bool doF = it < n_my, doB = it > 0, firstB = (it == 1);
            const uint32_t sf = base + (it & 1u) * kSetTiles * kTileBytes, sb = base + ((it & 1u) ^ 1u) * kSetTiles * kTileBytes;
            const uint32_t fF = sf + kTF * kTileBytes, fH1 = sf + kTH1 * kTileBytes, fG = sf + kTG * kTileBytes, fC1 = sf + kTC1 * kTileBytes, fC2 = sf + kTC2 * kTileBytes;
            const uint32_t bF = sb + kTF * kTileBytes, bH1 = sb + kTH1 * kTileBytes, bG = sb + kTG * kTileBytes, bC1 = sb + kTC1 * kTileBytes, bC2 = sb + kTC2 * kTileBytes, bX = sb + kTX * kTileBytes;
            iss_acquire_n(1); S3D_TR(0, 0, it);   // features stored
            { if (doF) for (uint32_t k = 0; k < 2; k++) mma_f16_if(lead, accF, desc_kmajor(fF, k), desc_kmajor(aWs0, k), id64, k > 0); mma_commit_if(lead, mbarF); S3D_TR(0, 1, it); }
            iss_acquire_n(2); S3D_TR(0, 2, it);   // dO in X (published before the sets met):  dC2 = dO . Wc2 ;  dWc2 += dO^T . C2
            {
              if (doB) {
                mma_f16_if(lead, accB, desc_kmajor(bX, 0), desc_mnmajor(aWc2, 0, kOTile), id64t, false);
              }
              mma_commit_if(lead, mbarB);
              if (doB) {
                if (tw) for (uint32_t k = 0; k < 8; k++) mma_f16_if(lead, accC2, desc_mnmajor(bX, k, kTileBytes), desc_mnmajor(bC2, k, kTileBytes), idw64, !(firstB && k == 0));
              }
              mma_commit_if(lead, mbarW); S3D_TR(0, 3, it);
            }
            iss_acquire_n(1); S3D_TR(0, 4, it);   // H1 written
            { if (doF) for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, accF, desc_kmajor(fH1, k), desc_kmajor(aWs1, k), id16, k > 0); mma_commit_if(lead, mbarF); S3D_TR(0, 5, it); }
            iss_acquire_n(2); S3D_TR(0, 6, it);   // dC2 in X:  dC1 = dC2 . Wc1 ;  dWc1 += dC2^T . C1
            {
              if (doB) {
                for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, accB, desc_kmajor(bX, k), desc_mnmajor(aWc1, k, kWTile), id64t, k > 0);
              }
              mma_commit_if(lead, mbarB);
              if (doB) {
                if (tw) for (uint32_t k = 0; k < 8; k++) mma_f16_if(lead, accC1, desc_mnmajor(bX, k, kTileBytes), desc_mnmajor(bC1, k, kTileBytes), idw64, !(firstB && k == 0));
              }
              mma_commit_if(lead, mbarW); S3D_TR(0, 7, it);
            }
            iss_acquire_n(1); S3D_TR(0, 8, it);   // [SH | geo] written
            {
                if (doF) {
                    for (uint32_t k = 0; k < 2; k++) mma_f16_if(lead, accF, desc_kmajor(fF, 2 + k), desc_kmajor(aWc0, k), id64, k > 0);
                    for (uint32_t k = 0; k < 2; k++) mma_f16_if(lead, accF, desc_kmajor(fG, k), desc_kmajor(aWc0, 2 + k), id64, true);
                }
                mma_commit_if(lead, mbarF); S3D_TR(0, 9, it);
            }
            iss_acquire_n(2); S3D_TR(0, 10, it);   // dC1 in X:  d[colour feats | SH | geo] = dC1 . Wc0p ;  dWc0 += dC1^T . [F | G]
            {
              if (doB) {
                for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, accB, desc_kmajor(bX, k), desc_mnmajor(aWc0, k, kWTile), id64t, k > 0);
              }
              mma_commit_if(lead, mbarB);
              if (doB) {
                if (tw) {
                    for (uint32_t k = 0; k < 8; k++) mma_f16_if(lead, accC0f, desc_mnmajor(bX, k, kTileBytes), desc_mnmajor(bF, k, kTileBytes), idw64, !(firstB && k == 0));
                    for (uint32_t k = 0; k < 8; k++) mma_f16_if(lead, accC0g, desc_mnmajor(bX, k, kTileBytes), desc_mnmajor(bG, k, kTileBytes), idw64, !(firstB && k == 0));
                }
              }
              mma_commit_if(lead, mbarW); S3D_TR(0, 11, it);
            }
            iss_acquire_n(1); S3D_TR(0, 12, it);   // C1 written
            { if (doF) for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, accF, desc_kmajor(fC1, k), desc_kmajor(aWc1, k), id64, k > 0); mma_commit_if(lead, mbarF); S3D_TR(0, 13, it); }
            iss_acquire_n(2); S3D_TR(0, 14, it);   // dh2 in X:  dH1 = dh2 . Ws1 ;  dWs1 += dh2^T . H1
            {
              if (doB) {
                mma_f16_if(lead, accB, desc_kmajor(bX, 0), desc_mnmajor(aWs1, 0, kOTile), id64t, false);
              }
              mma_commit_if(lead, mbarB);
              if (doB) {
                if (tw) for (uint32_t k = 0; k < 8; k++) mma_f16_if(lead, accS1, desc_mnmajor(bX, k, kTileBytes), desc_mnmajor(bH1, k, kTileBytes), idw64, !(firstB && k == 0));
              }
              mma_commit_if(lead, mbarW); S3D_TR(0, 15, it);
            }
            iss_acquire_n(1); S3D_TR(0, 16, it);   // C2 written
            { if (doF) for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, accF, desc_kmajor(fC2, k), desc_kmajor(aWc2, k), id16, k > 0); mma_commit_if(lead, mbarF); S3D_TR(0, 17, it); }
            iss_acquire_n(2); S3D_TR(0, 18, it);   // dH1 in X:  dF = dH1 . Ws0 ;  dWs0 += dH1^T . F
            {
              if (doB) {
                for (uint32_t k = 0; k < 4; k++) mma_f16_if(lead, accB, desc_kmajor(bX, k), desc_mnmajor(aWs0, k, kWTile), id64t, k > 0);
              }
              mma_commit_if(lead, mbarB);
              if (doB) {
                if (tw) for (uint32_t k = 0; k < 8; k++) mma_f16_if(lead, accS0, desc_mnmajor(bX, k, kTileBytes), desc_mnmajor(bF, k, kTileBytes), idw64, !(firstB && k == 0));
              }
              mma_commit_if(lead, mbarW); S3D_TR(0, 19, it);
            }
        }
    } else if (is_fwd) {
        // ================= forward recompute of tile it (set it & 1) =================
        Issue isF{mbarF, 0};
        const uint32_t accF = t_row + kAccF;
        FeatPre nf;
        float nin6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // warpgroup 0: dir xyz, dL/drgb of the next tile
        auto prefetch = [&](uint32_t t) {
            const uint32_t prow = t * kRows + r;
            const bool pin = t < a.n_tiles && prow < a.M;
            nf.load(a.feats, prow, hf, pin);
            if (hf == 0) {
#pragma unroll
                for (int j = 0; j < 3; j++) { nin6[j] = pin ? a.dirs[(size_t)prow * 3 + j] : 0.f; nin6[3 + j] = pin ? a.g_rgb[(size_t)prow * 3 + j] : 0.f; }
            }
        };
        prefetch(blockIdx.x);
        for (uint32_t it = 0; it <= n_my; it++) {
            const bool doF = it < n_my;
            uint8_t *sF = smem + (it & 1u) * kSetTiles * kTileBytes;
            uint8_t *fF = sF + kTF * kTileBytes, *fH1 = sF + kTH1 * kTileBytes, *fG = sF + kTG * kTileBytes, *fC1 = sF + kTC1 * kTileBytes, *fC2 = sF + kTC2 * kTileBytes, *fX = sF + kTX * kTileBytes;
            const uint32_t tile_f = blockIdx.x + it * gridDim.x, row_f = tile_f * kRows + r;
            const bool in_f = doF && row_f < a.M;
            float in6[6];
#pragma unroll
            for (int j = 0; j < 6; j++) in6[j] = nin6[j];
            if (doF) {
                nf.store(fF, r, hf);
                prefetch(tile_f + gridDim.x);
            }
            if (tid == 0) S3D_TR(1, 0, it); set_publish(1);
            isF.wait(); if (tid == 0) S3D_TR(1, 1, it);
            if (doF) relu_to_tile(accF, hf, fH1, r);
            if (tid == 0) S3D_TR(1, 2, it); set_publish(1);
            isF.wait(); if (tid == 0) S3D_TR(1, 3, it);
            if (doF) {
                if (hf == 0) {
                    float h2[16], g[32];
                    tmem_ld16(accF, h2);
                    s_h0[it & 1u][r] = h2[0];
                    sh4(in6[0], in6[1], in6[2], g);
#pragma unroll
                    for (int i = 0; i < 15; i++) g[16 + i] = h2[1 + i];
                    g[31] = 0.0f;
                    store_half_row(fG, r, 0, g);
                } else {
                    zero_half_row(fG, r, 1);
                }
            }
            if (tid == 0) S3D_TR(1, 4, it); set_publish(1);
            isF.wait(); if (tid == 0) S3D_TR(1, 5, it);
            if (doF) relu_to_tile(accF, hf, fC1, r);
            if (tid == 0) S3D_TR(1, 6, it); set_publish(1);
            isF.wait(); if (tid == 0) S3D_TR(1, 7, it);
            if (doF) relu_to_tile(accF, hf, fC2, r);
            if (tid == 0) S3D_TR(1, 8, it); set_publish(1);
            isF.wait(); if (tid == 0) S3D_TR(1, 9, it);
            if (doF) {
                if (hf == 0) {   // output gradient dO (3 meaningful columns, zero padded)
                    float o[16], d[32];
                    tmem_ld16(accF, o);
#pragma unroll
                    for (int i = 0; i < 32; i++) d[i] = 0.0f;
                    if (in_f) {
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const float sg = 1.0f / (1.0f + __expf(-o[c]));
                            d[c] = in6[3 + c] * sg * (1.0f - sg);
                        }
                    }
                    store_half_row(fX, r, 0, d);
                } else {
                    zero_half_row(fX, r, 1);
                }
            }
            fence_async_smem();
            fence_before_sync();
            if (tid == 0) S3D_TR(1, 20, it);
            bar_sync_n(3, 512);     // the sets meet: this set's tiles (and dO, h0) are complete, the other set is free
            fence_after_sync();
        }
    } else {
        // ================= backward of tile it - 1 (set (it - 1) & 1) =================
        Issue isB{mbarB, 0}, isW{mbarW, 0};
        const uint32_t accB = t_row + kAccB;
        for (uint32_t it = 0; it <= n_my; it++) {
            const bool doB = it > 0;
            uint8_t *sB = smem + ((it & 1u) ^ 1u) * kSetTiles * kTileBytes;
            uint8_t *bH1 = sB + kTH1 * kTileBytes, *bC1 = sB + kTC1 * kTileBytes, *bC2 = sB + kTC2 * kTileBytes, *bX = sB + kTX * kTileBytes;
            const uint32_t tile_b = blockIdx.x + (it - 1) * gridDim.x, row_b = tile_b * kRows + r;
            const bool in_b = doB && row_b < a.M;
            float gs_b = 0.f, h0_b = 0.f;
            if (in_b && hf == 0) { gs_b = a.g_sigma[row_b]; h0_b = s_h0[(it & 1u) ^ 1u][r]; }
            if (tid == 256) S3D_TR(2, 0, it); set_publish(2);
            isB.wait(); if (tid == 256) S3D_TR(2, 1, it);
            { uint4 pk[4]; if (doB) masked_half_row(accB, hf, bC2, r, pk); isW.wait(); if (doB) store_packed_half_row(bX, r, hf, pk); }      // dC2 (X is free once the weight-gradient MMAs of the round are done)
            if (tid == 256) S3D_TR(2, 2, it); set_publish(2);
            isB.wait(); if (tid == 256) S3D_TR(2, 3, it);
            { uint4 pk[4]; if (doB) masked_half_row(accB, hf, bC1, r, pk); isW.wait(); if (doB) store_packed_half_row(bX, r, hf, pk); }      // dC1
            if (tid == 256) S3D_TR(2, 4, it); set_publish(2);
            isB.wait(); if (tid == 256) S3D_TR(2, 5, it);
            if (doB) {
                if (hf == 1) {   // gradient w.r.t. the colour-grid features (dfeats cols 32..63)
                    float dfc[32];
                    tmem_ld32(accB, dfc);
                    if (in_b) {
                        uint4 *dst = reinterpret_cast<uint4 *>(a.dfeats + (size_t)row_b * 64) + 4;
#pragma unroll
                        for (int i = 0; i < 32; i++) dfc[i] *= a.out_scale;
#pragma unroll
                        for (uint32_t q = 0; q < 4; q += 2) stg_pair(dst + q, pack8(dfc + q * 8), pack8(dfc + q * 8 + 8));
                    }
                    isW.wait();
                    zero_half_row(bX, r, 1);
                } else {
                    float v[32], d[32];
                    tmem_ld32(accB + 32, v);  // cols 32-47 dSH (dropped), 48-62 dgeo
#pragma unroll
                    for (int i = 0; i < 32; i++) d[i] = 0.0f;
                    if (in_b) d[0] = gs_b * a.density_scale * __expf(fminf(fmaxf(h0_b, -15.0f), 15.0f));  // trunc_exp backward
#pragma unroll
                    for (int i = 0; i < 15; i++) d[1 + i] = v[16 + i];
                    isW.wait();
                    store_half_row(bX, r, 0, d);            // dh2
                }
            } else {
                isW.wait();
            }
            if (tid == 256) S3D_TR(2, 6, it); set_publish(2);
            isB.wait(); if (tid == 256) S3D_TR(2, 7, it);
            { uint4 pk[4]; if (doB) masked_half_row(accB, hf, bH1, r, pk); isW.wait(); if (doB) store_packed_half_row(bX, r, hf, pk); }      // dH1
            if (tid == 256) S3D_TR(2, 8, it); set_publish(2);
            isB.wait(); if (tid == 256) S3D_TR(2, 9, it);
            if (doB) {
                if (hf == 0) {
                    float dfs[32];
                    tmem_ld32(accB, dfs);
                    if (in_b) {
                        uint4 *dst = reinterpret_cast<uint4 *>(a.dfeats + (size_t)row_b * 64);
#pragma unroll
                        for (int i = 0; i < 32; i++) dfs[i] *= a.out_scale;
#pragma unroll
                        for (uint32_t q = 0; q < 4; q += 2) stg_pair(dst + q, pack8(dfs + q * 8), pack8(dfs + q * 8 + 8));
                    }
                }
            }
            isW.wait();   // the last weight-gradient MMAs still read this set's tiles
            fence_before_sync();
            if (tid == 256) S3D_TR(2, 20, it);
            bar_sync_n(3, 512);
            fence_after_sync();
            if (tid == 256) S3D_TR(2, 21, it);
        }

        // ---- flush weight gradients: accumulator rows (M = 64) live in lanes 0..15 of every 32-lane sub-partition;
        //      warpgroup hf takes the columns [32*hf, 32*hf+32) of every block ----
        if (a.train_mlp && n_my > 0) {
            const uint32_t lane = tid & 31, rr = (warp & 3u) * 16 + lane;   // output feature
            const bool rowok = lane < 16;
            float v[32];
            uint32_t *const nf = a.nonfinite;
            auto add = [nf](float *base, uint32_t idx, float x) {
                if constexpr (FIXED) fixed_add(reinterpret_cast<long long *>(base) + idx, x, nf);
                else atomicAdd(base + idx, x);
            };
            tmem_ld32(t_row + 64 + hf * 32, v);    // dWc2 [3 x 64]
            if (rowok && rr < 3) for (int i = 0; i < 32; i++) add(a.gw_c2, rr * 64 + hf * 32 + i, v[i]);
            tmem_ld32(t_row + 128 + hf * 32, v);   // dWc1 [64 x 64]
            if (rowok) for (int i = 0; i < 32; i++) add(a.gw_c1, rr * 64 + hf * 32 + i, v[i]);
            tmem_ld32(t_row + 320 + hf * 32, v);   // dWs1 [16 x 64]
            if (rowok && rr < 16) for (int i = 0; i < 32; i++) add(a.gw_s1, rr * 64 + hf * 32 + i, v[i]);
            if (hf == 1) {
                tmem_ld32(t_row + 192 + 32, v);    // dC1^T . F, columns 32..63 = colour-grid inputs = original columns 31..62
                if (rowok) for (int i = 0; i < 32; i++) add(a.gw_c0, rr * 63 + 31 + i, v[i]);
            } else {
                tmem_ld32(t_row + 256, v);         // dC1^T . G, columns 0..30 = [SH | geo] = original columns 0..30
                if (rowok) for (int i = 0; i < 31; i++) add(a.gw_c0, rr * 63 + i, v[i]);
                tmem_ld32(t_row + 384, v);         // dH1^T . F, columns 0..31 = sigma-grid inputs
                if (rowok) for (int i = 0; i < 32; i++) add(a.gw_s0, rr * 32 + i, v[i]);
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (is_issuer) tmem_dealloc(tmem, 512);
}

// interleave two [N,2] fp32 tables into one [N,4] fp16 table {s0,s1,c0,c1}
__global__ void k_interleave_tables(const float2 *__restrict__ ts, const float2 *__restrict__ tc, uint2 *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 s = ts[i], c = tc[i];
    __half2 a = __floats2half2_rn(s.x, s.y), b = __floats2half2_rn(c.x, c.y);
    out[i] = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
}

// Adam over the two hash tables with an interleaved gradient / shadow layout
__global__ void __launch_bounds__(256)
k_adam_tables(float2 *__restrict__ ps, float2 *__restrict__ pc, float4 *__restrict__ g4, float4 *__restrict__ m4, float4 *__restrict__ v4,
              uint8_t *__restrict__ shadow, uint32_t shadow_stride, size_t n, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps, float gscale,
              const float *__restrict__ scaler, float lr) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool skip = false;
    if (scaler) {   // device-side GradScaler state (csrc/train.cu): scale, skip decision, bias corrections of the applied-step count
        skip = scaler[2] != 0.0f;
        gscale = gscale / scaler[0];
        lr_over_bc1 = lr * scaler[4];
        inv_sqrt_bc2 = scaler[5];
    }
    const float4 g = g4[i];
    float4 m = m4[i];
    const bool gz = (g.x == 0.0f) & (g.y == 0.0f) & (g.z == 0.0f) & (g.w == 0.0f);
    const bool mz = (m.x == 0.0f) & (m.y == 0.0f) & (m.z == 0.0f) & (m.w == 0.0f);
    // An entry that has never received a gradient has g = m = v = 0 and dense Adam leaves it exactly unchanged
    // (0 / (0 + eps) = 0): skipping it is bit-identical and saves 104 of its 136 bytes of traffic.  Entries that were
    // touched once keep decaying their moments and moving, like torch.optim.Adam.
    if (gz && mz) return;
    if (!gz) g4[i] = make_float4(0, 0, 0, 0);
    if (skip) return;   // non-finite gradient somewhere: no update, gradient cleared
    float4 v = v4[i];
    float2 s = ps[i], c = pc[i];
    const float *G = &g.x;
    float *Mv = &m.x, *V = &v.x;
    float P[4] = {s.x, s.y, c.x, c.y};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float gr = G[k] * gscale;
        Mv[k] = b1 * Mv[k] + (1.0f - b1) * gr;
        V[k] = b2 * V[k] + (1.0f - b2) * gr * gr;
        P[k] -= lr_over_bc1 * Mv[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
    }
    m4[i] = m; v4[i] = v;
    ps[i] = make_float2(P[0], P[1]); pc[i] = make_float2(P[2], P[3]);
    __half2 a = __floats2half2_rn(P[0], P[1]), b = __floats2half2_rn(P[2], P[3]);
    *reinterpret_cast<uint2 *>(shadow + i * shadow_stride) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
}

// ------------------------------------------------------------------------------------------------
// Data parallel over NVLink peer memory: reduce + Adam + broadcast of the tables in ONE kernel
// ------------------------------------------------------------------------------------------------
// With N replicas the step used to end in all_reduce(98 MB gradient arena) (0.38 ms on 8 B200s) followed by every rank running
// the same Adam pass over all 6.1 M entries (0.135 ms).  Here every rank owns a contiguous 1/N of the entries.  For each entry of
// its shard it loads the N ranks' gradient vectors straight from their arenas over NVLink (peer pointers from a symmetric-memory
// rendezvous), adds them in rank order -- a fixed order, so every replica receives the same bits -- runs Adam with its shard of the
// moments, and stores the new fp32 entry and its fp16 shadow into every rank's tables.  Per rank: 7/8 of 98 MB in, 7/8 of the
// TOUCHED entries x 24 B out, both directions of the links in use at once, and no second pass over the parameters.
// Ordering is the caller's: a cross-rank barrier before (all scatters done) and after (all shards written, all arenas read) the
// launch; the local arena is cleared after the second barrier (its entries are read by their owners only).
constexpr int kMaxPeers = 16;
struct PeerTables {
    const float4 *grad[kMaxPeers];
    float2 *ps[kMaxPeers], *pc[kMaxPeers];
    uint8_t *shadow[kMaxPeers];
};

// E entries per thread (consecutive 256-entry slabs of a CTA's span), all E x world gradient loads issued before the first add: a
// remote load takes microseconds, so small worlds (few loads per entry) need more entries per thread to keep the links busy
// (world 2: 0.154 ms with E = 1 for half of the table -- slower than the whole local Adam pass)
template <int E, int WMAX>      // WMAX: the largest world this instantiation serves (bounds the register arrays)
__global__ void __launch_bounds__(256)
k_peer_adam_tables(const PeerTables P, uint32_t world, uint32_t self, float4 *__restrict__ m4, float4 *__restrict__ v4, uint32_t shadow_stride, size_t e0, size_t e1,
                   float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps, float gscale) {
    const size_t base = e0 + (size_t)blockIdx.x * (256 * E) + threadIdx.x;
    float4 gr[E][WMAX];
#pragma unroll
    for (int u = 0; u < E; u++) {
        const size_t i = base + (size_t)u * 256;
#pragma unroll
        for (int r = 0; r < WMAX; r++)
            if (r < (int)world && i < e1) gr[u][r] = __ldcg(P.grad[r] + i);
    }
#pragma unroll
    for (int u = 0; u < E; u++) {
        const size_t i = base + (size_t)u * 256;
        if (i >= e1) continue;
        float4 g = gr[u][0];
#pragma unroll
        for (int r = 1; r < WMAX; r++)
            if (r < (int)world) { g.x += gr[u][r].x; g.y += gr[u][r].y; g.z += gr[u][r].z; g.w += gr[u][r].w; }
        float4 m = m4[i];
        const bool gz = (g.x == 0.0f) & (g.y == 0.0f) & (g.z == 0.0f) & (g.w == 0.0f);
        const bool mz = (m.x == 0.0f) & (m.y == 0.0f) & (m.z == 0.0f) & (m.w == 0.0f);
        if (gz && mz) continue;      // never touched on any rank: dense Adam leaves it exactly unchanged (see k_adam_tables)
        float4 v = v4[i];
        const float2 s = P.ps[self][i], c = P.pc[self][i];      // replicas hold identical parameters: read the local copy
        const float *G = &g.x;
        float *Mv = &m.x, *V = &v.x;
        float Q[4] = {s.x, s.y, c.x, c.y};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float x = G[k] * gscale;
            Mv[k] = b1 * Mv[k] + (1.0f - b1) * x;
            V[k] = b2 * V[k] + (1.0f - b2) * x * x;
            Q[k] -= lr_over_bc1 * Mv[k] / (sqrtf(V[k]) * inv_sqrt_bc2 + eps);
        }
        m4[i] = m; v4[i] = v;
        const __half2 a = __floats2half2_rn(Q[0], Q[1]), b = __floats2half2_rn(Q[2], Q[3]);
        const uint2 sh = make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&b));
#pragma unroll
        for (int r = 0; r < WMAX; r++) {
            if (r < (int)world) {
                P.ps[r][i] = make_float2(Q[0], Q[1]);
                P.pc[r][i] = make_float2(Q[2], Q[3]);
                *reinterpret_cast<uint2 *>(P.shadow[r] + i * shadow_stride) = sh;
            }
        }
    }
}

struct PeerVec { const float *src[kMaxPeers]; };
// out[i] = sum over ranks (in the order given) of src[r][i]: the MLP's 12 K weight gradients, summed by every rank for itself
__global__ void k_peer_sum(const PeerVec P, uint32_t world, float *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = __ldcg(P.src[0] + i);
    for (uint32_t r = 1; r < world; r++) a += __ldcg(P.src[r] + i);
    out[i] = a;
}

size_t fwd_smem() { return 1024 + 2 * kTileBytes + 3 * kWTile + 2 * kOTile; }
size_t bwd_smem() { return 1024 + 2 * kSetTiles * kTileBytes + 3 * kWTile + 2 * kOTile; }

int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

// table: interleaved fp16 entries {s0,s1,c0,c1} at table + idx * table_stride (8 = stand-alone table4, 16 = one half of a
// paired table: pass the pointer already offset by 0 | 8 bytes); feats [M,64] fp16 (sigma feats 0..31, colour feats 32..63)
S3D_API int s3d_ngp_encode(const float *xyz, uint32_t M, float bound, const void *table, uint32_t table_stride, const int *offsets,
                           uint32_t L, float S, uint32_t H, void *feats, int sigma_only, void *stream) {
    if (M == 0) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
    if (table_stride != 8 && table_stride != 16) return S3D_EINVAL;
k_ngp_encode<<<div_up(M, 256u), 256, 0, as_stream(stream)>>>(xyz, M, bound, (const uint8_t *)table, table_stride, offsets, L, S, H, (__half *)feats, sigma_only);
    S3D_RETURN_LAST();
}

// paired table8 [N] x {teacher s0,s1,c0,c1 | student s0,s1,c0,c1} fp16 (16 bytes).  xyz_teacher / mask may be NULL (no proxy).
S3D_API int s3d_ngp_encode_pair(const float *xyz, const float *xyz_teacher, const uint8_t *mask, uint32_t M, float bound, const void *table8,
                                const int *offsets, uint32_t L, float S, uint32_t H, void *feats_teacher, void *feats_student, void *stream) {
    if (M == 0) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
k_ngp_encode_pair<<<div_up(M, 256u), 256, 0, as_stream(stream)>>>(xyz, xyz_teacher, xyz_teacher ? mask : nullptr, M, bound, (const uint4 *)table8, offsets, L, S, H,
                                                                     (__half *)feats_teacher, (__half *)feats_student);
    S3D_RETURN_LAST();
}

S3D_API int s3d_ngp_pair_tables(const void *teacher_table4, const void *student_table4, void *table8, uint64_t n_entries, void *stream) {
    if (n_entries == 0) return 0;
    k_pair_tables<<<(unsigned)div_up((size_t)n_entries, (size_t)256), 256, 0, as_stream(stream)>>>((const uint2 *)teacher_table4, (const uint2 *)student_table4, (uint4 *)table8, (size_t)n_entries);
    S3D_RETURN_LAST();
}

// grad4: interleaved fp32 gradient table [N] x {s0,s1,c0,c1}, accumulated into
S3D_API int s3d_ngp_scatter(const float *xyz, const void *dfeats, uint32_t M, float bound, float *grad4, const int *offsets, uint32_t L,
                            float S, uint32_t H, float grad_scale, void *stream) {
    if (M == 0) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
    // 128-thread CTAs: 1.92 ms against 1.98 with 256 (the kernel is warp-granular; smaller CTAs drain and refill the SMs more evenly)
    static const uint32_t blk = [] { const char *e = getenv("S3D_SCATTER_BLOCK"); const uint32_t v = e ? (uint32_t)atoi(e) : 128u; return (v >= 32 && v <= 256 && v % 32 == 0) ? v : 128u; }();
    k_ngp_scatter<false><<<div_up(M, blk), blk, 0, as_stream(stream)>>>(xyz, (const __half *)dfeats, M, bound, (float4 *)grad4, offsets, L, S, H, grad_scale, 0, 4, nullptr);
    S3D_RETURN_LAST();
}

// the same scatter restricted to levels [level_begin, level_end) (both multiples of 4): data-parallel training launches the
// levels in a few chunks so that the all-reduce of a finished chunk's slice of the gradient table runs under the next chunk
S3D_API int s3d_ngp_scatter_levels(const float *xyz, const void *dfeats, uint32_t M, float bound, float *grad4, const int *offsets, uint32_t L,
                                   float S, uint32_t H, float grad_scale, uint32_t level_begin, uint32_t level_end, void *stream) {
    if (M == 0 || level_begin >= level_end) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
    if ((level_begin & 3u) || ((level_end & 3u) && level_end != L) || level_end > L) return S3D_EINVAL;
    k_ngp_scatter<false><<<div_up(M, 256u), 256, 0, as_stream(stream)>>>(xyz, (const __half *)dfeats, M, bound, (float4 *)grad4, offsets, L, S, H, grad_scale,
                                                                           level_begin / 4, div_up(level_end, 4u), nullptr);
    S3D_RETURN_LAST();
}

// Deterministic mode.  fixed4: [N] x {s0,s1,c0,c1} as 64-bit fixed point (2^-32), accumulated into with integer reductions;
// nonfinite: one device word, set when a contribution was inf / NaN / out of range.
S3D_API int s3d_ngp_scatter_fixed(const float *xyz, const void *dfeats, uint32_t M, float bound, long long *fixed4, const int *offsets, uint32_t L,
                                  float S, uint32_t H, float grad_scale, uint32_t *nonfinite, void *stream) {
    if (M == 0) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
    if (!fixed4 || !nonfinite) return S3D_EINVAL;
    k_ngp_scatter<true><<<div_up(M, 256u), 256, 0, as_stream(stream)>>>(xyz, (const __half *)dfeats, M, bound, (float4 *)fixed4, offsets, L, S, H, grad_scale, 0, 4,
                                                                          nonfinite);
    S3D_RETURN_LAST();
}

// count[0] += the number of global reductions s3d_ngp_scatter issues for these samples, count[1] += those on hashed levels
// (pseudo-random addresses); device, 64-bit each (measurement aid)
S3D_API int s3d_ngp_scatter_count(const float *xyz, uint32_t M, float bound, const int *offsets, uint32_t L, float S, uint32_t H,
                                  unsigned long long *count, void *stream) {
    if (M == 0) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
    if (!count) return S3D_EINVAL;
    k_ngp_scatter_count<<<div_up(M, 256u), 256, 0, as_stream(stream)>>>(xyz, M, bound, offsets, L, S, H, count);
    S3D_RETURN_LAST();
}

// grad[i] += fixed[i] * 2^-32, fixed[i] = 0 (n values); a raised *nonfinite becomes a NaN in grad[0] and is cleared
S3D_API int s3d_fixed_to_float(long long *fixed, float *grad, size_t n, uint32_t *nonfinite, void *stream) {
    if (n == 0) return 0;
    if (!fixed || !grad || !nonfinite) return S3D_EINVAL;
    k_fixed_to_float<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(fixed, grad, n, nonfinite);
    S3D_RETURN_LAST();
}

// weights: fp16 row-major nn.Linear matrices sigma_net.0 [64,32], sigma_net.1 [16,64], color_net.0 [64,63], .1 [64,64], .2 [3,64]
S3D_API int s3d_ngp_mlp_forward(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                                const void *w_c1, const void *w_c2, float density_scale, float *sigma, float *rgb, float *geo,
                                int sigma_only, void *stream) {
    if (M == 0) return 0;
    const size_t smem = fwd_smem();
    cudaError_t e = cudaFuncSetAttribute(k_ngp_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    FwdArgs a;
    a.feats = (const __half *)feats; a.dirs = dirs;
    a.w = MlpWeights{(const __half *)w_s0, (const __half *)w_s1, (const __half *)w_c0, (const __half *)w_c1, (const __half *)w_c2};
    a.sigma = sigma; a.rgb = rgb; a.geo = geo; a.M = M; a.n_tiles = div_up(M, kRows); a.density_scale = density_scale; a.sigma_only = sigma_only;
    // S3D_MLP_FWD=ss selects the older kernel with activation tiles in shared memory (A/B runs); default = activations in TMEM
    static const bool use_ss = [] { const char *e = getenv("S3D_MLP_FWD"); return e && e[0] == 's' && e[1] == 's'; }();
    if (use_ss) {
        const uint32_t grid = min(a.n_tiles, (uint32_t)sm_count() * 3u);
        k_ngp_mlp_fwd<<<grid, kMlpThreads, smem, as_stream(stream)>>>(a);
        S3D_RETURN_LAST();
    }
    const size_t smem_ts = 1024 + 3 * kWTile + 2 * kOTile;
    e = cudaFuncSetAttribute(k_ngp_mlp_fwd_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts);
    if (e != cudaSuccess) return (int)e;
    static const uint32_t per_sm = [] { const char *e = getenv("S3D_MLP_FWD_CTAS"); const int v = e ? atoi(e) : 3; return (uint32_t)(v >= 1 && v <= 4 ? v : 3); }();
    const uint32_t grid = min(a.n_tiles, (uint32_t)sm_count() * per_sm);
    k_ngp_mlp_fwd_ts<<<grid, kMlpThreads, smem_ts, as_stream(stream)>>>(a);
    S3D_RETURN_LAST();
}

// teacher + student forward on one sample buffer: gather from the paired table + both MLPs in one kernel (k_ngp_pair_fwd).
// xyz_teacher / mask / dirs_teacher: the proxy-mapped positions / moved flags / mapped directions (xyz_teacher and mask may be
// NULL: no sample moved; dirs_teacher NULL = dirs).  weights_*: the five fp16 nn.Linear matrices of each model.
S3D_API int s3d_ngp_pair_forward(const float *xyz, const float *xyz_teacher, const uint8_t *mask, const float *dirs, const float *dirs_teacher,
                                 uint32_t M, float bound, const void *table8, const int *offsets, uint32_t L, float S, uint32_t H,
                                 const void *t_s0, const void *t_s1, const void *t_c0, const void *t_c1, const void *t_c2, const void *s_s0,
                                 const void *s_s1, const void *s_c0, const void *s_c1, const void *s_c2, float density_scale_teacher,
                                 float density_scale_student, float *sigma_t, float *rgb_t, float *sigma_s, float *rgb_s, void *feats_student,
                                 void *stream) {
    if (M == 0) return 0;
    if (L > kMaxLevels) return S3D_ENOTSUP;
    const size_t smem = 1024 + 2 * (3 * kWTile + 2 * kOTile);
    cudaError_t e = cudaFuncSetAttribute(k_ngp_pair_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    PairArgs a;
    a.xyz = xyz; a.xyz_teacher = xyz_teacher; a.mask = xyz_teacher ? mask : nullptr; a.dirs = dirs; a.dirs_teacher = dirs_teacher ? dirs_teacher : dirs;
    a.table8 = (const uint4 *)table8; a.offsets = offsets;
    a.wt = MlpWeights{(const __half *)t_s0, (const __half *)t_s1, (const __half *)t_c0, (const __half *)t_c1, (const __half *)t_c2};
    a.ws = MlpWeights{(const __half *)s_s0, (const __half *)s_s1, (const __half *)s_c0, (const __half *)s_c1, (const __half *)s_c2};
    a.sigma_t = sigma_t; a.rgb_t = rgb_t; a.sigma_s = sigma_s; a.rgb_s = rgb_s; a.feats_s = (__half *)feats_student;
    a.M = M; a.n_tiles = div_up(M, kRows); a.L = L; a.H = H; a.bound = bound; a.S = S; a.dscale_t = density_scale_teacher; a.dscale_s = density_scale_student;
    const uint32_t grid = min(div_up(a.n_tiles, kPairSlots), (uint32_t)sm_count());
    k_ngp_pair_fwd<<<grid, kPairThreads, smem, as_stream(stream)>>>(a);
    S3D_RETURN_LAST();
}

#ifdef S3D_TRACE
S3D_API int s3d_debug_trace(long long *host_out, int n) { return (int)cudaMemcpyFromSymbol(host_out, g_trace, (size_t)n * sizeof(long long)); }
#endif

static int mlp_backward_launch(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                               const void *w_c1, const void *w_c2, float density_scale, const float *g_sigma, const float *g_rgb,
                               void *dfeats, float out_scale, float *gw_s0, float *gw_s1, float *gw_c0, float *gw_c1, float *gw_c2,
                               int train_mlp, uint32_t *nonfinite, void *stream) {
    if (M == 0) return 0;
    const size_t smem = bwd_smem();
    cudaError_t e = cudaFuncSetAttribute(nonfinite ? k_ngp_mlp_bwd<true> : k_ngp_mlp_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    BwdArgs a;
    a.feats = (const __half *)feats; a.dirs = dirs;
    a.w = MlpWeights{(const __half *)w_s0, (const __half *)w_s1, (const __half *)w_c0, (const __half *)w_c1, (const __half *)w_c2};
    a.g_sigma = g_sigma; a.g_rgb = g_rgb; a.dfeats = (__half *)dfeats;
    a.gw_s0 = gw_s0; a.gw_s1 = gw_s1; a.gw_c0 = gw_c0; a.gw_c1 = gw_c1; a.gw_c2 = gw_c2;
    a.M = M; a.n_tiles = div_up(M, kRows); a.density_scale = density_scale; a.out_scale = out_scale; a.train_mlp = train_mlp;
    a.nonfinite = nonfinite;
    const uint32_t grid = min(a.n_tiles, (uint32_t)sm_count());
    if (nonfinite) k_ngp_mlp_bwd<true><<<grid, kBwdThreads, smem, as_stream(stream)>>>(a);
    else k_ngp_mlp_bwd<false><<<grid, kBwdThreads, smem, as_stream(stream)>>>(a);
    S3D_RETURN_LAST();
}

S3D_API int s3d_ngp_mlp_backward(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                                 const void *w_c1, const void *w_c2, float density_scale, const float *g_sigma, const float *g_rgb,
                                 void *dfeats, float out_scale, float *gw_s0, float *gw_s1, float *gw_c0, float *gw_c1, float *gw_c2,
                                 int train_mlp, void *stream) {
    return mlp_backward_launch(feats, dirs, M, w_s0, w_s1, w_c0, w_c1, w_c2, density_scale, g_sigma, g_rgb, dfeats, out_scale, gw_s0, gw_s1, gw_c0, gw_c1,
                               gw_c2, train_mlp, nullptr, stream);
}

// Deterministic mode: the weight gradients of the CTAs are added as 64-bit fixed point (2^-32; same element layout as the fp32
// matrices) so that their sum does not depend on the order the CTAs finish in; dfeats is unchanged.
S3D_API int s3d_ngp_mlp_backward_fixed(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                                       const void *w_c1, const void *w_c2, float density_scale, const float *g_sigma, const float *g_rgb,
                                       void *dfeats, float out_scale, long long *gw_s0, long long *gw_s1, long long *gw_c0, long long *gw_c1,
                                       long long *gw_c2, int train_mlp, uint32_t *nonfinite, void *stream) {
    if (!nonfinite) return S3D_EINVAL;
    return mlp_backward_launch(feats, dirs, M, w_s0, w_s1, w_c0, w_c1, w_c2, density_scale, g_sigma, g_rgb, dfeats, out_scale, (float *)gw_s0, (float *)gw_s1,
                               (float *)gw_c0, (float *)gw_c1, (float *)gw_c2, train_mlp, nonfinite, stream);
}

S3D_API int s3d_ngp_interleave_tables(const float *table_sigma, const float *table_color, void *table4, uint64_t n_entries, void *stream) {
    if (n_entries == 0) return 0;
    k_interleave_tables<<<(unsigned)div_up((size_t)n_entries, (size_t)256), 256, 0, as_stream(stream)>>>((const float2 *)table_sigma, (const float2 *)table_color, (uint2 *)table4, (size_t)n_entries);
    S3D_RETURN_LAST();
}

// Adam (torch semantics, main_SealNeRF.py:283-284) over both tables at once: params fp32 [N,2] each, interleaved
// grad / moments fp32 [N,4]; the fp16 shadow entry {s0,s1,c0,c1} of entry i is written at shadow + i * shadow_stride
// (8 = table4, 16 = the student half of a paired table8) in the same pass; the gradient is zeroed.
S3D_API int s3d_ngp_adam_tables(float *table_sigma, float *table_color, float *grad4, float *exp_avg4, float *exp_avg_sq4, void *shadow,
                                uint32_t shadow_stride, uint64_t n_entries, float lr, float beta1, float beta2, float eps, uint32_t step,
                                float grad_scale, const float *scaler_state, void *stream) {
    if (n_entries == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    k_adam_tables<<<(unsigned)div_up((size_t)n_entries, (size_t)256), 256, 0, as_stream(stream)>>>(
        (float2 *)table_sigma, (float2 *)table_color, (float4 *)grad4, (float4 *)exp_avg4, (float4 *)exp_avg_sq4, (uint8_t *)shadow, shadow_stride, (size_t)n_entries,
        (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps, grad_scale, scaler_state, lr);
    S3D_RETURN_LAST();
}

// Data parallel, peer-memory form of s3d_ngp_adam_tables (see k_peer_adam_tables).  *_peers: HOST arrays of `world` device
// pointers in rank order (the gradient sum is taken in array order: the same order on every rank gives every replica the same
// bits); `rank` = this process's index in them (its own tables are the ones read).  Entries [entry_begin, entry_end) are this rank's shard; exp_avg4 / exp_avg_sq4
// are local full-size moment tables of which only the shard is used.  grad_scale divides out the loss scale and the world size.
// The gradient arenas are NOT cleared (the caller clears its own after the closing barrier).
S3D_API int s3d_ngp_peer_adam_tables(const void *const *grad4_peers, void *const *sigma_peers, void *const *color_peers, void *const *shadow_peers,
                                     uint32_t world, uint32_t rank, float *exp_avg4, float *exp_avg_sq4, uint32_t shadow_stride, uint64_t entry_begin,
                                     uint64_t entry_end, float lr, float beta1, float beta2, float eps, uint32_t step, float grad_scale, void *stream) {
    if (entry_end <= entry_begin) return 0;
    if (rank >= world) return S3D_EINVAL;
    if (world == 0 || world > (uint32_t)kMaxPeers || !grad4_peers || !sigma_peers || !color_peers || !shadow_peers) return S3D_EINVAL;
    PeerTables P;
    for (uint32_t r = 0; r < (uint32_t)kMaxPeers; r++) {
        const uint32_t q = r < world ? r : 0;
        P.grad[r] = (const float4 *)grad4_peers[q]; P.ps[r] = (float2 *)sigma_peers[q]; P.pc[r] = (float2 *)color_peers[q]; P.shadow[r] = (uint8_t *)shadow_peers[q];
        if (!P.grad[r] || !P.ps[r] || !P.pc[r] || !P.shadow[r]) return S3D_EINVAL;
    }
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const size_t n = (size_t)(entry_end - entry_begin);
    const float a1 = (float)((double)lr / bc1), a2 = (float)(1.0 / sqrt(bc2));
    cudaStream_t st = as_stream(stream);
    // entries per thread: about 8 gradient loads in flight per thread whatever the world size
    if (world <= 2) k_peer_adam_tables<4, 2><<<(unsigned)div_up(n, (size_t)1024), 256, 0, st>>>(P, world, rank, (float4 *)exp_avg4, (float4 *)exp_avg_sq4, shadow_stride, (size_t)entry_begin, (size_t)entry_end, a1, a2, beta1, beta2, eps, grad_scale);
    else if (world <= 4) k_peer_adam_tables<2, 4><<<(unsigned)div_up(n, (size_t)512), 256, 0, st>>>(P, world, rank, (float4 *)exp_avg4, (float4 *)exp_avg_sq4, shadow_stride, (size_t)entry_begin, (size_t)entry_end, a1, a2, beta1, beta2, eps, grad_scale);
    else if (world <= 8) k_peer_adam_tables<1, 8><<<(unsigned)div_up(n, (size_t)256), 256, 0, st>>>(P, world, rank, (float4 *)exp_avg4, (float4 *)exp_avg_sq4, shadow_stride, (size_t)entry_begin, (size_t)entry_end, a1, a2, beta1, beta2, eps, grad_scale);
    else k_peer_adam_tables<1, kMaxPeers><<<(unsigned)div_up(n, (size_t)256), 256, 0, st>>>(P, world, rank, (float4 *)exp_avg4, (float4 *)exp_avg_sq4, shadow_stride, (size_t)entry_begin, (size_t)entry_end, a1, a2, beta1, beta2, eps, grad_scale);
    S3D_RETURN_LAST();
}

// out[i] = src_peers[0][i] + src_peers[1][i] + ... (array order), n floats; src_peers: HOST array of `world` device pointers
S3D_API int s3d_peer_sum(const void *const *src_peers, uint32_t world, float *out, uint64_t n, void *stream) {
    if (n == 0) return 0;
    if (world == 0 || world > (uint32_t)kMaxPeers || !src_peers || !out) return S3D_EINVAL;
    PeerVec P;
    for (uint32_t r = 0; r < (uint32_t)kMaxPeers; r++) { P.src[r] = (const float *)src_peers[r < world ? r : 0]; if (!P.src[r]) return S3D_EINVAL; }
    k_peer_sum<<<(unsigned)div_up((size_t)n, (size_t)256), 256, 0, as_stream(stream)>>>(P, world, out, (size_t)n);
    S3D_RETURN_LAST();
}
