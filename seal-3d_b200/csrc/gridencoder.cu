// Multiresolution hash / tiled grid encoding for sm_100a.
//
// Replaces gridencoder/src/gridencoder.cu of the reference (entry points gridencoder.h:12-15).
// What is different from the reference kernels (gridencoder.cu:88-242, :246-337):
//   * one thread owns a point for ALL levels (inputs read once, 16 independent gather groups in
//     flight) when the batch is large; small batches keep one (point, level) per thread so the
//     grid still covers the 148 SMs;
//   * every corner is fetched with ONE vector load of the whole C-channel entry (8 B for fp32 C=2,
//     4 B for fp16 C=2) through the read-only path -- the reference issues one scalar load per
//     channel (SASS: 32-bit LDG.E.CONSTANT / LDG.E.U16);
//   * fp16 tables are accumulated in fp32 and rounded once (the reference rounds to half after
//     every corner);
//   * the backward scatter first reduces inside the warp: samples arrive ray-major, so lanes that
//     fall in the same cell of a level form contiguous runs; a segmented shuffle reduction folds
//     each run into its head lane, which then issues one vector RED per corner
//     (red.global.add.v2.f32 / .noftz.f16x2) instead of one scalar atomic per (lane, channel);
//   * launches go to the caller's stream.
// Level constants follow gridencoder.cu:137-139 exactly: scale = exp2f(l*S)*H - 1 (one FMA),
// resolution = ceil(scale) + 1, hashmap_size = offsets[l+1] - offsets[l].
#include "common.cuh"

namespace {

__device__ __constant__ uint32_t kPrimes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};

struct LevelInfo {
    uint32_t hashmap_size, resolution;
    float scale;
    bool hashed;  // gridtype==hash and the dense index range exceeds the level's table
    bool pair_ok; // level starts at an even entry and holds an even number of entries: aligned entry pairs stay inside it
};

template <uint32_t D>
__device__ __forceinline__ LevelInfo level_info(const int *__restrict__ offsets, uint32_t level, float S, uint32_t H,
                                                uint32_t gridtype, bool align_corners) {
    LevelInfo li;
    const uint32_t first = (uint32_t)__ldg(offsets + level);
    li.hashmap_size = (uint32_t)__ldg(offsets + level + 1) - first;
    li.pair_ok = ((first | li.hashmap_size) & 1u) == 0;
    li.scale = __fmaf_rn(exp2f((float)level * S), (float)H, -1.0f);
    li.resolution = (uint32_t)ceilf(li.scale) + 1;
    // stride after the d-loop of get_grid_index (gridencoder.cu:72-75): stops multiplying once it exceeds the table
    uint32_t stride = 1;
#pragma unroll
    for (uint32_t d = 0; d < D; d++)
        if (stride <= li.hashmap_size) stride *= align_corners ? li.resolution : (li.resolution + 1);
    li.hashed = (gridtype == 0) && (stride > li.hashmap_size);
    return li;
}

// entry index (not yet multiplied by C) of one corner -- gridencoder.cu:51-84
template <uint32_t D>
__device__ __forceinline__ uint32_t corner_index(const LevelInfo &li, bool align_corners, const uint32_t (&pg)[D]) {
    uint32_t index = 0;
    if (li.hashed) {
#pragma unroll
        for (uint32_t d = 0; d < D; d++) index ^= pg[d] * kPrimes[d];
    } else {
        uint32_t stride = 1;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if (stride <= li.hashmap_size) {
                index += pg[d] * stride;
                stride *= align_corners ? li.resolution : (li.resolution + 1);
            }
        }
    }
    return index % li.hashmap_size;
}

template <typename T, uint32_t C> struct Entry;  // one table entry = C channels, loaded / reduced as one vector
template <> struct Entry<float, 1> { using V = float; };
template <> struct Entry<float, 2> { using V = float2; };
template <> struct Entry<float, 4> { using V = float4; };
template <> struct Entry<float, 8> { using V = float4; };
template <> struct Entry<__half, 1> { using V = __half; };
template <> struct Entry<__half, 2> { using V = __half2; };
template <> struct Entry<__half, 4> { using V = uint2; };
template <> struct Entry<__half, 8> { using V = uint4; };

template <typename T, uint32_t C>
__device__ __forceinline__ void load_entry(const T *__restrict__ grid, uint32_t index, float (&v)[C]) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (C == 1) { v[0] = __ldg(grid + index); }
        else if constexpr (C == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(grid) + index); v[0] = t.x; v[1] = t.y; }
        else {
#pragma unroll
            for (uint32_t q = 0; q < C / 4; q++) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(grid) + (size_t)index * (C / 4) + q);
                v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
            }
        }
    } else {
        if constexpr (C == 1) { v[0] = __half2float(__ldg(reinterpret_cast<const __half *>(grid) + index)); }
        else if constexpr (C == 2) {
            const float2 t = __half22float2(__ldg(reinterpret_cast<const __half2 *>(grid) + index));
            v[0] = t.x; v[1] = t.y;
        } else if constexpr (C == 4) {
            const uint2 t = __ldg(reinterpret_cast<const uint2 *>(grid) + index);
            const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
            const uint4 t = __ldg(reinterpret_cast<const uint4 *>(grid) + index);
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
                v[q * 2] = a.x; v[q * 2 + 1] = a.y;
            }
        }
    }
}

// Two entries that are the halves of one aligned entry pair, fetched with ONE double-width load.  Along x the two corners
// of a cell are such a pair whenever the first index is even and the second is its successor: on hashed levels
// (x ^ h) and ((x+1) ^ h) differ in bit 0 only for even x, on dense levels the index is x + ... itself.  The L1 tag
// stage pays per (lane, sector), so a merged pair costs one lookup instead of two.
template <typename T, uint32_t C> struct PairLoad { static constexpr bool kOk = false; };
template <> struct PairLoad<float, 2> {
    static constexpr bool kOk = true;
    __device__ static __forceinline__ void load(const float *grid, uint32_t even_index, float (&lo)[2], float (&hi)[2]) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(grid) + (even_index >> 1));
        lo[0] = t.x; lo[1] = t.y; hi[0] = t.z; hi[1] = t.w;
    }
};
template <> struct PairLoad<__half, 2> {
    static constexpr bool kOk = true;
    __device__ static __forceinline__ void load(const __half *grid, uint32_t even_index, float (&lo)[2], float (&hi)[2]) {
        const uint2 t = __ldg(reinterpret_cast<const uint2 *>(grid) + (even_index >> 1));
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
        lo[0] = a.x; lo[1] = a.y; hi[0] = b.x; hi[1] = b.y;
    }
};
template <> struct PairLoad<__half, 4> {
    static constexpr bool kOk = true;
    __device__ static __forceinline__ void load(const __half *grid, uint32_t even_index, float (&lo)[4], float (&hi)[4]) {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(grid) + (even_index >> 1));
        const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&t.y));
        const float2 c = __half22float2(*reinterpret_cast<const __half2 *>(&t.z)), d = __half22float2(*reinterpret_cast<const __half2 *>(&t.w));
        lo[0] = a.x; lo[1] = a.y; lo[2] = b.x; lo[3] = b.y; hi[0] = c.x; hi[1] = c.y; hi[2] = d.x; hi[3] = d.y;
    }
};

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }

template <typename T, uint32_t C, bool CS = false>
__device__ __forceinline__ void store_vec(T *__restrict__ p, const float (&v)[C]) {
    if constexpr (sizeof(T) == 4 && C == 2) {
        if constexpr (CS) __stcs(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
        else *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else if constexpr (sizeof(T) == 2 && C == 2) {
        if constexpr (CS) __stcs(reinterpret_cast<__half2 *>(p), __floats2half2_rn(v[0], v[1]));
        else *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(v[0], v[1]);
    }
    else {
#pragma unroll
        for (uint32_t c = 0; c < C; c++) p[c] = from_float<T>(v[c]);
    }
}

// position of a point inside one level: integer cell + fractional weights (gridencoder.cu:146-156)
template <uint32_t D>
__device__ __forceinline__ void locate(const float (&x)[D], const LevelInfo &li, bool align_corners, uint32_t interp,
                                       float (&pos)[D], float (&deriv)[D], uint32_t (&pg)[D]) {
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const float p = __fmaf_rn(x[d], li.scale, align_corners ? 0.0f : 0.5f);
        const float fl = floorf(p);
        pg[d] = (uint32_t)fl;
        float f = __fsub_rn(p, (float)pg[d]);
        // `float pos_deriv[D] = {1.0f}` in the reference (gridencoder.cu:143) sets element 0 only: with linear interpolation
        // dy_dx is non-zero for the first coordinate alone.  Reproduced (outputs of the reference kernel: tests/golden/gpu_ref.npz).
        deriv[d] = d == 0 ? 1.0f : 0.0f;
        if (interp == 1) {
            deriv[d] = 6 * f * (1.0f - f);
            f = f * f * (3.0f - 2.0f * f);
        }
        pos[d] = f;
    }
}

template <uint32_t D>
__device__ __forceinline__ bool out_of_range(const float (&x)[D]) {
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) oob |= (x[d] < 0.0f) || (x[d] > 1.0f);
    return oob;
}

// encode one (point, level): result[C] (fp32) and optionally dy_dx
template <typename T, uint32_t D, uint32_t C>
__device__ __forceinline__ void encode_level(const float (&x)[D], const T *__restrict__ grid_level, const LevelInfo &li,
                                             bool align_corners, uint32_t interp, float (&res)[C], T *__restrict__ dy_dx_lvl) {
    float pos[D], deriv[D];
    uint32_t pg[D];
    locate<D>(x, li, align_corners, interp, pos, deriv, pg);
#pragma unroll
    for (uint32_t c = 0; c < C; c++) res[c] = 0.0f;
    float val[1u << D][C];
    if constexpr (PairLoad<T, C>::kOk) {
        if (li.pair_ok) {  // uniform per level
            uint32_t i0[1u << (D - 1)], i1[1u << (D - 1)];
            bool merged[1u << (D - 1)];
#pragma unroll
            for (uint32_t j = 0; j < (1u << (D - 1)); j++) {
                uint32_t pl[D];
                pl[0] = pg[0];
#pragma unroll
                for (uint32_t d = 1; d < D; d++) pl[d] = pg[d] + ((j >> (d - 1)) & 1u);
                i0[j] = corner_index<D>(li, align_corners, pl);
                pl[0] = pg[0] + 1;
                i1[j] = corner_index<D>(li, align_corners, pl);
                merged[j] = (i0[j] ^ i1[j]) == 1u;
            }
            // all loads first (the second one predicated off for merged pairs), then the selects
            float lo[1u << (D - 1)][C], hi[1u << (D - 1)][C];
#pragma unroll
            for (uint32_t j = 0; j < (1u << (D - 1)); j++) {
                PairLoad<T, C>::load(grid_level, i0[j] & ~1u, lo[j], hi[j]);
                if (!merged[j]) load_entry<T, C>(grid_level, i1[j], val[2 * j + 1]);
            }
#pragma unroll
            for (uint32_t j = 0; j < (1u << (D - 1)); j++) {
                const bool odd = i0[j] & 1u;
#pragma unroll
                for (uint32_t c = 0; c < C; c++) {
                    val[2 * j][c] = odd ? hi[j][c] : lo[j][c];
                    if (merged[j]) val[2 * j + 1][c] = odd ? lo[j][c] : hi[j][c];
                }
            }
        } else {
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                uint32_t pl[D];
#pragma unroll
                for (uint32_t d = 0; d < D; d++) pl[d] = pg[d] + ((idx >> d) & 1u);
                load_entry<T, C>(grid_level, corner_index<D>(li, align_corners, pl), val[idx]);
            }
        }
    } else {
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); idx++) {
            uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) pl[d] = pg[d] + ((idx >> d) & 1u);
            load_entry<T, C>(grid_level, corner_index<D>(li, align_corners, pl), val[idx]);
        }
    }
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = 1.0f;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) w = __fmul_rn(w, ((idx >> d) & 1u) ? pos[d] : __fsub_rn(1.0f, pos[d]));
#pragma unroll
        for (uint32_t c = 0; c < C; c++) res[c] = __fmaf_rn(w, val[idx][c], res[c]);
    }
    if (dy_dx_lvl) {  // [D, C] block of dy_dx[B, L, D, C]  (gridencoder.cu:198-241)
#pragma unroll
        for (uint32_t gd = 0; gd < D; gd++) {
            float rg[C];
#pragma unroll
            for (uint32_t c = 0; c < C; c++) rg[c] = 0.0f;
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                if ((idx >> gd) & 1u) continue;  // enumerate pairs (left = idx, right = idx | 1<<gd)
                float w = li.scale;
#pragma unroll
                for (uint32_t d = 0; d < D; d++)
                    if (d != gd) w *= ((idx >> d) & 1u) ? pos[d] : (1.0f - pos[d]);
#pragma unroll
                for (uint32_t c = 0; c < C; c++) rg[c] += w * (val[idx | (1u << gd)][c] - val[idx][c]) * deriv[gd];
            }
#pragma unroll
            for (uint32_t c = 0; c < C; c++) dy_dx_lvl[gd * C + c] = from_float<T>(rg[c]);
        }
    }
}

// ALL_LEVELS: blockIdx.y unused, thread loops over levels.  Otherwise blockIdx.y = level.
template <typename T, uint32_t D, uint32_t C, bool ALL_LEVELS, int UNROLL = 2, bool CS = false, int MAXT = 256>
__global__ void __launch_bounds__(MAXT)
k_grid_forward(const float *__restrict__ inputs, const T *__restrict__ grid, const int *__restrict__ offsets,
               T *__restrict__ outputs, uint32_t B, uint32_t L, float S, uint32_t H, T *__restrict__ dy_dx,
               uint32_t gridtype, bool align_corners, uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) x[d] = __ldg(inputs + (size_t)b * D + d);
    const bool oob = out_of_range<D>(x);
    const uint32_t l0 = ALL_LEVELS ? 0 : blockIdx.y, l1 = ALL_LEVELS ? L : blockIdx.y + 1;
#pragma unroll UNROLL
    for (uint32_t level = l0; level < l1; level++) {
        float res[C];
        T *dd = dy_dx ? dy_dx + ((size_t)b * L + level) * D * C : nullptr;
        if (oob) {
#pragma unroll
            for (uint32_t c = 0; c < C; c++) res[c] = 0.0f;
            if (dd)
                for (uint32_t i = 0; i < D * C; i++) dd[i] = from_float<T>(0.0f);
        } else {
            const LevelInfo li = level_info<D>(offsets, level, S, H, gridtype, align_corners);
            encode_level<T, D, C>(x, grid + (size_t)(uint32_t)__ldg(offsets + level) * C, li, align_corners, interp, res, dd);
        }
        store_vec<T, C, CS>(outputs + ((size_t)level * B + b) * C, res);
    }
}

// ---- backward ----------------------------------------------------------------------------

template <typename T, uint32_t C>
__device__ __forceinline__ void red_add_entry(T *__restrict__ p, const float (&v)[C]) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (C == 1) atomicAdd(p, v[0]);
        else if constexpr (C == 2) atomicAdd(reinterpret_cast<float2 *>(p), make_float2(v[0], v[1]));
        else {
#pragma unroll
            for (uint32_t q = 0; q < C / 4; q++)
                atomicAdd(reinterpret_cast<float4 *>(p) + q, make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]));
        }
    } else {
        // fp16 gradient table (the reference's AMP path accumulates in half too, gridencoder.cu:321-334): fire-and-forget
        // reductions written in PTX -- atomicAdd(__half2 *) compiles to the RETURNING form (ATOM.E.ADD.F16x2), which made this
        // path slower than the fp32 one (4.30 vs 2.69 ms for 2^22 points)
        if constexpr (C == 1) {
            const __half h = __float2half_rn(v[0]);
            asm volatile("red.global.add.noftz.f16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<const unsigned short *>(&h)) : "memory");
        } else {
            uint32_t w[C / 2];
#pragma unroll
            for (uint32_t q = 0; q < C / 2; q++) {
                const __half2 h = __floats2half2_rn(v[q * 2], v[q * 2 + 1]);
                w[q] = *reinterpret_cast<const uint32_t *>(&h);
            }
            if constexpr (C == 2) asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
            else if constexpr (C == 4) asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
            else asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
    }
}

// scatter one (point, level) with an in-warp segmented pre-reduction over runs of equal cells.
// All 32 lanes of the warp must call this (inactive lanes pass valid=false).
template <typename T, uint32_t D, uint32_t C>
__device__ __forceinline__ void scatter_level(bool valid, const float (&x)[D], const float (&g)[C], T *__restrict__ gg_level,
                                              const LevelInfo &li, bool align_corners, uint32_t interp) {
    float pos[D], deriv[D];
    uint32_t pg[D] = {0};
    if (valid) locate<D>(x, li, align_corners, interp, pos, deriv, pg);
    else {
#pragma unroll
        for (uint32_t d = 0; d < D; d++) pos[d] = 0.0f;
    }
    // key of the cell: lanes with equal keys address the same 2^D entries
    unsigned long long key = valid ? 0ull : ~0ull;
    if (valid) {
        constexpr uint32_t kBits = 64 / D;  // 32 / 21 / 16 bits per axis: cell coordinates are far below that
#pragma unroll
        for (uint32_t d = 0; d < D; d++) key |= (unsigned long long)pg[d] << (kBits * d);
    }
    const uint32_t lane = lane_id();
    const unsigned long long prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    float wv[1u << D][C];
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = valid ? 1.0f : 0.0f;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) w *= ((idx >> d) & 1u) ? pos[d] : (1.0f - pos[d]);
#pragma unroll
        for (uint32_t c = 0; c < C; c++) wv[idx][c] = w * g[c];
    }
    bool writer = valid;
    if (__popc(heads) <= 16) {  // at least half of the lanes can be merged: segmented suffix reduction to the head lane
        const uint32_t after = heads >> 1 >> lane;  // heads strictly after this lane
        const uint32_t seg_left = after ? (uint32_t)(__ffs(after) - 1) : (31u - lane);  // lanes after me in my run
#pragma unroll
        for (uint32_t dlt = 1; dlt < 32; dlt <<= 1) {
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << D); idx++)
#pragma unroll
                for (uint32_t c = 0; c < C; c++) {
                    const float o = __shfl_down_sync(0xffffffffu, wv[idx][c], dlt);
                    if (dlt <= seg_left) wv[idx][c] += o;
                }
        }
        writer = valid && head;
    }
    if (writer) {
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); idx++) {
            uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) pl[d] = pg[d] + ((idx >> d) & 1u);
            red_add_entry<T, C>(gg_level + (size_t)corner_index<D>(li, align_corners, pl) * C, wv[idx]);
        }
    }
}

template <typename T, uint32_t D, uint32_t C, bool ALL_LEVELS>
__global__ void __launch_bounds__(256)
k_grid_backward(const T *__restrict__ grad, const float *__restrict__ inputs, const int *__restrict__ offsets,
                T *__restrict__ grad_grid, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                bool align_corners, uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;  // whole warps stay alive: no early return
    const bool in_range = b < B;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) x[d] = in_range ? __ldg(inputs + (size_t)b * D + d) : -1.0f;
    const bool valid = in_range && !out_of_range<D>(x);
    const uint32_t l0 = ALL_LEVELS ? 0 : blockIdx.y, l1 = ALL_LEVELS ? L : blockIdx.y + 1;
    for (uint32_t level = l0; level < l1; level++) {
        const LevelInfo li = level_info<D>(offsets, level, S, H, gridtype, align_corners);
        float g[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) g[c] = valid ? to_float<T>(grad[((size_t)level * B + b) * C + c]) : 0.0f;
        scatter_level<T, D, C>(valid, x, g, grad_grid + (size_t)(uint32_t)__ldg(offsets + level) * C, li, align_corners, interp);
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]      (gridencoder.cu:341-366)
template <typename T>
__global__ void k_grid_input_backward(const T *__restrict__ grad, const T *__restrict__ dy_dx, T *__restrict__ grad_inputs,
                                      uint32_t B, uint32_t D, uint32_t C, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    float r = 0.0f;
    for (uint32_t l = 0; l < L; l++)
        for (uint32_t c = 0; c < C; c++)
            r += to_float<T>(grad[((size_t)l * B + b) * C + c]) * to_float<T>(dy_dx[(((size_t)b * L + l) * D + d) * C + c]);
    grad_inputs[t] = from_float<T>(r);
}

// total-variation regulariser gradient (gridencoder.cu:504-607), float tables only in practice
template <typename T, uint32_t D, uint32_t C>
__global__ void k_grad_tv(const T *__restrict__ inputs, const T *__restrict__ grid, T *__restrict__ grad,
                          const int *__restrict__ offsets, float weight, uint32_t B, uint32_t L, float S, uint32_t H,
                          uint32_t gridtype, bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) x[d] = to_float<T>(inputs[(size_t)b * D + d]);
    if (out_of_range<D>(x)) return;
    const LevelInfo li = level_info<D>(offsets, level, S, H, gridtype, align_corners);
    const T *gl = grid + (size_t)(uint32_t)offsets[level] * C;
    T *gg = grad + (size_t)(uint32_t)offsets[level] * C;
    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) pg[d] = (uint32_t)floorf(__fmaf_rn(x[d], li.scale, align_corners ? 0.0f : 0.5f));
    float center[C], res[C], idelta[C];
    const uint32_t index = corner_index<D>(li, align_corners, pg);
    load_entry<T, C>(gl, index, center);
#pragma unroll
    for (uint32_t c = 0; c < C; c++) { res[c] = 0.0f; idelta[c] = 0.0f; }
    const float w = weight / (2 * D);
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t cur = pg[d];
        float nb[C];
        if (cur < li.resolution) {
            pg[d] = cur + 1;
            load_entry<T, C>(gl, corner_index<D>(li, align_corners, pg), nb);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) { const float gv = center[c] - nb[c]; res[c] += gv; idelta[c] += gv * gv; }
        }
        if (cur > 0) {
            pg[d] = cur - 1;
            load_entry<T, C>(gl, corner_index<D>(li, align_corners, pg), nb);
#pragma unroll
            for (uint32_t c = 0; c < C; c++) { const float gv = center[c] - nb[c]; res[c] += gv; idelta[c] += gv * gv; }
        }
        pg[d] = cur;
    }
    float out[C];
#pragma unroll
    for (uint32_t c = 0; c < C; c++) out[c] = w * res[c] * rsqrtf(idelta[c] + 1e-9f);
    red_add_entry<T, C>(gg + (size_t)index * C, out);
}

__global__ void k_level_scales(uint32_t L, float S, uint32_t H, float *__restrict__ out) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < L) out[l] = __fmaf_rn(exp2f((float)l * S), (float)H, -1.0f);
}

// ---- dispatch ----------------------------------------------------------------------------

constexpr uint32_t kBigBatch = 1u << 17;  // from here one thread walks all levels of its point

template <typename T, uint32_t D, uint32_t C>
int launch_forward(const float *inputs, const T *emb, const int *offsets, T *outputs, uint32_t B, uint32_t L, float S,
                   uint32_t H, T *dy_dx, uint32_t gridtype, bool ac, uint32_t interp, cudaStream_t st) {
    // all-levels form: 512-thread CTAs and streaming (evict-first) stores of the level-major outputs -- the best of the
    // measured launch variants (profiles/r1d_experiments.md, runs 27 and 37); the outputs are written once and never re-read
    if (B >= kBigBatch && dy_dx == nullptr)
        k_grid_forward<T, D, C, true, 2, true, 512><<<dim3(div_up(B, 512u), 1), 512, 0, st>>>(inputs, emb, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac, interp);
    else
        k_grid_forward<T, D, C, false><<<dim3(div_up(B, 256u), L), 256, 0, st>>>(inputs, emb, offsets, outputs, B, L, S, H, dy_dx, gridtype, ac, interp);
    return (int)cudaPeekAtLastError();
}

template <typename T, uint32_t D, uint32_t C>
int launch_backward(const T *grad, const float *inputs, const int *offsets, T *grad_emb, uint32_t B, uint32_t L, float S,
                    uint32_t H, uint32_t gridtype, bool ac, uint32_t interp, cudaStream_t st) {
    if (B >= kBigBatch)
        k_grid_backward<T, D, C, true><<<dim3(div_up(B, 256u), 1), 256, 0, st>>>(grad, inputs, offsets, grad_emb, B, L, S, H, gridtype, ac, interp);
    else
        k_grid_backward<T, D, C, false><<<dim3(div_up(B, 256u), L), 256, 0, st>>>(grad, inputs, offsets, grad_emb, B, L, S, H, gridtype, ac, interp);
    return (int)cudaPeekAtLastError();
}

#define S3D_DISPATCH_DC(D_, C_, CALL)                                             \
    switch ((D_) * 16 + (C_)) {                                                   \
        case 2 * 16 + 1: { constexpr uint32_t kD = 2, kC = 1; CALL; } break;      \
        case 2 * 16 + 2: { constexpr uint32_t kD = 2, kC = 2; CALL; } break;      \
        case 2 * 16 + 4: { constexpr uint32_t kD = 2, kC = 4; CALL; } break;      \
        case 2 * 16 + 8: { constexpr uint32_t kD = 2, kC = 8; CALL; } break;      \
        case 3 * 16 + 1: { constexpr uint32_t kD = 3, kC = 1; CALL; } break;      \
        case 3 * 16 + 2: { constexpr uint32_t kD = 3, kC = 2; CALL; } break;      \
        case 3 * 16 + 4: { constexpr uint32_t kD = 3, kC = 4; CALL; } break;      \
        case 3 * 16 + 8: { constexpr uint32_t kD = 3, kC = 8; CALL; } break;      \
        case 4 * 16 + 1: { constexpr uint32_t kD = 4, kC = 1; CALL; } break;      \
        case 4 * 16 + 2: { constexpr uint32_t kD = 4, kC = 2; CALL; } break;      \
        case 4 * 16 + 4: { constexpr uint32_t kD = 4, kC = 4; CALL; } break;      \
        case 4 * 16 + 8: { constexpr uint32_t kD = 4, kC = 8; CALL; } break;      \
        case 5 * 16 + 1: { constexpr uint32_t kD = 5, kC = 1; CALL; } break;      \
        case 5 * 16 + 2: { constexpr uint32_t kD = 5, kC = 2; CALL; } break;      \
        case 5 * 16 + 4: { constexpr uint32_t kD = 5, kC = 4; CALL; } break;      \
        case 5 * 16 + 8: { constexpr uint32_t kD = 5, kC = 8; CALL; } break;      \
        default: return S3D_EINVAL; /* reference: D in 2..5 (gridencoder.cu:388-396), C in {1,2,4,8} (:373-379) */ \
    }

}  // namespace

// the per-level scale exp2f(l*S)*H - 1 exactly as the kernels evaluate it (gridencoder.cu:138 of the reference);
// diagnostic entry used by the parity tests (see oracle/seal_oracle.c: orc_set_level_scales)
S3D_API int s3d_grid_level_scales(uint32_t L, float S, uint32_t H, float *scales, void *stream) {
    if (L == 0) return 0;
    k_level_scales<<<1, 64, 0, as_stream(stream)>>>(L, S, H, scales);
    S3D_RETURN_LAST();
}

// dtype: 0 = float32 table/outputs, 1 = float16 table/outputs (inputs are always float32)
S3D_API int s3d_grid_encode_forward(const float *inputs, const void *embeddings, const int *offsets, void *outputs,
                                    uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void *dy_dx,
                                    uint32_t gridtype, int align_corners, uint32_t interp, int dtype, void *stream) {
    if (B == 0) return 0;
    if (dtype != 0 && dtype != 1) return S3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    int rc = 0;
    if (dtype == 0) {
        S3D_DISPATCH_DC(D, C, (rc = launch_forward<float, kD, kC>(inputs, (const float *)embeddings, offsets, (float *)outputs, B, L, S, H, (float *)dy_dx, gridtype, align_corners != 0, interp, st)))
    } else {
        S3D_DISPATCH_DC(D, C, (rc = launch_forward<__half, kD, kC>(inputs, (const __half *)embeddings, offsets, (__half *)outputs, B, L, S, H, (__half *)dy_dx, gridtype, align_corners != 0, interp, st)))
    }
    return rc;
}

// grad_embeddings is accumulated into (the caller zero-fills it, grid.py:77); grad_inputs is overwritten.
S3D_API int s3d_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings, const int *offsets,
                                     void *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                     uint32_t H, const void *dy_dx, void *grad_inputs, uint32_t gridtype,
                                     int align_corners, uint32_t interp, int dtype, void *stream) {
    (void)embeddings;
    if (B == 0) return 0;
    if (dtype != 0 && dtype != 1) return S3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    int rc = 0;
    if (dtype == 0) {
        S3D_DISPATCH_DC(D, C, (rc = launch_backward<float, kD, kC>((const float *)grad, inputs, offsets, (float *)grad_embeddings, B, L, S, H, gridtype, align_corners != 0, interp, st)))
    } else {
        S3D_DISPATCH_DC(D, C, (rc = launch_backward<__half, kD, kC>((const __half *)grad, inputs, offsets, (__half *)grad_embeddings, B, L, S, H, gridtype, align_corners != 0, interp, st)))
    }
    if (rc) return rc;
    if (dy_dx && grad_inputs) {
        if (dtype == 0) k_grid_input_backward<float><<<div_up(B * D, 256u), 256, 0, st>>>((const float *)grad, (const float *)dy_dx, (float *)grad_inputs, B, D, C, L);
        else k_grid_input_backward<__half><<<div_up(B * D, 256u), 256, 0, st>>>((const __half *)grad, (const __half *)dy_dx, (__half *)grad_inputs, B, D, C, L);
    }
    S3D_RETURN_LAST();
}

S3D_API int s3d_grad_total_variation(const void *inputs, const void *embeddings, void *grad, const int *offsets,
                                     float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                     uint32_t gridtype, int align_corners, int dtype, void *stream) {
    if (B == 0) return 0;
    if (dtype != 0) return S3D_ENOTSUP;  // the reference wrapper runs it with autocast disabled (grid.py:164)
    cudaStream_t st = as_stream(stream);
    S3D_DISPATCH_DC(D, C, (k_grad_tv<float, kD, kC><<<dim3(div_up(B, 256u), L), 256, 0, st>>>((const float *)inputs, (const float *)embeddings, (float *)grad, offsets, weight, B, L, S, H, gridtype, align_corners != 0)))
    S3D_RETURN_LAST();
}
