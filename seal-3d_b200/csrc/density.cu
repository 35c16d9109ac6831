// Density-grid refresh (nerf/renderer.py:445-538 update_extra_state) as a chain of launches with NO host read in it:
//
//   pick cells      partial update (:487-499): H^3/4 uniformly random cells + H^3/4 cells drawn (with repetition) from the
//                   occupied set {density_grid > 0}.  The reference builds the occupied set with torch.nonzero (a device
//                   sync) -- here it is a count -> scan -> write stream compaction whose size stays on the device, and the
//                   draws index it there.
//   cells -> xyz    :469-479 / :501-509, cell centre of the cascade + uniform jitter of half a cell, one rounding per torch op
//   (density)       the caller's field (fused: k_ngp_encode + k_ngp_mlp_fwd, sigma only)
//   scatter         tmp_grid[cell] = sigma * density_scale (:483, :513).  A cell drawn twice gets "one of" its values in the
//                   reference (index_put without accumulate is order-dependent on the GPU); here the LARGEST, by an integer
//                   atomicMax on the float bits (sigma >= 0) -- deterministic, and one of the reference's possible outcomes.
//   EMA-max + mean  :521-524, with the mean reduced in a FIXED order (per-block partials, then one block): data-parallel
//                   replicas must derive bit-identical thresholds, which float atomics do not give.
//   packbits        :529-530 with density_thresh = min(mean_density, thresh) read from device memory.
//
// Random numbers: counter-based (PCG hash of (seed, index)), so every rank draws the same cells and jitter for the same
// seed with no broadcast.  The reference uses torch's global generator, unseeded: its stream cannot be reproduced, so parity
// is defined on an injected stream (oracle/: orc_update_extra_state takes the draws as arrays; tests regenerate this hash).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t pcg(uint32_t v) {
    v = v * 747796405u + 2891336453u;
    const uint32_t w = ((v >> ((v >> 28u) + 4u)) ^ v) * 277803737u;
    return (w >> 22u) ^ w;
}
__device__ __forceinline__ uint32_t draw(uint32_t seed, uint32_t counter) { return pcg(seed ^ pcg(counter)); }

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t compact_bits10(uint32_t x) {
    x &= 0x49249249u; x = (x | (x >> 2)) & 0xc30c30c3u; x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu; x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

constexpr uint32_t kBlk = 1024;

// ---- occupied-cell compaction -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlk) k_occ_count(const float *__restrict__ grid, uint32_t n, uint32_t *__restrict__ block_count) {
    const uint32_t i = blockIdx.x * kBlk + threadIdx.x;
    const int c = __syncthreads_count(i < n && grid[i] > 0.0f);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}
// exclusive scan of up to 4096 block counts in one block; prefix[nb] = total
__global__ void __launch_bounds__(kBlk) k_occ_scan(const uint32_t *__restrict__ block_count, uint32_t nb, uint32_t *__restrict__ prefix) {
    __shared__ uint32_t s[kBlk];
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t j = threadIdx.x * 4 + k; v[k] = j < nb ? block_count[j] : 0; sum += v[k]; }
    s[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < kBlk; d <<= 1) {
        const uint32_t t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = s[threadIdx.x] - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t j = threadIdx.x * 4 + k; if (j < nb) prefix[j] = run; run += v[k]; }
    if (threadIdx.x == kBlk - 1) prefix[nb] = s[kBlk - 1];
}
__global__ void __launch_bounds__(kBlk) k_occ_write(const float *__restrict__ grid, uint32_t n, const uint32_t *__restrict__ prefix, int *__restrict__ list) {
    __shared__ uint32_t warp_base[kBlk / 32];
    const uint32_t i = blockIdx.x * kBlk + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool occ = i < n && grid[i] > 0.0f;
    const uint32_t m = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) warp_base[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        uint32_t c = warp_base[lane], x = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, d); if (lane >= (uint32_t)d) x += t; }
        warp_base[lane] = x - c;
    }
    __syncthreads();
    if (occ) list[prefix[blockIdx.x] + warp_base[warp] + __popc(m & ((1u << lane) - 1u))] = (int)i;   // ascending, like torch.nonzero
}

// cells_out[0, n_uniform): morton3D of uniformly random coords; [n_uniform, n_uniform + n_occ): list[floor(u * Nz)].
// Nz = 0 (the reference raises in torch.randint(0, 0)): the occupied half repeats the uniform half.
__global__ void k_pick_cells(uint32_t H, uint32_t n_uniform, uint32_t n_occ, uint32_t seed, const int *__restrict__ list,
                             const uint32_t *__restrict__ n_list, int *__restrict__ cells_out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_uniform) {
        uint32_t c[3];
#pragma unroll
        for (uint32_t d = 0; d < 3; d++) c[d] = (uint32_t)(((uint64_t)draw(seed, j * 3u + d) * H) >> 32);
        cells_out[j] = (int)(expand_bits10(c[0]) | (expand_bits10(c[1]) << 1) | (expand_bits10(c[2]) << 2));
    } else if (j < n_uniform + n_occ) {
        const uint32_t nz = *n_list, k = j - n_uniform;
        if (nz == 0) {
            uint32_t c[3];
#pragma unroll
            for (uint32_t d = 0; d < 3; d++) c[d] = (uint32_t)(((uint64_t)draw(seed, (k % max(n_uniform, 1u)) * 3u + d) * H) >> 32);
            cells_out[j] = (int)(expand_bits10(c[0]) | (expand_bits10(c[1]) << 1) | (expand_bits10(c[2]) << 2));
        } else {
            cells_out[j] = list[(uint32_t)(((uint64_t)draw(seed ^ 0x9e3779b9u, k) * nz) >> 32)];
        }
    }
}

// ---- cell -> jittered position ------------------------------------------------------------------------------------
// nerf/renderer.py:470-479: xyzs = 2 * coords.float() / (H - 1) - 1; cas_xyzs = xyzs * (bound - hgs); cas_xyzs += (rand * 2 - 1) * hgs
// -- torch rounds after every op, so every op is an explicit _rn intrinsic (no FMA contraction).
__global__ void k_cells_to_xyz(const int *__restrict__ cells, uint32_t n, uint32_t H, float bound_cas, uint32_t seed, float *__restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t idx = cells ? (uint32_t)cells[i] : i;
    const uint32_t c[3] = {compact_bits10(idx), compact_bits10(idx >> 1), compact_bits10(idx >> 2)};
    const float hgs = __fdiv_rn(bound_cas, (float)H), span = __fsub_rn(bound_cas, hgs);
#pragma unroll
    for (uint32_t d = 0; d < 3; d++) {
        const float u = (float)(draw(seed ^ 0x85ebca6bu, i * 3u + d) >> 8) * (1.0f / 16777216.0f);  // [0,1), 24 bits like torch.rand
        const float base = __fmul_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)c[d]), (float)(H - 1)), 1.0f), span);
        xyz[(size_t)i * 3 + d] = __fadd_rn(base, __fmul_rn(__fsub_rn(__fmul_rn(u, 2.0f), 1.0f), hgs));
    }
}

__global__ void k_scatter_max(const int *__restrict__ cells, const float *__restrict__ sigma, uint32_t n, float scale, float *__restrict__ tmp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = __fmul_rn(sigma[i], scale);
    if (!cells) { tmp[i] = v; return; }
    // v >= 0 and tmp starts at -1: as signed integers the float bit patterns order like the floats (NaN sorts above +inf)
    atomicMax(reinterpret_cast<int *>(tmp) + (uint32_t)cells[i], __float_as_int(v));
}

// ---- EMA-max + deterministic mean ---------------------------------------------------------------------------------
constexpr uint32_t kEmaBlocks = 1024;
__global__ void __launch_bounds__(256) k_ema_partial(float *__restrict__ grid, const float *__restrict__ tmp, uint32_t n, float decay, double *__restrict__ partial) {
    __shared__ double s[256];
    double acc = 0.0;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += kEmaBlocks * 256) {
        float g = grid[i];
        const float t = tmp[i];
        if (g >= 0.0f && t >= 0.0f) { g = fmaxf(__fmul_rn(g, decay), t); grid[i] = g; }
        acc += (double)fmaxf(g, 0.0f);
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
// stats[0] = mean_density, stats[1] = min(mean_density, density_thresh)
__global__ void __launch_bounds__(kEmaBlocks) k_ema_final(const double *__restrict__ partial, uint32_t n, float density_thresh, float *__restrict__ stats) {
    __shared__ double s[kEmaBlocks];
    s[threadIdx.x] = partial[threadIdx.x];
    __syncthreads();
    for (uint32_t d = kEmaBlocks / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) s[threadIdx.x] += s[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mean = (float)(s[0] / (double)n);
        stats[0] = mean;
        stats[1] = fminf(mean, density_thresh);
    }
}

// packbits with the threshold in device memory (same packing as k_packbits in raymarching.cu)
__global__ void k_packbits_dev(const float *__restrict__ grid, uint32_t N, const float *__restrict__ thresh_p, uint8_t *__restrict__ bitfield) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t first = w * 4;
    if (first >= N) return;
    const float thresh = __ldg(thresh_p);
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

// mean_count = int(sum(step_counter[:total, 0]) / total)  (nerf/renderer.py:533-536), on the device
__global__ void k_mean_count(const int *__restrict__ step_counter, uint32_t total, int *__restrict__ out) {
    long long s = 0;
    for (uint32_t i = 0; i < total; i++) s += step_counter[i * 2];
    // Python: int(tensor_sum.item() / total) -- true division in double, truncation toward zero
    *out = total ? (int)((double)s / (double)total) : -1;
}

}  // namespace

S3D_API int s3d_density_pick_cells(const float *density_grid_cas, uint32_t H, uint32_t n_uniform, uint32_t n_occ, uint32_t seed,
                                   int *cells_out, uint32_t *n_occupied_out, void *stream) {
    const uint32_t n = H * H * H, nb = div_up(n, kBlk);
    if (H == 0 || H > 1024 || nb > 4 * kBlk) return S3D_EINVAL;
    if (n_uniform + n_occ == 0) return 0;
    cudaStream_t st = as_stream(stream);
    uint32_t *scratch = nullptr;   // [nb] counts | [nb + 1] prefix | [n] list
    cudaError_t e = scratch_alloc((void **)&scratch, ((size_t)2 * nb + 1 + n) * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    uint32_t *counts = scratch, *prefix = scratch + nb;
    int *list = (int *)(scratch + 2 * nb + 1);
    k_occ_count<<<nb, kBlk, 0, st>>>(density_grid_cas, n, counts);
    k_occ_scan<<<1, kBlk, 0, st>>>(counts, nb, prefix);
    k_occ_write<<<nb, kBlk, 0, st>>>(density_grid_cas, n, prefix, list);
    k_pick_cells<<<div_up(n_uniform + n_occ, 256u), 256, 0, st>>>(H, n_uniform, n_occ, seed, list, prefix + nb, cells_out);
    if (n_occupied_out) cudaMemcpyAsync(n_occupied_out, prefix + nb, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
    cudaFreeAsync(scratch, st);
    S3D_RETURN_LAST();
}

S3D_API int s3d_density_cells_to_xyz(const int *cell_morton, uint32_t n, uint32_t H, float bound_cas, uint32_t seed, float *xyz, void *stream) {
    if (n == 0) return 0;
    k_cells_to_xyz<<<div_up(n, 256u), 256, 0, as_stream(stream)>>>(cell_morton, n, H, bound_cas, seed, xyz);
    S3D_RETURN_LAST();
}

S3D_API int s3d_density_scatter(const int *cell_morton, const float *sigma, uint32_t n, float density_scale, float *tmp_grid, void *stream) {
    if (n == 0) return 0;
    k_scatter_max<<<div_up(n, 256u), 256, 0, as_stream(stream)>>>(cell_morton, sigma, n, density_scale, tmp_grid);
    S3D_RETURN_LAST();
}

S3D_API int s3d_density_grid_update(float *grid, const float *tmp_grid, uint32_t n, float decay, float density_thresh, float *stats_out, void *stream) {
    if (n == 0) return 0;
    cudaStream_t st = as_stream(stream);
    double *partial = nullptr;
    cudaError_t e = scratch_alloc((void **)&partial, kEmaBlocks * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    k_ema_partial<<<kEmaBlocks, 256, 0, st>>>(grid, tmp_grid, n, decay, partial);
    k_ema_final<<<1, kEmaBlocks, 0, st>>>(partial, n, density_thresh, stats_out);
    cudaFreeAsync(partial, st);
    S3D_RETURN_LAST();
}

S3D_API int s3d_packbits_dev_thresh(const float *grid, uint32_t N, const float *density_thresh_dev, uint8_t *bitfield, void *stream) {
    if (N == 0) return 0;
    k_packbits_dev<<<div_up(div_up(N, 4u), 256u), 256, 0, as_stream(stream)>>>(grid, N, density_thresh_dev, bitfield);
    S3D_RETURN_LAST();
}

S3D_API int s3d_mean_count(const int *step_counter, uint32_t total_step, int *mean_count_out, void *stream) {
    k_mean_count<<<1, 1, 0, as_stream(stream)>>>(step_counter, total_step, mean_count_out);
    S3D_RETURN_LAST();
}
