// Fully fused MLP, wide configurations: hidden width 128 / 256 and / or output width > 16 (sm_100a, tcgen05).
//
// ffmlp/src/ffmlp.cu:653-658 dispatches hidden 16..256; for output_dim > 16 the reference runs the last layer as a separate
// CUTLASS GEMM (:661-670) and the weight gradients as split-K CUTLASS GEMMs on side streams (:814-894).  The narrow kernels of
// ffmlp.cu (hidden <= 64: every matrix resident in shared memory, weight gradients accumulated in TMEM) do not scale to these
// shapes -- a 256 x 256 weight-gradient accumulator alone is all of TMEM -- so the wide path is three kernels:
//
//   k_wide_forward    persistent; thread i owns batch row i of a 128-row tile.  The activation lives in TENSOR MEMORY as the A
//                     operand (tcgen05.mma TS form, tc05.cuh): inputs go global -> registers -> tcgen05.st, every layer's
//                     epilogue reads the fp32 accumulator row, applies the activation, packs to fp16 and stores it back as
//                     the next A operand (and to forward_buffer, which the API asks for).  Shared memory holds only weights:
//                     all layers when they fit (hidden 128; hidden 256 with one hidden matrix), otherwise one layer at a time.
//   k_wide_backward   the data-gradient chain the same way: dA in TMEM, weights read MN-major (= W^T without a transpose),
//                     act'(h) taken from forward_buffer, backward_buffer / grad_inputs written as the API asks.
//   k_wide_wgrad      dW = dA^T . H as tcgen05 MMAs with both operands MN-major (K = batch rows) straight from the row-major
//                     forward / backward buffers: one CTA per (128-row block of dW, batch slice), fp32 accumulation in TMEM
//                     over the slice, one atomic flush -- the split-K GEMM of the reference without its side streams.
#include "tc05.cuh"

namespace {

using namespace tc05;

enum Act : uint32_t { kReLU = 0, kExp = 1, kSine = 2, kSigmoid = 3, kSquareplus = 4, kSoftplus = 5, kNone = 6 };

__device__ __forceinline__ float wact_fwd(uint32_t a, float x) {
    switch (a) {
        case kReLU: return fmaxf(x, 0.0f);
        case kExp: return __expf(x);
        case kSine: return __sinf(x);
        case kSigmoid: return 1.0f / (1.0f + __expf(-x));
        case kSquareplus: return 0.5f * (x + sqrtf(x * x + 4.0f));
        case kSoftplus: return __logf(__expf(x) + 1.0f);
        default: return x;
    }
}
__device__ __forceinline__ float wact_bwd(uint32_t a, float y) {   // as a function of the stored activation (ffmlp/src/utils.h:537-582)
    switch (a) {
        case kReLU: return y > 0.0f ? 1.0f : 0.0f;
        case kExp: return y;
        case kSigmoid: return y * (1.0f - y);
        case kSquareplus: { const float y2 = y * y; return y2 / (y2 + 1.0f); }
        case kSoftplus: return 1.0f - __expf(-y);
        case kNone: return 1.0f;
        default: return 0.0f;
    }
}

constexpr uint32_t kRows = 128;

struct Dims {
    uint32_t B, in_dim, out_dim, hidden, num_layers, act, out_act;
    uint32_t in_pad, out_pad;     // multiples of 16
    uint32_t n_tiles;
    int resident;                 // all weight matrices stay in shared memory
};

// matrix m of the network: 0 = input layer, 1 .. num_layers-1 = hidden, num_layers = output layer
__device__ __forceinline__ void matrix_shape(const Dims &d, uint32_t m, uint32_t &rows, uint32_t &cols, uint32_t &pad_rows, size_t &goff) {
    if (d.num_layers == 0) { rows = d.out_dim; cols = d.in_dim; pad_rows = d.out_pad; goff = 0; }   // a single bias-free Linear (s3d_linear_*)
    else if (m == 0) { rows = d.hidden; cols = d.in_dim; pad_rows = d.hidden; goff = 0; }
    else if (m < d.num_layers) { rows = d.hidden; cols = d.hidden; pad_rows = d.hidden; goff = (size_t)d.hidden * d.in_dim + (size_t)(m - 1) * d.hidden * d.hidden; }
    else { rows = d.out_dim; cols = d.hidden; pad_rows = d.out_pad; goff = (size_t)d.hidden * d.in_dim + (size_t)(d.num_layers - 1) * d.hidden * d.hidden; }
}
__host__ __device__ inline uint32_t matrix_bytes(uint32_t pad_rows, uint32_t cols) { return ((cols + 63) / 64) * pad_rows * 128; }

// row-major fp16 [rows x cols] -> K-major SW128 tiles, one [pad_rows x 64] tile per 64-column block, zero padded
__device__ void load_w(uint8_t *smem, const __half *__restrict__ g, uint32_t rows, uint32_t cols, uint32_t pad_rows) {
    const uint32_t blocks = (cols + 63) / 64, tile_bytes = pad_rows * 128, chunks = pad_rows * 8 * blocks;
    for (uint32_t i = threadIdx.x; i < chunks; i += blockDim.x) {
        const uint32_t blk = i / (pad_rows * 8), rem = i - blk * pad_rows * 8, r = rem >> 3, c16 = rem & 7;
        const uint32_t col = blk * 64 + c16 * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows && col < cols) {
            if (col + 8 <= cols && (cols & 7u) == 0) v = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)r * cols + col));
            else {
                __half h[8];
                for (uint32_t j = 0; j < 8; j++) h[j] = (col + j < cols) ? g[(size_t)r * cols + col + j] : __float2half_rn(0.0f);
                v = *reinterpret_cast<const uint4 *>(h);
            }
        }
        *reinterpret_cast<uint4 *>(smem + blk * tile_bytes + sw128_off(r, c16)) = v;
    }
}

// byte offset of matrix m in the shared weight area (resident mode) -- streaming mode always uses offset 0
__device__ __forceinline__ uint32_t matrix_smem_off(const Dims &d, uint32_t m) {
    uint32_t off = 0;
    for (uint32_t j = 0; j < m; j++) {
        uint32_t r, c, p; size_t g;
        matrix_shape(d, j, r, c, p, g);
        off += matrix_bytes(p, c);
    }
    return off;
}

__device__ __forceinline__ void sync_all() {
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
}

// thread `row` stores `n_half` fp16 values of a global row (zero padded to a multiple of 32) into TMEM as an A operand.
// The loads of four 32-value chunks (16 x 16 bytes) are issued before the first tcgen05.st: one exposed global-memory
// latency per 128 values instead of one per chunk (ncu: the kernel's warps sat on the long scoreboard at 12 % occupancy).
// `half_sel` = 0 / 1: the thread's warpgroup takes the 32-value chunks with an even / odd chunk index
// G = chunks per staging group (4 * G sixteen-byte loads in flight per thread).  row_load / row_store are the two halves of one
// group, so that a caller can issue the loads of the NEXT tile before it waits for the current tile's last MMA.
template <uint32_t G>
__device__ __forceinline__ void row_load(uint4 (&v)[4 * G], const __half *__restrict__ g_row, uint32_t n_half, bool in_range, uint32_t half_sel, uint32_t i0) {
    const bool vec = (n_half & 7u) == 0;
    const bool wide = (n_half & 15u) == 0 && aligned32(g_row);     // rows are multiples of 32 bytes: 256-bit loads
#pragma unroll
    for (uint32_t q = 0; q < 4 * G; q += 2) {
        v[q] = make_uint4(0, 0, 0, 0);
        v[q + 1] = make_uint4(0, 0, 0, 0);
        const uint32_t col = 32 * (2 * (i0 + q / 4) + half_sel) + (q & 3u) * 8;
        if (in_range && wide && col + 16 <= n_half) { ldg256(g_row + col, v[q], v[q + 1]); continue; }
#pragma unroll
        for (uint32_t u = 0; u < 2; u++) {
            const uint32_t cu = col + 8 * u;
            if (in_range && cu < n_half) {
                if (vec && cu + 8 <= n_half) v[q + u] = __ldg(reinterpret_cast<const uint4 *>(g_row + cu));
                else {
                    __half h[8];
                    for (uint32_t j = 0; j < 8; j++) h[j] = (cu + j < n_half) ? g_row[cu + j] : __float2half_rn(0.0f);
                    v[q + u] = *reinterpret_cast<const uint4 *>(h);
                }
            }
        }
    }
}
template <uint32_t G>
__device__ __forceinline__ void row_store(uint32_t t_a, const uint4 (&v)[4 * G], uint32_t n_half, uint32_t half_sel, uint32_t i0) {
#pragma unroll
    for (uint32_t ch = 0; ch < G; ch++) {
        const uint32_t c0 = 32 * (2 * (i0 + ch) + half_sel);
        if (c0 < n_half) {
            uint32_t pk[16];
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) { pk[4 * q] = v[ch * 4 + q].x; pk[4 * q + 1] = v[ch * 4 + q].y; pk[4 * q + 2] = v[ch * 4 + q].z; pk[4 * q + 3] = v[ch * 4 + q].w; }
            tmem_st16(t_a + c0 / 2, pk);
        }
    }
}
// groups [i_begin, ...) of a row, loaded and stored in place
template <uint32_t G>
__device__ __forceinline__ void row_to_tmem(uint32_t t_a, const __half *__restrict__ g_row, uint32_t n_half, bool in_range, uint32_t half_sel, uint32_t i_begin = 0) {
    for (uint32_t i0 = i_begin; 32 * (2 * i0 + half_sel) < n_half; i0 += G) {
        uint4 v[4 * G];
        row_load<G>(v, g_row, n_half, in_range, half_sel, i0);
        row_store<G>(t_a, v, n_half, half_sel, i0);
    }
}
struct FwdP {
    const __half *inputs, *weights;
    __half *forward_buffer, *outputs;
    Dims d;
    uint32_t tmem_cols, a_col;    // accumulator at column 0, A operand at a_col
};

// 256 threads: thread (r, hf) owns batch row r = tid & 127 of the tile and the 32-value chunks with chunk index % 2 == hf of
// every row it touches (input staging, accumulator read-back, activation store) -- two warpgroups per chain halve the serial
// epilogue of a layer and double the loads / stores in flight (TMEM allows only two 208-column chains per SM at hidden 128)
__global__ void __launch_bounds__(256, 2)
k_wide_forward(const FwdP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const Dims &d = p.d;
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();
    const uint32_t mbar = smem_u32(&s_mbar);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
    if (tid == 0) mbar_init(mbar, 1);
    const uint32_t n_mat = d.num_layers + 1;
    if (d.resident) {
        for (uint32_t m = 0; m < n_mat; m++) {
            uint32_t r, c, pr; size_t g;
            matrix_shape(d, m, r, c, pr, g);
            load_w(smem + matrix_smem_off(d, m), p.weights + g, r, c, pr);
        }
        fence_async_smem();
    }
    sync_all();
    const uint32_t r = tid & 127u, hf = tid >> 7;
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t t_acc = t_row, t_a = t_row + p.a_col;
    uint32_t parity = 0;
    // the first 192 input values of a row (group 0: three chunks per thread) are loaded one tile ahead: issued before the wait for the previous tile's
    // last layer, so their latency runs under that MMA and epilogue instead of in front of this tile's first MMA
    uint4 nx[12];
    row_load<3>(nx, p.inputs + (size_t)(blockIdx.x * kRows + r) * d.in_dim, d.in_pad, blockIdx.x * kRows + r < d.B, hf, 0);
    for (uint32_t tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x) {
        const uint32_t row = tile * kRows + r;
        const bool in_range = row < d.B;
        row_store<3>(t_a, nx, d.in_pad, hf, 0);
        row_to_tmem<3>(t_a, p.inputs + (size_t)row * d.in_dim, d.in_pad, in_range, hf, 3);
        tmem_st_wait();
        for (uint32_t m = 0; m < n_mat; m++) {
            uint32_t r, c, pr; size_t g;
            matrix_shape(d, m, r, c, pr, g);
            const uint32_t w_off = d.resident ? matrix_smem_off(d, m) : 0;
            if (!d.resident) {
                sync_all();                       // the previous layer's MMAs are done with the buffer (waited below), all threads past it
                load_w(smem, p.weights + g, r, c, pr);
                fence_async_smem();
            }
            sync_all();                           // A stores + weights visible to the issuing thread
            const bool last = (m == n_mat - 1);
            const uint32_t N = last ? d.out_pad : d.hidden, K = (m == 0) ? d.in_pad : d.hidden;   // num_layers = 0: m = 0 is also the last
            if (warp == 0) {
                const bool lead = elect_one();
                const uint32_t idesc = make_idesc(128, N, false, false);
                const uint32_t wbase = smem_u32(smem + w_off), tile_bytes = pr * 128;
                for (uint32_t k = 0; k < K / 16; k++)
                    mma_f16_ts_if(lead, tmem, tmem + p.a_col + 8 * k, desc_kmajor(wbase + (k >> 2) * tile_bytes, k & 3), idesc, k > 0);
                mma_commit_if(lead, mbar);
            }
            if (last) {
                const uint32_t nrow = row + gridDim.x * kRows;      // (r is shadowed by the matrix shape here)
                row_load<3>(nx, p.inputs + (size_t)nrow * d.in_dim, d.in_pad, tile + gridDim.x < d.n_tiles && nrow < d.B, hf, 0);
            }
            mbar_wait(mbar, parity);
            parity ^= 1;
            fence_after_sync();
            if (!last) {
                __half *fb = (p.forward_buffer && in_range) ? p.forward_buffer + ((size_t)m * d.B + row) * d.hidden : nullptr;
                for (uint32_t c0 = hf * 32; c0 < d.hidden; c0 += 64) {
                    float v[32];
                    tmem_ld32(t_acc + c0, v);
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) pk[i] = pack_half2(wact_fwd(d.act, v[2 * i]), wact_fwd(d.act, v[2 * i + 1]));
                    tmem_st16(t_a + c0 / 2, pk);
                    if (fb) {
#pragma unroll
                        for (uint32_t q = 0; q < 4; q += 2) {
                            const uint4 lo = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]), hi = make_uint4(pk[4 * q + 4], pk[4 * q + 5], pk[4 * q + 6], pk[4 * q + 7]);
                            if (c0 + q * 8 + 16 <= d.hidden) stg_pair(fb + c0 + q * 8, lo, hi);
                            else if (c0 + q * 8 < d.hidden) *reinterpret_cast<uint4 *>(fb + c0 + q * 8) = lo;
                        }
                    }
                }
                tmem_st_wait();
            } else {
                for (uint32_t c0 = hf * 16; c0 < d.out_pad; c0 += 32) {
                    float v[16];
                    tmem_ld16(t_acc + c0, v);
                    if (in_range) {
                        __half *o = p.outputs + (size_t)row * d.out_dim + c0;
                        if (c0 + 16 <= d.out_dim && (d.out_dim & 7u) == 0) {
                            float w[16];
#pragma unroll
                            for (int i = 0; i < 16; i++) w[i] = wact_fwd(d.out_act, v[i]);
                            stg_pair(o, pack8(w), pack8(w + 8));
                        } else {
                            for (uint32_t i = 0; i < 16 && c0 + i < d.out_dim; i++) o[i] = __float2half_rn(wact_fwd(d.out_act, v[i]));
                        }
                    }
                }
            }
        }
        sync_all();   // accumulator reads of this tile are done before the next tile's first MMA overwrites it
    }
    sync_all();
    if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

#ifdef S3D_WTRACE
__device__ long long g_wtrace[4096];
// tile index within the CTA (it), step, event -> clock64 of thread 0 / thread 128 (who = 0 / 1) of CTA 0
#define S3D_WT(who_tid, it, step, ev) do { if (blockIdx.x == 0 && threadIdx.x == (who_tid) && (it) >= 2 && (it) < 6) g_wtrace[((((it) - 2) * 2 + ((who_tid) ? 1 : 0)) * 4 + (step)) * 8 + (ev)] = clock64(); } while (0)
#else
#define S3D_WT(who_tid, it, step, ev) do { } while (0)
#endif

struct BwdP {
    const __half *grad, *weights, *forward_buffer;
    __half *backward_buffer, *grad_inputs;
    Dims d;
    uint32_t tmem_cols, a_col;
};

// data gradients, top down.  step s = 0: through W_out (K = out_pad, N = hidden); s = 1 .. n_hid: through hidden matrix
// num_layers - s; s = n_hid + 1: through W_0 (N = in_dim), only for grad_inputs.  backward_buffer[s] = dL/d(pre-activation of
// hidden activation num_layers-1-s) as in the narrow kernel (ffmlp.cu).
__global__ void __launch_bounds__(256, 2)
k_wide_backward(const BwdP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const Dims &d = p.d;
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();
    const uint32_t mbar = smem_u32(&s_mbar);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
    if (tid == 0) mbar_init(mbar, 1);
    const uint32_t n_mat = d.num_layers + 1;
    if (d.resident) {
        for (uint32_t m = 0; m < n_mat; m++) {
            uint32_t r, c, pr; size_t g;
            matrix_shape(d, m, r, c, pr, g);
            load_w(smem + matrix_smem_off(d, m), p.weights + g, r, c, pr);
        }
        fence_async_smem();
    }
    sync_all();
    const uint32_t r = tid & 127u, hf = tid >> 7;
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t t_acc = t_row, t_a = t_row + p.a_col;
    uint32_t parity = 0;
    const uint32_t n_steps = d.num_layers + (p.grad_inputs ? 1u : 0u);
    // Loads that do not depend on the running MMA are issued BEFORE its completion wait (the kernel was bound by exposed
    // global-load latency: 2 CTAs x 8 warps per SM): the first 128 gradient values of the next tile's row during the last step,
    // the activation row a step's ReLU mask needs while that step's MMA runs.
    uint4 ng[8];
    row_load<2>(ng, p.grad + (size_t)(blockIdx.x * kRows + r) * d.out_dim, d.out_dim, blockIdx.x * kRows + r < d.B, hf, 0);
    uint32_t wt_it = 0;
    for (uint32_t tile = blockIdx.x; tile < d.n_tiles; tile += gridDim.x, wt_it++) {
        const uint32_t row = tile * kRows + r;
        const bool in_range = row < d.B;
        S3D_WT(0, wt_it, 3, 0);
        row_store<2>(t_a, ng, d.out_dim, hf, 0);
        row_to_tmem<2>(t_a, p.grad + (size_t)row * d.out_dim, d.out_dim, in_range, hf, 2);
        tmem_st_wait();
        S3D_WT(0, wt_it, 3, 1);
        for (uint32_t s = 0; s < n_steps; s++) {
            const uint32_t m = d.num_layers - s;           // matrix the gradient flows through
            uint32_t r, c, pr; size_t g;
            matrix_shape(d, m, r, c, pr, g);
            const uint32_t w_off = d.resident ? matrix_smem_off(d, m) : 0;
            if (!d.resident) {
                sync_all();
                load_w(smem, p.weights + g, r, c, pr);
                fence_async_smem();
            }
            S3D_WT(0, wt_it, s, 0); S3D_WT(128, wt_it, s, 0);
            sync_all();
            S3D_WT(0, wt_it, s, 1); S3D_WT(128, wt_it, s, 1);
            const uint32_t K = pr;                          // rows of the matrix = width of the incoming gradient (padded)
            const uint32_t N = (m == 0) ? d.in_pad : d.hidden;
            if (warp == 0) {
                const bool lead = elect_one();
                const uint32_t idesc = make_idesc(128, N, false, true);   // B read MN-major: B[n][k] = W[k][n]
                const uint32_t wbase = smem_u32(smem + w_off), tile_bytes = pr * 128;
                for (uint32_t k = 0; k < K / 16; k++)
                    mma_f16_ts_if(lead, tmem, tmem + p.a_col + 8 * k, desc_mnmajor(wbase, k, tile_bytes), idesc, k > 0);
                mma_commit_if(lead, mbar);
            }
            S3D_WT(0, wt_it, s, 2);
            // h = hidden activation m-1 (forward_buffer[m-1]); this thread's chunks, two at a time
            const __half *h = m > 0 ? p.forward_buffer + ((size_t)(m - 1) * d.B + row) * d.hidden : nullptr;
            uint4 hu[8];
            if (m > 0) row_load<2>(hu, h, d.hidden, in_range, hf, 0);
            if (s + 1 == n_steps) {
                const uint32_t nrow = row + gridDim.x * kRows;      // (r is shadowed by the matrix shape here)
                row_load<2>(ng, p.grad + (size_t)nrow * d.out_dim, d.out_dim, tile + gridDim.x < d.n_tiles && nrow < d.B, hf, 0);
            }
            S3D_WT(0, wt_it, s, 3); S3D_WT(128, wt_it, s, 3);
            mbar_wait(mbar, parity);
            parity ^= 1;
            fence_after_sync();
            S3D_WT(0, wt_it, s, 4); S3D_WT(128, wt_it, s, 4);
            if (m > 0) {
                // dA = D (.) act'(h)
                __half *bb = (p.backward_buffer && in_range) ? p.backward_buffer + ((size_t)s * d.B + row) * d.hidden : nullptr;
                for (uint32_t i0 = 0; 32 * (2 * i0 + hf) < d.hidden; i0 += 2) {
                    if (i0 > 0) row_load<2>(hu, h, d.hidden, in_range, hf, i0);
#pragma unroll
                    for (uint32_t ch = 0; ch < 2; ch++) {
                        const uint32_t c0 = 32 * (2 * (i0 + ch) + hf);
                        if (c0 < d.hidden) {
                            float v[32];
                            tmem_ld32(t_acc + c0, v);
                            uint32_t pk[16];
#pragma unroll
                            for (uint32_t q = 0; q < 4; q++) {
                                float hv[8];
                                unpack8(hu[ch * 4 + q], hv);
#pragma unroll
                                for (int i = 0; i < 4; i++)
                                    pk[4 * q + i] = pack_half2(v[q * 8 + 2 * i] * wact_bwd(d.act, hv[2 * i]), v[q * 8 + 2 * i + 1] * wact_bwd(d.act, hv[2 * i + 1]));
                            }
                            tmem_st16(t_a + c0 / 2, pk);
                            if (bb) {
#pragma unroll
                                for (uint32_t q = 0; q < 4; q += 2) {
                                    const uint4 lo = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]), hi = make_uint4(pk[4 * q + 4], pk[4 * q + 5], pk[4 * q + 6], pk[4 * q + 7]);
                                    if (c0 + q * 8 + 16 <= d.hidden) stg_pair(bb + c0 + q * 8, lo, hi);
                                    else if (c0 + q * 8 < d.hidden) *reinterpret_cast<uint4 *>(bb + c0 + q * 8) = lo;
                                }
                            }
                        }
                    }
                }
                tmem_st_wait();
                S3D_WT(0, wt_it, s, 5); S3D_WT(128, wt_it, s, 5);
            } else {
                for (uint32_t c0 = hf * 16; c0 < d.in_pad; c0 += 32) {
                    float v[16];
                    tmem_ld16(t_acc + c0, v);
                    if (in_range && c0 < d.in_dim) {
                        __half *o = p.grad_inputs + (size_t)row * d.in_dim + c0;
                        stg_pair(o, pack8(v), pack8(v + 8));
                    }
                }
            }
        }
        sync_all();
    }
    sync_all();
    if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

// dW [R x C] += dA^T . H over a slice of the batch.  dA [B, ldA] (R valid columns), H [B, ldH] (C valid columns), row-major
// fp16 in global memory.  blockIdx.x = 128-row block of dW, blockIdx.y = batch slice.
struct WgP {
    const __half *dA, *H;
    float *gw;                    // [R x C] fp32, accumulated into
    uint32_t B, R, C, ldA, ldH, n_tiles, tmem_cols;
};

// batch rows [row0, row0 + 128) x columns [col0, col0 + 64 * blocks) of a row-major matrix -> SW128 tiles (rows = batch = the
// K index of an MN-major operand), zero padded
__device__ void load_rows_sw128(uint8_t *smem, const __half *__restrict__ g, uint32_t ld, uint32_t n_rows, uint32_t n_cols, uint32_t row0,
                                uint32_t col0, uint32_t blocks) {
    const bool vec = (ld & 7u) == 0;
    const uint32_t total = blocks * kRows * 8;
    for (uint32_t i0 = threadIdx.x; i0 < total; i0 += blockDim.x * 8) {
        uint4 v[8];
        uint32_t dst[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; u++) {     // eight 16-byte loads in flight per thread before the first shared-memory store
            const uint32_t i = i0 + u * blockDim.x;
            v[u] = make_uint4(0, 0, 0, 0);
            dst[u] = 0xffffffffu;
            if (i < total) {
                const uint32_t blk = i / (kRows * 8), rem = i - blk * kRows * 8, r = rem >> 3, c16 = rem & 7;
                const uint32_t row = row0 + r, col = col0 + blk * 64 + c16 * 8;
                dst[u] = blk * (kRows * 128) + sw128_off(r, c16);
                if (row < n_rows && col < n_cols) {
                    if (vec && col + 8 <= n_cols) v[u] = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)row * ld + col));
                    else {
                        __half h[8];
                        for (uint32_t j = 0; j < 8; j++) h[j] = (col + j < n_cols) ? g[(size_t)row * ld + col + j] : __float2half_rn(0.0f);
                        v[u] = *reinterpret_cast<const uint4 *>(h);
                    }
                }
            }
        }
#pragma unroll
        for (uint32_t u = 0; u < 8; u++)
            if (dst[u] != 0xffffffffu) *reinterpret_cast<uint4 *>(smem + dst[u]) = v[u];
    }
}

// 256 threads: twice the loads in flight while a tile pair is staged (the kernel is bound by the latency of its global loads,
// ncu: long scoreboard 23 cycles per issue at 12 % occupancy); warps 0-3 / 4-7 flush the low / high half of the accumulator columns
__global__ void __launch_bounds__(256)
k_wide_wgrad(const WgP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();
    const uint32_t mbar = smem_u32(&s_mbar);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
    if (tid == 0) mbar_init(mbar, 1);
    sync_all();
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t r0 = blockIdx.x * 128, c_blocks = (p.C + 63) / 64, Npad = c_blocks * 64;
    uint8_t *tA = smem, *tH = smem + 2 * kRows * 128;
    uint32_t parity = 0;
    bool first = true;
    for (uint32_t tile = blockIdx.y; tile < p.n_tiles; tile += gridDim.y) {
        load_rows_sw128(tA, p.dA, p.ldA, p.B, p.R, tile * kRows, r0, 2);
        load_rows_sw128(tH, p.H, p.ldH, p.B, p.C, tile * kRows, 0, c_blocks);
        fence_async_smem();
        sync_all();
        if (warp == 0) {
            const bool lead = elect_one();
            const uint32_t idesc = make_idesc(128, Npad, true, true);
            for (uint32_t k = 0; k < kRows / 16; k++)
                mma_f16_if(lead, tmem, desc_mnmajor(smem_u32(tA), k, kRows * 128), desc_mnmajor(smem_u32(tH), k, kRows * 128), idesc, !(first && k == 0));
            mma_commit_if(lead, mbar);
        }
        first = false;
        mbar_wait(mbar, parity);     // the tiles may be overwritten
        parity ^= 1;
        fence_after_sync();
    }
    if (!first) {
        const uint32_t rr = r0 + (tid & 127u);     // M = 128: accumulator row i lives in lane i
        const uint32_t half = tid >> 7;            // column half handled by this warpgroup
        for (uint32_t c0 = half * 32; c0 < Npad; c0 += 64) {
            float v[32];
            tmem_ld32(t_row + c0, v);
            if (rr < p.R) {
#pragma unroll
                for (int i = 0; i < 32; i++)
                    if (c0 + i < p.C) atomicAdd(p.gw + (size_t)rr * p.C + c0 + i, v[i]);
            }
        }
    }
    sync_all();
    if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

// ---- the same split-K weight gradient as a cp.async pipeline -----------------------------------------------------------------
// k_wide_wgrad stages a tile pair through registers and waits for it before every MMA group: with 2-3 CTAs per SM the global
// loads were not deep enough to cover their latency (3 GEMMs of the 160-128-128-16 head: 0.59 ms against 0.29 ms of HBM time).
// Here one CTA per SM keeps S stages of (dA tile, H tile) in shared memory; 16-byte cp.async copies land directly in the
// swizzled operand layout (zero-filled outside the matrix), S - 1 tiles are in flight while the tensor core consumes one.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void stage_rows_async(uint32_t smem_addr, const __half *__restrict__ g, uint32_t ld, uint32_t n_rows, uint32_t n_cols,
                                                 uint32_t row0, uint32_t col0, uint32_t blocks) {
    const uint32_t total = blocks * kRows * 8;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        const uint32_t blk = i / (kRows * 8), rem = i - blk * kRows * 8, r = rem >> 3, c16 = rem & 7;
        const uint32_t row = row0 + r, col = col0 + blk * 64 + c16 * 8;
        const bool ok = row < n_rows && col < n_cols;      // n_cols is a multiple of 8 here: a chunk is inside or outside
        cp_async16(smem_addr + blk * (kRows * 128) + sw128_off(r, c16), ok ? (const void *)(g + (size_t)row * ld + col) : (const void *)g, ok ? 16u : 0u);
    }
}

template <int S>
__global__ void __launch_bounds__(256, 1)
k_wide_wgrad_pipe(const WgP p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();
    const uint32_t mbar = smem_u32(&s_mbar);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), p.tmem_cols);
    if (tid == 0) mbar_init(mbar, 1);
    sync_all();
    const uint32_t tmem = s_tmem, t_row = tmem + (((warp & 3u) * 32u) << 16);
    const uint32_t r0 = blockIdx.x * 128, c_blocks = (p.C + 63) / 64, Npad = c_blocks * 64;
    const uint32_t stage_bytes = (2 + c_blocks) * kRows * 128, base = smem_u32(smem);
    const uint32_t n_my = blockIdx.y < p.n_tiles ? (p.n_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
    auto stage = [&](uint32_t j) {     // tile j of this CTA -> stage j % S
        const uint32_t a = base + (j % S) * stage_bytes, row0 = (blockIdx.y + j * gridDim.y) * kRows;
        stage_rows_async(a, p.dA, p.ldA, p.B, p.R, row0, r0, 2);
        stage_rows_async(a + 2 * kRows * 128, p.H, p.ldH, p.B, p.C, row0, 0, c_blocks);
    };
    for (uint32_t j = 0; j + 1 < S; j++) {
        if (j < n_my) stage(j);
        cp_async_commit();
    }
    uint32_t parity = 0;
    for (uint32_t j = 0; j < n_my; j++) {
        if (j > 0) {                    // the MMAs of tile j - 1 are done: its stage is the one refilled below
            mbar_wait(mbar, parity);
            parity ^= 1;
            fence_after_sync();
        }
        if (j + S - 1 < n_my) stage(j + S - 1);
        cp_async_commit();
        cp_async_wait<S - 1>();         // this thread's copies of tile j have landed
        fence_async_smem();
        sync_all();                     // ... and everybody's
        if (warp == 0) {
            const bool lead = elect_one();
            const uint32_t idesc = make_idesc(128, Npad, true, true);
            const uint32_t a = base + (j % S) * stage_bytes;
            for (uint32_t k = 0; k < kRows / 16; k++)
                mma_f16_if(lead, tmem, desc_mnmajor(a, k, kRows * 128), desc_mnmajor(a + 2 * kRows * 128, k, kRows * 128), idesc, !(j == 0 && k == 0));
            mma_commit_if(lead, mbar);
        }
    }
    if (n_my > 0) {
        mbar_wait(mbar, parity);
        fence_after_sync();
        const uint32_t rr = r0 + (tid & 127u);     // M = 128: accumulator row i lives in lane i
        const uint32_t half = tid >> 7;            // column half handled by this warpgroup
        for (uint32_t c0 = half * 32; c0 < Npad; c0 += 64) {
            float v[32];
            tmem_ld32(t_row + c0, v);
            if (rr < p.R) {
#pragma unroll
                for (int i = 0; i < 32; i++)
                    if (c0 + i < p.C) atomicAdd(p.gw + (size_t)rr * p.C + c0 + i, v[i]);
            }
        }
    }
    sync_all();
    if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

__global__ void k_wide_f32_to_f16(const float *__restrict__ src, __half *__restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}

uint32_t pow2_cols(uint32_t need) {
    uint32_t a = 32;
    while (a < need) a <<= 1;
    return a;
}

int make_dims(Dims &d, uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers, uint32_t act, uint32_t out_act, size_t &w_bytes) {
    const bool linear = num_layers == 0;       // s3d_linear_*: one matrix, no hidden layer
    if (!linear && hidden != 16 && hidden != 32 && hidden != 64 && hidden != 128 && hidden != 256) return S3D_ENOTSUP;   // ffmlp.cu:658
    if (in_dim == 0 || in_dim % 16 != 0 || in_dim > 256) return S3D_EINVAL;
    if (out_dim == 0 || out_dim > 256) return S3D_EINVAL;
    if (!linear && (num_layers < 2 || num_layers > 8)) return S3D_EINVAL;
    if (linear) hidden = 0;
    d.B = B; d.in_dim = in_dim; d.out_dim = out_dim; d.hidden = hidden; d.num_layers = num_layers; d.act = act; d.out_act = out_act;
    d.in_pad = in_dim; d.out_pad = (out_dim + 15) / 16 * 16;
    d.n_tiles = div_up(B, kRows);
    const size_t total = linear ? matrix_bytes(d.out_pad, in_dim)
                                : matrix_bytes(hidden, in_dim) + (size_t)(num_layers - 1) * matrix_bytes(hidden, hidden) + matrix_bytes(d.out_pad, hidden);
    const size_t largest = linear ? total : max(max((size_t)matrix_bytes(hidden, in_dim), (size_t)matrix_bytes(hidden, hidden)), (size_t)matrix_bytes(d.out_pad, hidden));
    d.resident = total <= 200 * 1024;
    w_bytes = d.resident ? total : largest;
    return 0;
}

int sm_count_w() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace

#ifdef S3D_WTRACE
S3D_API int s3d_debug_wtrace(long long *host_out, int n) { return (int)cudaMemcpyFromSymbol(host_out, g_wtrace, (size_t)n * sizeof(long long)); }
#endif

// called by the s3d_ffmlp_* entry points of ffmlp.cu for hidden > 64 or output_dim > 16
int s3d_ffmlp_wide_forward(const __half *inputs, const __half *weights, uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden,
                           uint32_t num_layers, uint32_t act, uint32_t out_act, __half *forward_buffer, __half *outputs, cudaStream_t st) {
    if (B == 0) return 0;
    FwdP p;
    size_t w_bytes = 0;
    if (int rc = make_dims(p.d, B, in_dim, out_dim, hidden, num_layers, act, out_act, w_bytes)) return rc;
    p.inputs = inputs; p.weights = weights; p.forward_buffer = forward_buffer; p.outputs = outputs;
    hidden = p.d.hidden;
    const uint32_t acc_cols = max(hidden, p.d.out_pad), a_cols = (max(max(p.d.in_pad, hidden), 32u) + 31) / 32 * 16;
    p.a_col = acc_cols;
    p.tmem_cols = pow2_cols(acc_cols + a_cols);
    if (p.tmem_cols > 512) return S3D_ENOTSUP;
    const size_t smem = 1024 + w_bytes;
    cudaError_t e = cudaFuncSetAttribute(k_wide_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const uint32_t per_sm = max(1u, min(512u / p.tmem_cols, (uint32_t)((220 * 1024) / smem)));
    k_wide_forward<<<min(p.d.n_tiles, (uint32_t)sm_count_w() * per_sm), 256, smem, st>>>(p);
    return (int)cudaPeekAtLastError();
}

int s3d_ffmlp_wide_backward(const __half *grad, const __half *inputs, const __half *weights, const __half *forward_buffer, uint32_t B,
                            uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers, uint32_t act, int calc_grad_inputs,
                            __half *backward_buffer, __half *grad_inputs, __half *grad_weights, cudaStream_t st) {
    if (B == 0) return 0;
    if (act == kSine) return S3D_ENOTSUP;     // needs pre-activations the API does not store (ffmlp/src/utils.h:552-556: the reference returns garbage)
    if (num_layers > 0 && (!backward_buffer || !forward_buffer)) return S3D_EINVAL;
    BwdP p;
    size_t w_bytes = 0;
    if (int rc = make_dims(p.d, B, in_dim, out_dim, hidden, num_layers, act, 6, w_bytes)) return rc;
    p.grad = grad; p.weights = weights; p.forward_buffer = forward_buffer; p.backward_buffer = backward_buffer;
    p.grad_inputs = calc_grad_inputs ? grad_inputs : nullptr;
    hidden = p.d.hidden;
    const uint32_t acc_cols = max(hidden, p.d.in_pad), a_cols = (max(max(p.d.out_pad, hidden), 32u) + 31) / 32 * 16;
    p.a_col = acc_cols;
    p.tmem_cols = pow2_cols(acc_cols + a_cols);
    if (p.tmem_cols > 512) return S3D_ENOTSUP;
    const size_t smem = 1024 + w_bytes;
    cudaError_t e = cudaFuncSetAttribute(k_wide_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const uint32_t per_sm = max(1u, min(512u / p.tmem_cols, (uint32_t)((220 * 1024) / smem)));
    const int sms = sm_count_w();
    k_wide_backward<<<min(p.d.n_tiles, (uint32_t)sms * per_sm), 256, smem, st>>>(p);
    e = cudaPeekAtLastError();
    if (e != cudaSuccess) return (int)e;
    // weight gradients: one split-K GEMM per matrix over the buffers the data pass just wrote
    const size_t nW = num_layers == 0 ? (size_t)out_dim * in_dim
                                      : (size_t)hidden * in_dim + (size_t)hidden * hidden * (num_layers - 1) + (size_t)out_dim * hidden;
    float *gw32 = nullptr;
    e = scratch_alloc((void **)&gw32, nW * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    cudaMemsetAsync(gw32, 0, nW * sizeof(float), st);
    const uint32_t n_hid = num_layers > 0 ? num_layers - 1 : 0;
    for (uint32_t m = 0; m <= num_layers; m++) {
        WgP w;
        w.B = B; w.n_tiles = p.d.n_tiles;
        if (m == num_layers) { w.dA = grad; w.ldA = out_dim; w.R = out_dim; }
        else { w.dA = backward_buffer + (size_t)(n_hid - m) * B * hidden; w.ldA = hidden; w.R = hidden; }   // backward_buffer[s], s = num_layers - 1 - m
        if (m == 0) { w.H = inputs; w.ldH = in_dim; w.C = in_dim; w.gw = gw32; }
        else {
            w.H = forward_buffer + (size_t)(m - 1) * B * hidden; w.ldH = hidden; w.C = hidden;
            w.gw = gw32 + (size_t)hidden * in_dim + (size_t)(m - 1) * hidden * hidden;
        }
        const uint32_t c_blocks = (w.C + 63) / 64;
        w.tmem_cols = pow2_cols(c_blocks * 64);
        static const bool no_pipe = [] { const char *e = getenv("S3D_WGRAD_PIPE"); return e && e[0] == '0'; }();
        if (!no_pipe && !((w.ldA | w.ldH | w.R | w.C) & 7u)) {
            // cp.async pipeline: S stages of (2 + c_blocks) 16 KB tiles, one CTA per SM
            const size_t stage_bytes = (size_t)(2 + c_blocks) * kRows * 128;
            const int S = (int)min((size_t)4, (size_t)(200 * 1024) / stage_bytes);
            if (S >= 2) {
                const size_t smem_p = 1024 + S * stage_bytes;
                const uint32_t m_blocks = div_up(w.R, 128u);
                const dim3 grid(m_blocks, max(1u, min(w.n_tiles, (uint32_t)sms / m_blocks)));
                auto launch = [&](auto kern) {
                    cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
                    if (ee == cudaSuccess) kern<<<grid, 256, smem_p, st>>>(w);
                    return ee;
                };
                e = S == 2 ? launch(k_wide_wgrad_pipe<2>) : (S == 3 ? launch(k_wide_wgrad_pipe<3>) : launch(k_wide_wgrad_pipe<4>));
                if (e != cudaSuccess) break;
                continue;
            }
        }
        const size_t smem_w = 1024 + (size_t)(2 + c_blocks) * kRows * 128;
        e = cudaFuncSetAttribute(k_wide_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
        if (e != cudaSuccess) break;
        const uint32_t m_blocks = div_up(w.R, 128u);
        const uint32_t slices = max(1u, min(w.n_tiles, (uint32_t)(2 * sms) / m_blocks));
        k_wide_wgrad<<<dim3(m_blocks, slices), 256, smem_w, st>>>(w);
    }
    if (e == cudaSuccess) {
        k_wide_f32_to_f16<<<(unsigned)div_up(nW, (size_t)256), 256, 0, st>>>(gw32, grad_weights, nW);
        e = cudaPeekAtLastError();
    }
    cudaFreeAsync(gw32, st);
    return (int)e;
}

// A single bias-free Linear on the same kernels (num_layers = 0): y [B,out] = x [B,in] . W^T, W [out,in] row-major, all fp16,
// fp32 accumulation.  in a multiple of 16 (<= 256), out <= 256.  Used for TensoRF's basis_mat (tensoRF/network.py:42, :155).
S3D_API int s3d_linear_forward(const void *x, const void *w, uint32_t B, uint32_t in_dim, uint32_t out_dim, void *y, void *stream) {
    return s3d_ffmlp_wide_forward((const __half *)x, (const __half *)w, B, in_dim, out_dim, 0, 0, kNone, kNone, nullptr, (__half *)y, as_stream(stream));
}
// grad_x [B,in] (NULL: not wanted), grad_w [out,in] (overwritten)
S3D_API int s3d_linear_backward(const void *grad_y, const void *x, const void *w, uint32_t B, uint32_t in_dim, uint32_t out_dim, void *grad_x,
                                void *grad_w, void *stream) {
    return s3d_ffmlp_wide_backward((const __half *)grad_y, (const __half *)x, (const __half *)w, nullptr, B, in_dim, out_dim, 0, 0, kNone,
                                   grad_x != nullptr, nullptr, (__half *)grad_x, (__half *)grad_w, as_stream(stream));
}
