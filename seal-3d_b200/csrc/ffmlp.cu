// Fully fused bias-free MLP on tcgen05 tensor cores (sm_100a).
//
// Replaces ffmlp/src/ffmlp.cu of the reference (entry points ffmlp.h:8-14): fp16 inputs / weights,
// `num_layers` hidden activations (= num_layers + 1 matmuls), hidden width <= 64 (16 / 32 are zero
// padded to 64), output width 16, input width a multiple of 16 up to 256.
//
// The reference is a wmma m16n16k16 kernel with fp16 accumulation plus CUTLASS-2.8 Volta split-K
// GEMMs on side streams for the weight gradients (ffmlp.cu:332-518, :749-894).  Here:
//   * a CTA owns 128 batch rows; one elected thread issues tcgen05.mma (M=128, N=64|16, K=16 per
//     instruction) with operands in SWIZZLE_128B shared-memory tiles and the fp32 accumulator in
//     TMEM; thread i reads back accumulator row i (tcgen05.ld 32x32b), applies the activation and
//     writes the fp16 row straight into the next layer's operand tile -- activations never leave
//     the SM except for the forward_buffer / backward_buffer the API asks for;
//   * the backward kernel is persistent: data gradients chain through the same tiles with the
//     weight tile read MN-major (W^T without a transpose), and the weight gradients
//     dW_l = dA_l^T . A_{l-1} are tcgen05 MMAs too (both operands MN-major, K = the 128 batch rows)
//     accumulated in TMEM across all tiles of the CTA and flushed once -- no split-K, no side streams;
//   * everything runs on the caller's stream; allocate_splitk/free_splitk are kept as no-ops.
#include "tc05.cuh"

int s3d_ffmlp_wide_forward(const __half *inputs, const __half *weights, uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden,
                           uint32_t num_layers, uint32_t act, uint32_t out_act, __half *forward_buffer, __half *outputs, cudaStream_t st);
int s3d_ffmlp_wide_backward(const __half *grad, const __half *inputs, const __half *weights, const __half *forward_buffer, uint32_t B,
                            uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers, uint32_t act, int calc_grad_inputs,
                            __half *backward_buffer, __half *grad_inputs, __half *grad_weights, cudaStream_t st);

namespace {

using namespace tc05;

enum Act : uint32_t { kReLU = 0, kExp = 1, kSine = 2, kSigmoid = 3, kSquareplus = 4, kSoftplus = 5, kNone = 6 };

__device__ __forceinline__ float act_fwd(uint32_t a, float x) {
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
// derivative as a function of the stored forward activation y (ffmlp/src/utils.h:537-582)
__device__ __forceinline__ float act_bwd(uint32_t a, float y) {
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

constexpr uint32_t kRows = 128;        // batch rows per tile
constexpr uint32_t kW = 64;            // padded hidden width
constexpr uint32_t kTileBytes = kRows * 128;   // 16 KB: [128 x 64] fp16
constexpr uint32_t kWTileBytes = kW * 128;     // 8 KB:  [64 x 64] fp16
constexpr uint32_t kOTileBytes = 16 * 128;     // 2 KB:  [16 x 64] fp16

// copy a row-major fp16 matrix [rows x cols] (cols % 8 == 0, cols <= 64*blocks) from global memory into
// SW128 tiles of `tile_rows` rows (one tile per 64-column block), zero-filling up to pad_rows x 64.
__device__ void load_matrix_sw128(uint8_t *smem, const __half *__restrict__ g, uint32_t rows, uint32_t cols,
                                  uint32_t pad_rows, uint32_t blocks, uint32_t tile_bytes) {
    const uint32_t chunks = pad_rows * 8 * blocks;
    for (uint32_t i = threadIdx.x; i < chunks; i += blockDim.x) {
        const uint32_t blk = i / (pad_rows * 8), rem = i - blk * pad_rows * 8, r = rem >> 3, c16 = rem & 7;
        const uint32_t col = blk * 64 + c16 * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows && col < cols) v = __ldg(reinterpret_cast<const uint4 *>(g + (size_t)r * cols + col));
        *reinterpret_cast<uint4 *>(smem + blk * tile_bytes + sw128_off(r, c16)) = v;
    }
}

// thread `row` loads its batch row (cols halves, zero padded to 64*blocks) into the activation tile(s)
__device__ __forceinline__ void load_row_sw128(uint8_t *tile, const __half *__restrict__ g_row, uint32_t row, uint32_t cols,
                                               uint32_t blocks, bool in_range) {
    const bool wide = aligned32(g_row);      // 256-bit loads: a row-per-thread access costs 32 L1 wavefronts whatever its width
    for (uint32_t blk = 0; blk < blocks; blk++) {
#pragma unroll
        for (uint32_t c16 = 0; c16 < 8; c16 += 2) {
            const uint32_t col = blk * 64 + c16 * 8;
            uint4 v = make_uint4(0, 0, 0, 0), w = v;
            if (in_range && wide && col + 16 <= cols) ldg256(g_row + col, v, w);
            else {
                if (in_range && col < cols) v = __ldg(reinterpret_cast<const uint4 *>(g_row + col));
                if (in_range && col + 8 < cols) w = __ldg(reinterpret_cast<const uint4 *>(g_row + col + 8));
            }
            *reinterpret_cast<uint4 *>(tile + blk * kTileBytes + sw128_off(row, c16)) = v;
            *reinterpret_cast<uint4 *>(tile + blk * kTileBytes + sw128_off(row, c16 + 1)) = w;
        }
    }
}

struct FwdParams {
    const __half *inputs, *weights;
    __half *forward_buffer, *outputs;
    uint32_t B, in_dim, out_dim, hidden, num_layers, act, out_act;
};

// dynamic smem: [A tiles: in_blocks x 16K] [W0: in_blocks x 8K] [hidden W: (num_layers-1) x 8K] [Wout 2K]
__global__ void __launch_bounds__(128)
k_ffmlp_forward(const FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t in_blocks = (p.in_dim + 63) / 64;
    uint8_t *tA = smem;
    uint8_t *tW0 = tA + in_blocks * kTileBytes;
    uint8_t *tWh = tW0 + in_blocks * kWTileBytes;
    uint8_t *tWo = tWh + (p.num_layers - 1) * kWTileBytes;

    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();   // warp-uniform for the compiler (see tc05.cuh)
    const uint32_t mbar = smem_u32(&s_mbar);
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 64);
    if (tid == 0) mbar_init(mbar, 1);

    // weights -> shared (layer 0 may span several 64-column K blocks)
    const __half *w = p.weights;
    load_matrix_sw128(tW0, w, p.hidden, p.in_dim, kW, in_blocks, kWTileBytes);
    w += (size_t)p.hidden * p.in_dim;
    for (uint32_t l = 0; l + 1 < p.num_layers; l++, w += (size_t)p.hidden * p.hidden)
        load_matrix_sw128(tWh + l * kWTileBytes, w, p.hidden, p.hidden, kW, 1, kWTileBytes);
    load_matrix_sw128(tWo, w, p.out_dim, p.hidden, 16, 1, kOTileBytes);

    const uint32_t row = blockIdx.x * kRows + tid;
    const bool in_range = row < p.B;
    load_row_sw128(tA, p.inputs + (size_t)row * p.in_dim, tid, p.in_dim, in_blocks, in_range);

    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t t_row = tmem + ((warp * 32u) << 16);  // this warp's 32 TMEM lanes
    uint32_t parity = 0;

    const uint32_t n_mm = p.num_layers + 1;
    for (uint32_t l = 0; l < n_mm; l++) {
        const bool last = (l == n_mm - 1);
        if (warp == 0) {
            const bool lead = elect_one();
            const uint32_t idesc = make_idesc(128, last ? 16 : 64, false, false);
            if (l == 0) {
                const uint32_t ksteps = (p.in_dim + 15) / 16;
                for (uint32_t k = 0; k < ksteps; k++) {
                    const uint32_t blk = k >> 2, kk = k & 3;
                    mma_f16_if(lead, tmem, desc_kmajor(smem_u32(tA + blk * kTileBytes), kk), desc_kmajor(smem_u32(tW0 + blk * kWTileBytes), kk), idesc, k > 0);
                }
            } else {
                const uint8_t *wt = last ? tWo : (tWh + (l - 1) * kWTileBytes);
                const uint32_t ksteps = (p.hidden + 15) / 16;
                for (uint32_t k = 0; k < ksteps; k++)
                    mma_f16_if(lead, tmem, desc_kmajor(smem_u32(tA), k), desc_kmajor(smem_u32(wt), k), idesc, k > 0);
            }
            mma_commit_if(lead, mbar);
        }
        mbar_wait(mbar, parity);
        parity ^= 1;
        fence_after_sync();

        if (!last) {
            // hidden activation: TMEM row -> act -> fp16 -> operand tile (+ forward_buffer)
            __half *fb = p.forward_buffer ? p.forward_buffer + ((size_t)l * p.B + row) * p.hidden : nullptr;
#pragma unroll
            for (uint32_t half = 0; half < 2; half++) {
                float v[32];
                tmem_ld32(t_row + half * 32, v);
#pragma unroll
                for (int i = 0; i < 32; i++) v[i] = act_fwd(p.act, v[i]);
#pragma unroll
                for (uint32_t q = 0; q < 4; q += 2) {
                    const uint32_t c16 = half * 4 + q;
                    const uint4 u = pack8(v + q * 8), u2 = pack8(v + q * 8 + 8);
                    *reinterpret_cast<uint4 *>(tA + sw128_off(tid, c16)) = u;
                    *reinterpret_cast<uint4 *>(tA + sw128_off(tid, c16 + 1)) = u2;
                    if (fb && in_range) {
                        if (c16 * 8 + 16 <= p.hidden) stg_pair(fb + c16 * 8, u, u2);
                        else if (c16 * 8 < p.hidden) *reinterpret_cast<uint4 *>(fb + c16 * 8) = u;
                    }
                }
            }
            fence_async_smem();
        } else {
            float v[16];
            tmem_ld16(t_row, v);
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = act_fwd(p.out_act, v[i]);
            if (in_range) {
                __half *o = p.outputs + (size_t)row * p.out_dim;
                if (p.out_dim == 16) {
                    stg_pair(o, pack8(v), pack8(v + 8));
                } else {
                    for (uint32_t i = 0; i < p.out_dim; i++) o[i] = __float2half_rn(v[i]);
                }
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    if (warp == 0) tmem_dealloc(tmem, 64);
}

struct BwdParams {
    const __half *grad, *inputs, *weights, *forward_buffer;
    __half *backward_buffer, *grad_inputs;
    float *gw32;  // fp32 accumulation workspace, same flat layout as the weights
    uint32_t B, in_dim, out_dim, hidden, num_layers, act, n_tiles;
};

// TMEM columns: [0,64) data-gradient accumulator; then one 64-column weight-gradient accumulator per
// 64x64 block: W_out (1), hidden matrices (num_layers-1), W_0 (in_blocks).  <= 512 columns total.
__global__ void __launch_bounds__(128)
k_ffmlp_backward(const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ uint32_t s_tmem;
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t in_blocks = (p.in_dim + 63) / 64;
    const uint32_t n_hid = p.num_layers - 1;
    uint8_t *tD = smem;                       // dA tile (current layer's activation gradient), 16 KB
    uint8_t *tH = tD + kTileBytes;            // forward activation tiles h_k / inputs, in_blocks x 16 KB
    uint8_t *tW0 = tH + in_blocks * kTileBytes;
    uint8_t *tWh = tW0 + in_blocks * kWTileBytes;
    uint8_t *tWo = tWh + n_hid * kWTileBytes;

    const uint32_t tid = threadIdx.x, warp = warp_idx_sync();   // warp-uniform for the compiler (see tc05.cuh)
    const uint32_t mbar = smem_u32(&s_mbar);
    const uint32_t n_acc = 1 + n_hid + in_blocks;              // weight-gradient accumulators
    uint32_t ncols = 64 * (1 + n_acc), alloc = 32;
    while (alloc < ncols) alloc <<= 1;
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), alloc);
    if (tid == 0) mbar_init(mbar, 1);

    const __half *w = p.weights;
    load_matrix_sw128(tW0, w, p.hidden, p.in_dim, kW, in_blocks, kWTileBytes);
    w += (size_t)p.hidden * p.in_dim;
    for (uint32_t l = 0; l < n_hid; l++, w += (size_t)p.hidden * p.hidden)
        load_matrix_sw128(tWh + l * kWTileBytes, w, p.hidden, p.hidden, kW, 1, kWTileBytes);
    load_matrix_sw128(tWo, w, p.out_dim, p.hidden, 16, 1, kOTileBytes);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const uint32_t tmem = s_tmem;
    const uint32_t t_row = tmem + ((warp * 32u) << 16);
    const uint32_t acc_out = tmem + 64, acc_hid = acc_out + 64, acc_in = acc_hid + 64 * n_hid;
    uint32_t parity = 0;
    const uint32_t idesc_wg = make_idesc(64, 64, true, true);  // dW (64 x 64) += A^T(MN-major) . B(MN-major), K = 16 rows

    bool first_tile = true;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, first_tile = false) {
        const uint32_t row = tile * kRows + tid;
        const bool in_range = row < p.B;
        // stage 0: output gradient (16 wide, zero padded to 64) -> tD ; h_last -> tH
        {
#pragma unroll
            for (uint32_t c16 = 0; c16 < 8; c16++) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (in_range && c16 * 8 < p.out_dim) v = __ldg(reinterpret_cast<const uint4 *>(p.grad + (size_t)row * p.out_dim + c16 * 8));
                *reinterpret_cast<uint4 *>(tD + sw128_off(tid, c16)) = v;
            }
            load_row_sw128(tH, p.forward_buffer + ((size_t)(p.num_layers - 1) * p.B + row) * p.hidden, tid, p.hidden, 1, in_range);
        }
        fence_async_smem();
        fence_before_sync();
        __syncthreads();
        fence_after_sync();

        // layer loop: s = 0 handles W_out, s = 1..n_hid the hidden matrices (top down), s = n_hid+1 the input layer
        for (uint32_t s = 0; s <= n_hid + 1; s++) {
            const bool is_out = (s == 0), is_in = (s == n_hid + 1);
            if (warp == 0) {
                const bool lead = elect_one();
                // (1) weight gradient of this matrix: dW += dA^T . H     (K = 128 batch rows, 8 MMAs of 16 rows)
                const uint32_t nblk = is_in ? in_blocks : 1;
                for (uint32_t blk = 0; blk < nblk; blk++) {
                    const uint32_t acc = is_out ? acc_out : (is_in ? acc_in + 64 * blk : acc_hid + 64 * (n_hid - s));
                    for (uint32_t k = 0; k < kRows / 16; k++)
                        mma_f16_if(lead, acc, desc_mnmajor(smem_u32(tD), k, kTileBytes), desc_mnmajor(smem_u32(tH + blk * kTileBytes), k, kTileBytes),
                                idesc_wg, !(first_tile && k == 0));
                }
                // (2) data gradient through this matrix: D = dA . W     (W tile read MN-major = W^T)
                if (is_out) {
                    mma_f16_if(lead, tmem, desc_kmajor(smem_u32(tD), 0), desc_mnmajor(smem_u32(tWo), 0, kOTileBytes), make_idesc(128, 64, false, true), false);
                } else if (!is_in) {
                    const uint8_t *wt = tWh + (n_hid - s) * kWTileBytes;
                    for (uint32_t k = 0; k < (p.hidden + 15) / 16; k++)
                        mma_f16_if(lead, tmem, desc_kmajor(smem_u32(tD), k), desc_mnmajor(smem_u32(wt), k, kWTileBytes), make_idesc(128, 64, false, true), k > 0);
                } else if (p.grad_inputs) {
                    // grad_inputs block by block reuses the accumulator; handled below one block at a time
                    for (uint32_t k = 0; k < (p.hidden + 15) / 16; k++)
                        mma_f16_if(lead, tmem, desc_kmajor(smem_u32(tD), k), desc_mnmajor(smem_u32(tW0), k, kWTileBytes), make_idesc(128, 64, false, true), k > 0);
                }
                mma_commit_if(lead, mbar);
            }
            mbar_wait(mbar, parity);
            parity ^= 1;
            fence_after_sync();

            if (!is_in) {
                // dA_next = D (.) act'(h) with h = the forward activation this gradient flows into (held in tH, own row)
                const uint32_t bb_idx = s;  // backward_buffer[0] = gradient w.r.t. h_last, ...
                __half *bb = p.backward_buffer ? p.backward_buffer + ((size_t)bb_idx * p.B + row) * p.hidden : nullptr;
#pragma unroll
                for (uint32_t half = 0; half < 2; half++) {
                    float v[32];
                    tmem_ld32(t_row + half * 32, v);
#pragma unroll
                    for (uint32_t q = 0; q < 4; q += 2) {
                        const uint32_t c16 = half * 4 + q;
                        uint4 u[2];
#pragma unroll
                        for (uint32_t e = 0; e < 2; e++) {
                            float h[8];
                            unpack8(*reinterpret_cast<const uint4 *>(tH + sw128_off(tid, c16 + e)), h);
#pragma unroll
                            for (int i = 0; i < 8; i++) v[(q + e) * 8 + i] *= act_bwd(p.act, h[i]);
                            u[e] = pack8(v + (q + e) * 8);
                            *reinterpret_cast<uint4 *>(tD + sw128_off(tid, c16 + e)) = u[e];
                        }
                        if (bb && in_range) {
                            if (c16 * 8 + 16 <= p.hidden) stg_pair(bb + c16 * 8, u[0], u[1]);
                            else if (c16 * 8 < p.hidden) *reinterpret_cast<uint4 *>(bb + c16 * 8) = u[0];
                        }
                    }
                }
                // next forward activation: h_{k-1}, or the network inputs for the last step
                __syncthreads();  // every thread is done reading its tH row before it is overwritten (rows are private, but keep the phases aligned)
                if (s < n_hid) load_row_sw128(tH, p.forward_buffer + ((size_t)(n_hid - 1 - s) * p.B + row) * p.hidden, tid, p.hidden, 1, in_range);
                else load_row_sw128(tH, p.inputs + (size_t)row * p.in_dim, tid, p.in_dim, in_blocks, in_range);
                fence_async_smem();
            } else if (p.grad_inputs) {
                // grad_inputs = dA_0 . W_0, 64 columns per pass
                for (uint32_t blk = 0; blk < in_blocks; blk++) {
                    if (blk > 0) {
                        fence_before_sync();
                        __syncthreads();
                        fence_after_sync();
                        if (warp == 0) {
                            const bool lead = elect_one();
                            for (uint32_t k = 0; k < (p.hidden + 15) / 16; k++)
                                mma_f16_if(lead, tmem, desc_kmajor(smem_u32(tD), k), desc_mnmajor(smem_u32(tW0 + blk * kWTileBytes), k, kWTileBytes), make_idesc(128, 64, false, true), k > 0);
                            mma_commit_if(lead, mbar);
                        }
                        mbar_wait(mbar, parity);
                        parity ^= 1;
                        fence_after_sync();
                    }
#pragma unroll
                    for (uint32_t half = 0; half < 2; half++) {
                        float v[32];
                        tmem_ld32(t_row + half * 32, v);
#pragma unroll
                        for (uint32_t q = 0; q < 4; q += 2) {
                            const uint32_t col = blk * 64 + (half * 4 + q) * 8;
                            __half *o = p.grad_inputs + (size_t)row * p.in_dim + col;
                            if (in_range && col + 16 <= p.in_dim) stg_pair(o, pack8(v + q * 8), pack8(v + q * 8 + 8));
                            else if (in_range && col < p.in_dim) *reinterpret_cast<uint4 *>(o) = pack8(v + q * 8);
                        }
                    }
                }
            }
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
        }
    }

    // flush the weight-gradient accumulators: M = 64 rows live in lanes 0..15 of each 32-lane sub-partition
    // (row i -> lane 32*(i/16) + i%16), so warp w / lane t < 16 holds row 16*w + t.
    {
        const uint32_t lane = tid & 31;
        const uint32_t r = warp * 16 + lane;  // output-feature row of the weight matrix
        const size_t off_hid = (size_t)p.hidden * p.in_dim, off_out = off_hid + (size_t)p.hidden * p.hidden * n_hid;
        for (uint32_t a = 0; a < n_acc; a++) {
#pragma unroll
            for (uint32_t half = 0; half < 2; half++) {
                float v[32];
                tmem_ld32(t_row + 64 * (1 + a) + half * 32, v);  // warp-collective: all lanes execute
                float *dst;
                uint32_t ld, rows, cols, col0;
                if (a == 0) { dst = p.gw32 + off_out; ld = p.hidden; rows = p.out_dim; cols = p.hidden; col0 = 0; }
                else if (a <= n_hid) { dst = p.gw32 + off_hid + (size_t)(a - 1) * p.hidden * p.hidden; ld = p.hidden; rows = p.hidden; cols = p.hidden; col0 = 0; }
                else { dst = p.gw32; ld = p.in_dim; rows = p.hidden; cols = p.in_dim; col0 = (a - 1 - n_hid) * 64; }
                if (lane < 16 && r < rows) {
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const uint32_t c = col0 + half * 32 + i;
                        if (c < cols) atomicAdd(dst + (size_t)r * ld + c, v[i]);
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, alloc);
}

__global__ void k_f32_to_f16(const float *__restrict__ src, __half *__restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}

int check_dims(uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers) {
    if (hidden != 16 && hidden != 32 && hidden != 64) return S3D_ENOTSUP;  // 128 / 256 and output > 16: ffmlp_wide.cu
    if (in_dim == 0 || in_dim % 16 != 0 || in_dim > 256) return S3D_EINVAL;
    if (out_dim == 0 || out_dim > 16) return S3D_ENOTSUP;
    if (num_layers < 2 || num_layers > 5) return S3D_EINVAL;
    return 0;
}
// hidden 128 / 256 (ffmlp/src/ffmlp.cu:653-658) and output_dim > 16 (:661-670) take the wide kernels; so does a narrow network
// that is deeper than the TMEM-resident weight-gradient accumulators of the narrow backward allow
bool is_wide(uint32_t hidden, uint32_t out_dim, uint32_t num_layers) { return hidden > 64 || out_dim > 16 || num_layers > 5; }

int run_forward(const __half *inputs, const __half *weights, uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden,
                uint32_t num_layers, uint32_t act, uint32_t out_act, __half *forward_buffer, __half *outputs, cudaStream_t st) {
    if (B == 0) return 0;
    if (is_wide(hidden, out_dim, num_layers))
        return s3d_ffmlp_wide_forward(inputs, weights, B, in_dim, out_dim, hidden, num_layers, act, out_act, forward_buffer, outputs, st);
    if (int rc = check_dims(in_dim, out_dim, hidden, num_layers)) return rc;
    const uint32_t in_blocks = (in_dim + 63) / 64;
    const size_t smem = 1024 + (size_t)in_blocks * (kTileBytes + kWTileBytes) + (size_t)(num_layers - 1) * kWTileBytes + kOTileBytes;
    cudaError_t e = cudaFuncSetAttribute(k_ffmlp_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    FwdParams p{inputs, weights, forward_buffer, outputs, B, in_dim, out_dim, hidden, num_layers, act, out_act};
    k_ffmlp_forward<<<div_up(B, kRows), 128, smem, st>>>(p);
    return (int)cudaPeekAtLastError();
}

}  // namespace

// ffmlp.h:8  (forward_buffer [num_layers, B, hidden], outputs [B, output_dim])
S3D_API int s3d_ffmlp_forward(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                              uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                              void *forward_buffer, void *outputs, void *stream) {
    return run_forward((const __half *)inputs, (const __half *)weights, B, input_dim, output_dim, hidden_dim, num_layers,
                       activation, output_activation, (__half *)forward_buffer, (__half *)outputs, as_stream(stream));
}

// ffmlp.h:9  (inference_buffer is only written by the reference when output_dim > 16; unused here)
S3D_API int s3d_ffmlp_inference(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                                void *inference_buffer, void *outputs, void *stream) {
    (void)inference_buffer;
    return run_forward((const __half *)inputs, (const __half *)weights, B, input_dim, output_dim, hidden_dim, num_layers,
                       activation, output_activation, nullptr, (__half *)outputs, as_stream(stream));
}

// ffmlp.h:11.  grad [B, output_dim]; backward_buffer [num_layers, B, hidden]; grad_inputs [B, input_dim] (written
// only when calc_grad_inputs); grad_weights flat fp16 (overwritten).  The output activation is ignored like
// in the reference (ffmlp.cu:781).
S3D_API int s3d_ffmlp_backward(const void *grad, const void *inputs, const void *weights, const void *forward_buffer,
                               uint32_t B, uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                               uint32_t activation, uint32_t output_activation, int calc_grad_inputs, void *backward_buffer,
                               void *grad_inputs, void *grad_weights, void *stream) {
    (void)output_activation;
    if (B == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const uint32_t in_blocks = (input_dim + 63) / 64;
    if (is_wide(hidden_dim, output_dim, num_layers) || 64 * (1 + 1 + (num_layers - 1) + in_blocks) > 512)
        return s3d_ffmlp_wide_backward((const __half *)grad, (const __half *)inputs, (const __half *)weights, (const __half *)forward_buffer, B, input_dim,
                                       output_dim, hidden_dim, num_layers, activation, calc_grad_inputs, (__half *)backward_buffer,
                                       (__half *)grad_inputs, (__half *)grad_weights, st);
    if (int rc = check_dims(input_dim, output_dim, hidden_dim, num_layers)) return rc;
    // sine needs the pre-activations, which the API does not store: the reference's backward returns without writing
    // (ffmlp/src/utils.h:552-556, "assert(false)" commented out); here the call is refused
    if (activation == kSine) return S3D_ENOTSUP;
    const size_t nW = (size_t)hidden_dim * input_dim + (size_t)hidden_dim * hidden_dim * (num_layers - 1) + (size_t)output_dim * hidden_dim;
    float *gw32 = nullptr;
    cudaError_t e = scratch_alloc((void **)&gw32, nW * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    cudaMemsetAsync(gw32, 0, nW * sizeof(float), st);
    const size_t smem = 1024 + kTileBytes + (size_t)in_blocks * (kTileBytes + kWTileBytes) + (size_t)(num_layers - 1) * kWTileBytes + kOTileBytes;
    e = cudaFuncSetAttribute(k_ffmlp_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaFreeAsync(gw32, st); return (int)e; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t n_tiles = div_up(B, kRows);
    BwdParams p{(const __half *)grad, (const __half *)inputs, (const __half *)weights, (const __half *)forward_buffer,
                (__half *)backward_buffer, calc_grad_inputs ? (__half *)grad_inputs : nullptr, gw32,
                B, input_dim, output_dim, hidden_dim, num_layers, activation, n_tiles};
    k_ffmlp_backward<<<min(n_tiles, (uint32_t)sms), 128, smem, st>>>(p);
    k_f32_to_f16<<<(unsigned)div_up(nW, (size_t)256), 256, 0, st>>>(gw32, (__half *)grad_weights, nW);
    e = cudaPeekAtLastError();
    cudaFreeAsync(gw32, st);
    return (int)e;
}

// ffmlp.h:13-14: the reference creates side streams + events for its split-K GEMMs.  Nothing to allocate here.
S3D_API int s3d_allocate_splitk(size_t size) { (void)size; return 0; }
S3D_API int s3d_free_splitk(void) { return 0; }
