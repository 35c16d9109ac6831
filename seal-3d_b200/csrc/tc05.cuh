// tcgen05 / TMEM / mbarrier helpers (sm_100a inline PTX) shared by the fused-MLP kernels.
//
// Shared-memory operand tiles all use one layout: rows of 64 fp16 (128 B), SWIZZLE_128B, 8-row
// groups of 1024 B (tile base 1024-B aligned):
//     byte(r, c16) = (r / 8) * 1024 + (r % 8) * 128 + ((c16 ^ (r % 8)) * 16),   c16 = 16-byte chunk 0..7
// The same bytes serve as
//   * a K-major operand   (rows = M or N index, the 64 columns = K)          -- forward / dgrad A, forward B
//   * an MN-major operand (rows = K index, the 64 columns = M or N index)    -- dgrad B (= W^T), wgrad A and B
// which is what lets one activation tile feed the forward GEMM, the data-gradient GEMM and the
// weight-gradient GEMM without any transposition in shared memory.
#pragma once
#include "common.cuh"

namespace tc05 {

constexpr uint32_t kTileRowBytes = 128;
constexpr uint32_t kGroupBytes = 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of 16-byte chunk c16 of row r inside a SW128 tile
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c16) {
    return (r >> 3) * kGroupBytes + (r & 7) * kTileRowBytes + ((c16 ^ (r & 7)) << 4);
}

// ---- descriptors ---------------------------------------------------------------------------
// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | base_offset [49,52) | lbo_mode [52] | layout_type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major tile: one MMA consumes 16 K-elements = 32 B of every row -> advance the start address by 32 B
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, uint32_t kstep) {
    return make_desc_sw128(tile_addr + kstep * 32u, 16u, kGroupBytes);
}
// MN-major tile: one MMA consumes 16 K-rows = two 8-row groups -> advance by 2048 B; LBO = next 64-wide MN block
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, uint32_t kstep, uint32_t mn_block_stride_bytes) {
    return make_desc_sw128(tile_addr + kstep * 2u * kGroupBytes, mn_block_stride_bytes, kGroupBytes);
}

// InstrDescriptor for kind::f16, fp16 operands, fp32 accumulate:
// c_format=1 [4,6) | a_format=0 [7,10) | b_format=0 [10,13) | a_major [15] | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- MMA issue / completion ----------------------------------------------------------------
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// warp index / elected lane in the form ptxas recognises as warp-uniform (what CUTLASS calls canonical_warp_idx_sync /
// elect_one_sync): a role branch on `threadIdx.x >> 5` is divergent as far as the compiler can tell, and inside a divergent
// region every tcgen05 operand is moved to a uniform register through an ELECT + R2UR loop (~25 instructions per MMA).
__device__ __forceinline__ uint32_t warp_idx_sync() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Issue-predicated form for an issuer WARP: all 32 lanes run the surrounding code, which keeps the control flow warp-uniform
// and therefore the descriptors / addresses in uniform registers (UTCHMMA takes its operands from URs; under a divergent
// `if (lane == 0)` every operand costs a dependent R2UR, ~100 cycles per MMA measured); only the lane whose `issue` is true
// executes the tcgen05 instruction itself.
__device__ __forceinline__ void mma_f16_if(bool issue, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)issue)
        : "memory");
}
__device__ __forceinline__ void mma_commit_if(bool issue, uint32_t mbar_addr) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(mbar_addr), "r"((uint32_t)issue)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t mbar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar_addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_addr), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// bounded wait: a wrong descriptor must surface as a launch failure, not as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t mbar_addr, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 24); it++) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(mbar_addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// ---- TMEM ----------------------------------------------------------------------------------
// one full warp allocates; the address lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst_addr, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst_addr), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets row (lane_base + i), columns col..col+31.
// A warp may only address the 32 TMEM lanes of its own sub-partition (warp_id % 4).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    __syncwarp();  // .sync.aligned: the whole warp must arrive converged
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// ---- A operand in TMEM (tcgen05.mma "TS" form) -----------------------------------------------------------------------
// An fp16 A tile [128 rows x K] lives in TMEM as lane = row, 32-bit column j = {A[row][2j] (low half), A[row][2j+1]}
// (cute::UMMA::tmem_frg_1sm<half_t, half_t>: dense 16-bit values, K-major -- "A from TMEM can't be transposed").  One
// K = 16 MMA consumes 8 columns.  The epilogue thread that owns a row stores its packed activations with tcgen05.st, so the
// activation never touches shared memory: no STS, no fence.proxy.async, and the MMA reads only its B operand (2 KB per
// M128 N64 K16 instruction = 16 cycles of shared-memory bandwidth against 32 cycles of tensor time) instead of A + B
// (6 KB = 48 cycles), which is what made the SS-form MLP kernels shared-memory bound at 13 % tensor-pipe utilisation.
__device__ __forceinline__ void mma_f16_ts_if(bool issue, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)issue)
        : "memory");
}
// 32 lanes x 16 columns: thread i of the warp writes 16 packed half2 words into row (lane_base + i), columns col..col+15
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    __syncwarp();
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// plain (non-tcgen05) arrival of one thread on an mbarrier
__device__ __forceinline__ void mbar_arrive(uint32_t mbar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_addr) : "memory");
}
// ---- 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256) --------------------------------------------------------------
// The MLP kernels own one matrix ROW per thread (TMEM lane = row), so a warp-wide access touches 32 different lines and costs 32
// L1 wavefronts whatever its width: moving a row's contiguous bytes 32 at a time instead of 16 halves the wavefronts (the wide
// FFMLP epilogues were bound by exactly that, see profiles/r2_experiments.md).  The address must be 32-byte aligned.
__device__ __forceinline__ bool aligned32(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }
__device__ __forceinline__ void ldg256(const void *p, uint4 &a, uint4 &b) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, const uint4 &a, const uint4 &b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
                 "r"(b.z), "r"(b.w)
                 : "memory");
}
// two adjacent 16-byte pieces of a row: one 256-bit store when the address allows it
__device__ __forceinline__ void stg_pair(void *p, const uint4 &a, const uint4 &b) {
    if (aligned32(p)) stg256(p, a, b);
    else { reinterpret_cast<uint4 *>(p)[0] = a; reinterpret_cast<uint4 *>(p)[1] = b; }
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// pack 8 floats into one 16-byte chunk of fp16
__device__ __forceinline__ uint4 pack8(const float *v) {
    uint4 u;
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]),
            d = __floats2half2_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t *>(&a); u.y = *reinterpret_cast<uint32_t *>(&b);
    u.z = *reinterpret_cast<uint32_t *>(&c); u.w = *reinterpret_cast<uint32_t *>(&d);
    return u;
}
__device__ __forceinline__ void unpack8(const uint4 &u, float *v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
        v[q * 2] = f.x; v[q * 2 + 1] = f.y;
    }
}

}  // namespace tc05
