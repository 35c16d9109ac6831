// micro-benchmark: cost of global reductions by width.  N_ENT random entries of a 98 MB table per warp instruction,
// RED.128 (v4.f32) vs RED.64 (v2.f32) vs RED.32 vs RED.64 of packed halfs (v2.f16x2).  Development tool (scripts/r2).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__global__ void k(float *tab, uint32_t n_entries, uint32_t per_thread) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = 0; i < per_thread; i++) {
        const uint32_t e = hash(t * 977u + i * 0x9e3779b9u) % n_entries;
        float *p = tab + (size_t)e * 4;
        if (MODE == 0) atomicAdd(reinterpret_cast<float4 *>(p), make_float4(1.f, 2.f, 3.f, 4.f));
        if (MODE == 1) atomicAdd(reinterpret_cast<float2 *>(p), make_float2(1.f, 2.f));
        if (MODE == 2) atomicAdd(p, 1.f);
        if (MODE == 3) { atomicAdd(reinterpret_cast<float2 *>(p), make_float2(1.f, 2.f)); atomicAdd(reinterpret_cast<float2 *>(p) + 1, make_float2(3.f, 4.f)); }
        if (MODE == 4) {
            const uint32_t a = 0x3c003c00u, b = 0x3c003c00u;
            asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
        }
        if (MODE == 6) {   // the two x-corners of a hashed cell: adjacent 16-byte entries of one 32-byte sector, two instructions
            float *q = tab + (size_t)(e & ~1u) * 4;
            atomicAdd(reinterpret_cast<float4 *>(q), make_float4(1.f, 2.f, 3.f, 4.f));
            atomicAdd(reinterpret_cast<float4 *>(q) + 1, make_float4(1.f, 2.f, 3.f, 4.f));
        }
        if (MODE == 7) {   // four adjacent entries (64 bytes, one half line), four instructions
            float *q = tab + (size_t)(e & ~3u) * 4;
            for (int j = 0; j < 4; j++) atomicAdd(reinterpret_cast<float4 *>(q) + j, make_float4(1.f, 2.f, 3.f, 4.f));
        }
        if (MODE == 8) {   // the same entry for the whole warp (what a run of equal cells would do without the in-warp merge)
            const uint32_t w = hash((t >> 5) * 977u + i * 0x9e3779b9u) % n_entries;
            atomicAdd(reinterpret_cast<float4 *>(tab + (size_t)w * 4), make_float4(1.f, 2.f, 3.f, 4.f));
        }
        if (MODE == 5) {
            const uint32_t a = 0x3f803f80u, b = 0x3f803f80u;
            asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
        }
    }
}
int main() {
    const uint32_t n_entries = 6098120;   // the benchmark's table: 6.1 M entries x 16 B = 98 MB
    float *tab;
    cudaMalloc(&tab, (size_t)n_entries * 16);
    cudaMemset(tab, 0, (size_t)n_entries * 16);
    const uint32_t threads = 148 * 8 * 256 * 4, per = 64;
    const char *names[] = {"RED.128 v4.f32", "RED.64 v2.f32", "RED.32 f32", "2 x RED.64 v2.f32 (one entry)", "RED.64 v2.f16x2", "RED.64 v2.bf16x2",
                           "2 x RED.128, one 32 B sector", "4 x RED.128, one 64 B half line", "RED.128, whole warp -> one entry"};
    const double per_iter[] = {1, 1, 1, 1, 1, 1, 2, 4, 1};
    for (int mode = 0; mode < 9; mode++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: k<0><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 1: k<1><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 2: k<2><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 3: k<3><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 4: k<4><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 5: k<5><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 6: k<6><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 7: k<7><<<threads / 256, 256>>>(tab, n_entries, per); break;
                case 8: k<8><<<threads / 256, 256>>>(tab, n_entries, per); break;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double n = (double)threads * per * per_iter[mode];
        printf("%-34s %8.3f ms  %7.1f G reductions/s   (%s)\n", names[mode], best, n / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
