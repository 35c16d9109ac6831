// Shared helpers for the sm_100a kernels behind include/seal3d_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define S3D_API extern "C" __attribute__((visibility("default")))

// C-ABI error convention: 0 = ok, >0 = cudaError_t of the launch, <0 = argument error
#define S3D_EINVAL (-22)
#define S3D_ENOTSUP (-95)

#define S3D_RETURN_LAST()                          \
    do {                                           \
        cudaError_t e__ = cudaPeekAtLastError();   \
        return (int)e__;                           \
    } while (0)

template <typename T>
__host__ __device__ inline T div_up(T a, T b) { return (a + b - 1) / b; }

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Stream-ordered scratch (cudaMallocAsync): keep freed blocks in the device's default pool instead of returning them to
// the OS at every synchronisation point (the default release threshold is 0, which turns a per-step scratch buffer
// into a cudaMalloc + cudaFree per step as soon as the host reads anything back).
static inline cudaError_t scratch_alloc(void **p, size_t bytes, cudaStream_t st) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        configured[dev] = true;
    }
    return cudaMallocAsync(p, bytes, st);
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

__device__ __forceinline__ uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}
