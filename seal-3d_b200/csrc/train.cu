// Training-step glue kernels for sm_100a: distillation losses (+ their gradients), fused Adam over the
// gradient arena with fp16 shadow refresh, density-grid EMA update.
//
//   pretrain (per sample)  SealNeRF/trainer.py:456-469   L1(sigma) + L1(rgb), mean reductions
//   finetune (per ray)     nerf/utils.py:484-489,530     mean_rays(mean_c (rgb - gt)^2) + mean |depth - gt|
//   Adam                   main_SealNeRF.py:283-284      torch.optim.Adam(betas=(0.9,0.99), eps=1e-15), no weight decay
//   density grid           nerf/renderer.py:521-524      grid = max(grid*decay, tmp) where both >= 0; mean of clamp(grid,0)
#include "common.cuh"

namespace {

__device__ __forceinline__ float sgnf(float v) { return (v > 0.0f) - (v < 0.0f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// loss[0] += sum|ds| / M + sum|dc| / (3M); grads are those of the mean reductions
__global__ void k_pretrain_loss(const float *__restrict__ sig_s, const float *__restrict__ rgb_s, const float *__restrict__ sig_t,
                                const float *__restrict__ rgb_t, uint32_t M, float *__restrict__ loss,
                                float *__restrict__ g_sig, float *__restrict__ g_rgb) {
    float acc = 0.0f;
    const float inv_m = 1.0f / (float)M, inv_3m = 1.0f / (3.0f * (float)M);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const float ds = sig_s[i] - sig_t[i];
        acc += fabsf(ds) * inv_m;
        if (g_sig) g_sig[i] = sgnf(ds) * inv_m;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float dc = rgb_s[(size_t)i * 3 + c] - rgb_t[(size_t)i * 3 + c];
            acc += fabsf(dc) * inv_3m;
            if (g_rgb) g_rgb[(size_t)i * 3 + c] = sgnf(dc) * inv_3m;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc);
}

// student image = comp + (1 - ws) * bg.  loss[0] += MSE part, loss[1] += L1 depth part.
// grad_image = d loss / d comp ; grad_ws = d loss / d ws = -bg * sum_c grad_image_c.
__global__ void k_finetune_loss(const float *__restrict__ comp_s, const float *__restrict__ ws_s, const float *__restrict__ depth_s,
                                const float *__restrict__ image_t, const float *__restrict__ depth_t, uint32_t N, float bg,
                                float *__restrict__ loss, float *__restrict__ grad_image, float *__restrict__ grad_ws) {
    float a0 = 0.0f, a1 = 0.0f;
    const float inv_n = 1.0f / (float)N, k = 2.0f / (3.0f * (float)N);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float back = (1.0f - ws_s[i]) * bg;
        float gsum = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float d = comp_s[(size_t)i * 3 + c] + back - image_t[(size_t)i * 3 + c];
            a0 += d * d * (inv_n / 3.0f);
            const float g = k * d;
            grad_image[(size_t)i * 3 + c] = g;
            gsum += g;
        }
        grad_ws[i] = -bg * gsum;
        if (depth_t) a1 += fabsf(depth_s[i] - depth_t[i]) * inv_n;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1);
    if ((threadIdx.x & 31) == 0) { atomicAdd(loss, a0); atomicAdd(loss + 1, a1); }
}

// One pass over the arena: g = grad * grad_scale; Adam; optional fp16 shadow of the new parameter; grad zeroed.
template <typename G>
__global__ void __launch_bounds__(256)
k_adam(float *__restrict__ p, G *__restrict__ g, float *__restrict__ m, float *__restrict__ v, __half *__restrict__ shadow,
       size_t n, float lr_over_bc1, float inv_sqrt_bc2, float beta1, float beta2, float eps, float grad_scale, int zero_grad,
       const float *__restrict__ scaler, float lr) {
    bool skip = false;
    if (scaler) {   // GradScaler semantics (see s3d_grad_scaler_check): device-side scale, skip decision and bias corrections
        skip = scaler[2] != 0.0f;
        grad_scale = grad_scale / scaler[0];
        lr_over_bc1 = lr * scaler[4];
        inv_sqrt_bc2 = scaler[5];
    }
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    if (skip) {   // a non-finite gradient somewhere in the arena: the optimizer step is skipped, the gradient is still cleared
        if (zero_grad)
            for (size_t i = i0; i < n && i < i0 + 4; i++) {
                if constexpr (sizeof(G) == 4) reinterpret_cast<float *>(g)[i] = 0.0f;
                else reinterpret_cast<__half *>(g)[i] = __float2half_rn(0.0f);
            }
        return;
    }
    const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                           reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(shadow) & 7) == 0;
    if (i0 + 4 <= n && sizeof(G) == 4 && aligned) {
        float4 pp = *reinterpret_cast<float4 *>(p + i0), gg = *reinterpret_cast<float4 *>(reinterpret_cast<float *>(g) + i0);
        float4 mm = *reinterpret_cast<float4 *>(m + i0), vv = *reinterpret_cast<float4 *>(v + i0);
        float *P = &pp.x, *Gp = &gg.x, *Mp = &mm.x, *Vp = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float gr = Gp[k] * grad_scale;
            Mp[k] = beta1 * Mp[k] + (1.0f - beta1) * gr;
            Vp[k] = beta2 * Vp[k] + (1.0f - beta2) * gr * gr;
            P[k] -= lr_over_bc1 * Mp[k] / (sqrtf(Vp[k]) * inv_sqrt_bc2 + eps);
        }
        *reinterpret_cast<float4 *>(p + i0) = pp;
        *reinterpret_cast<float4 *>(m + i0) = mm;
        *reinterpret_cast<float4 *>(v + i0) = vv;
        if (zero_grad) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(g) + i0) = make_float4(0, 0, 0, 0);
        if (shadow) {
            __half2 a = __floats2half2_rn(pp.x, pp.y), b = __floats2half2_rn(pp.z, pp.w);
            *reinterpret_cast<uint2 *>(shadow + i0) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
        }
    } else {
        for (size_t i = i0; i < n && i < i0 + 4; i++) {
            float gr;
            if constexpr (sizeof(G) == 4) gr = reinterpret_cast<float *>(g)[i] * grad_scale;
            else gr = __half2float(reinterpret_cast<__half *>(g)[i]) * grad_scale;
            const float mi = beta1 * m[i] + (1.0f - beta1) * gr;
            const float vi = beta2 * v[i] + (1.0f - beta2) * gr * gr;
            m[i] = mi; v[i] = vi;
            const float pn = p[i] - lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
            p[i] = pn;
            if (shadow) shadow[i] = __float2half_rn(pn);
            if (zero_grad) {
                if constexpr (sizeof(G) == 4) reinterpret_cast<float *>(g)[i] = 0.0f;
                else reinterpret_cast<__half *>(g)[i] = __float2half_rn(0.0f);
            }
        }
    }
}

// ---- GradScaler state on the device (torch.cuda.amp.GradScaler semantics, nerf/utils.py:857-859) ----------------
// state = float[16], two 8-float blocks with the same slot meaning, one per parameter family (torch.optim.Adam keeps a step
// count per parameter: after a pretraining stage that trains the hash tables only, nerf SealNeRF/trainer.py:484-488, the MLP's
// bias correction must start at t = 1 while the tables' continues):
//   block 0 (tables, and any flat arena):  [0] scale | [1] growth tracker | [2] found_inf | [3] optimizer steps applied so far
//                                          | [4] 1/(1-beta1^t) | [5] 1/sqrt(1-beta2^t)
//   block 1 (state + 8, the MLP arena):    [8] scale (copy) | [10] found_inf (copy) | [11] steps | [12], [13] bias corrections
// The Adam kernels read slots 0, 2, 4, 5 of whichever block they are handed.
__global__ void __launch_bounds__(256)
k_found_inf(const float *__restrict__ g, size_t n, float *__restrict__ state) {
    bool bad = false;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (size_t)gridDim.x * blockDim.x * 4) {
        if (i + 4 <= n && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            const float4 v = *reinterpret_cast<const float4 *>(g + i);
            bad |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
        } else {
            for (size_t j = i; j < n && j < i + 4; j++) bad |= !isfinite(g[j]);
        }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) state[2] = 1.0f;
}
__global__ void k_scaler_prepare(float *__restrict__ state, float beta1, float beta2, int advance_mlp) {
    if (state[2] == 0.0f) {   // the step will be applied: advance the optimizer's step counts, refresh the bias corrections
        const float t = state[3] + 1.0f;
        state[3] = t;
        state[4] = (float)(1.0 / (1.0 - pow((double)beta1, (double)t)));
        state[5] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)t)));
        if (advance_mlp) {
            const float tm = state[11] + 1.0f;
            state[11] = tm;
            state[12] = (float)(1.0 / (1.0 - pow((double)beta1, (double)tm)));
            state[13] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)tm)));
        }
    }
    state[8] = state[0];
    state[10] = state[2];
}
__global__ void k_scaler_update(float *__restrict__ state, float growth, float backoff, float interval) {
    if (state[2] != 0.0f) { state[0] *= backoff; state[1] = 0.0f; }
    else {
        const float tr = state[1] + 1.0f;
        if (tr >= interval) { state[0] *= growth; state[1] = 0.0f; }
        else state[1] = tr;
    }
    state[2] = 0.0f;
    state[8] = state[0];
    state[10] = 0.0f;
}

// torch_ema.ExponentialMovingAverage.update: shadow -= (1 - decay) * (shadow - param)
__global__ void __launch_bounds__(256)
k_ema_update(float *__restrict__ shadow, const float *__restrict__ param, size_t n, float one_minus_decay) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float sv = shadow[i];
        shadow[i] = sv - one_minus_decay * (sv - param[i]);
    }
}

__global__ void k_f32_to_f16_arena(const float *__restrict__ src, __half *__restrict__ dst, size_t n) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const float2 f = *reinterpret_cast<const float2 *>(src + i);
        *reinterpret_cast<__half2 *>(dst + i) = __floats2half2_rn(f.x, f.y);
    } else if (i < n) dst[i] = __float2half_rn(src[i]);
}

// grid[i] = max(grid[i]*decay, tmp[i]) where both >= 0 ; sum += max(grid[i], 0)
__global__ void k_density_ema(float *__restrict__ grid, const float *__restrict__ tmp, uint32_t n, float decay, float *__restrict__ sum) {
    float acc = 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        const float t = tmp[i];
        if (g >= 0.0f && t >= 0.0f) { g = fmaxf(g * decay, t); grid[i] = g; }
        acc += fmaxf(g, 0.0f);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, acc);
}

// nerf/renderer.py:379-443 mark_untrained_grid: a cell of cascade `cas` is "trained" if its centre, taken to camera space
// (cam = (x - t) . R, poses are cam2world), lies in front of at least one camera and inside its frustum widened by one cell:
// |cam.x| < cx/fx * cam.z + 2*half_grid_size (and the same for y).  Cells no camera sees get density -1.  The reference does
// this with a 5-level Python loop of batched matmuls; here one thread owns one (cascade, cell) and walks the cameras
// (poses in shared memory).  Float32 with the reference's operation order: centre = (2*c/(H-1) - 1) * (bound - hgs).
__global__ void __launch_bounds__(256)
k_mark_untrained(float *__restrict__ grid, const float *__restrict__ poses, uint32_t B, float kx, float ky, uint32_t C, uint32_t H,
                 float bound, int *__restrict__ count_out) {
    extern __shared__ float s_pose[];   // per camera: R (9, row-major) + t (3)
    for (uint32_t i = threadIdx.x; i < B * 12; i += blockDim.x) {
        const uint32_t b = i / 12, k = i - b * 12;
        s_pose[i] = k < 9 ? poses[(size_t)b * 16 + (k / 3) * 4 + (k % 3)] : poses[(size_t)b * 16 + (k - 9) * 4 + 3];
    }
    __syncthreads();
    const uint32_t H3 = H * H * H;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * H3) return;
    const uint32_t cas = t / H3, cell = t - cas * H3;
    const uint32_t cx = cell / (H * H), cy = (cell / H) % H, cz = cell % H;
    auto spread = [](uint32_t v) {
        v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
        v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v; };
    const uint32_t morton = spread(cx) | (spread(cy) << 1) | (spread(cz) << 2);
    const float bound_cas = fminf((float)(1u << cas), bound);
    const float hgs = __fdiv_rn(bound_cas, (float)H);
    const float sc = __fsub_rn(bound_cas, hgs);
    float w[3];
    const uint32_t c3[3] = {cx, cy, cz};
#pragma unroll
    for (int d = 0; d < 3; d++) w[d] = __fmul_rn(__fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)c3[d]), (float)(H - 1)), 1.0f), sc);
    const float margin = __fmul_rn(hgs, 2.0f);
    int count = 0;
    for (uint32_t b = 0; b < B; b++) {
        const float *P = s_pose + b * 12;
        const float dx = __fsub_rn(w[0], P[9]), dy = __fsub_rn(w[1], P[10]), dz = __fsub_rn(w[2], P[11]);
        // cam_j = dx*R[0][j] + dy*R[1][j] + dz*R[2][j]  (row vector times R)
        const float camx = __fadd_rn(__fadd_rn(__fmul_rn(dx, P[0]), __fmul_rn(dy, P[3])), __fmul_rn(dz, P[6]));
        const float camy = __fadd_rn(__fadd_rn(__fmul_rn(dx, P[1]), __fmul_rn(dy, P[4])), __fmul_rn(dz, P[7]));
        const float camz = __fadd_rn(__fadd_rn(__fmul_rn(dx, P[2]), __fmul_rn(dy, P[5])), __fmul_rn(dz, P[8]));
        const bool in = camz > 0.0f && fabsf(camx) < __fadd_rn(__fmul_rn(kx, camz), margin) && fabsf(camy) < __fadd_rn(__fmul_rn(ky, camz), margin);
        count += in ? 1 : 0;
    }
    if (count_out) count_out[(size_t)cas * H3 + morton] = count;
    if (count == 0) grid[(size_t)cas * H3 + morton] = -1.0f;
}

// nerf/utils.py:53-140 get_rays: pixel (row * W + col) of view b -> ray origin (the camera centre) and unit direction
// d = R . normalize((col + 0.5 - cx) / fx, (row + 0.5 - cy) / fy, 1), float32 in the reference's operation order
__global__ void k_get_rays(const float *__restrict__ poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t W,
                           const long long *__restrict__ inds, uint32_t inds_stride, uint32_t N, float *__restrict__ rays_o,
                           float *__restrict__ rays_d) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * N) return;
    const uint32_t b = (uint32_t)(t / N), n = (uint32_t)(t - (uint64_t)b * N);
    const long long pix = inds ? inds[(size_t)b * inds_stride + n] : (long long)n;
    const float i = __fadd_rn((float)(pix % W), 0.5f), j = __fadd_rn((float)(pix / W), 0.5f);
    const float x = __fdiv_rn(__fsub_rn(i, cx), fx), y = __fdiv_rn(__fsub_rn(j, cy), fy);
    const float norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), 1.0f));
    const float dx = __fdiv_rn(x, norm), dy = __fdiv_rn(y, norm), dz = __fdiv_rn(1.0f, norm);
    const float *P = poses + (size_t)b * 16;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        rays_d[t * 3 + k] = __fadd_rn(__fadd_rn(__fmul_rn(dx, P[k * 4]), __fmul_rn(dy, P[k * 4 + 1])), __fmul_rn(dz, P[k * 4 + 2]));
        rays_o[t * 3 + k] = P[k * 4 + 3];
    }
}


}  // namespace

S3D_API int s3d_pretrain_loss(const float *sigma_s, const float *rgb_s, const float *sigma_t, const float *rgb_t, uint32_t M,
                              float *loss, float *grad_sigma, float *grad_rgb, void *stream) {
    if (M == 0) return 0;
    k_pretrain_loss<<<min(div_up(M, 256u), 2048u), 256, 0, as_stream(stream)>>>(sigma_s, rgb_s, sigma_t, rgb_t, M, loss, grad_sigma, grad_rgb);
    S3D_RETURN_LAST();
}

S3D_API int s3d_finetune_loss(const float *comp_s, const float *ws_s, const float *depth_s, const float *image_t, const float *depth_t,
                              uint32_t N, float bg_color, float *loss, float *grad_image, float *grad_ws, void *stream) {
    if (N == 0) return 0;
    k_finetune_loss<<<min(div_up(N, 256u), 2048u), 256, 0, as_stream(stream)>>>(comp_s, ws_s, depth_s, image_t, depth_t, N, bg_color, loss, grad_image, grad_ws);
    S3D_RETURN_LAST();
}

// grad_dtype: 0 = float32, 1 = float16.  step >= 1 (bias correction).  shadow may be NULL.
S3D_API int s3d_adam_step(float *params, void *grads, float *exp_avg, float *exp_avg_sq, void *shadow_f16, uint64_t n, float lr,
                          float beta1, float beta2, float eps, uint32_t step, float grad_scale, int zero_grad, int grad_dtype,
                          const float *scaler_state, void *stream) {
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float lr_over_bc1 = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const unsigned blocks = (unsigned)div_up((size_t)n, (size_t)1024);
    if (grad_dtype == 0)
        k_adam<float><<<blocks, 256, 0, as_stream(stream)>>>(params, (float *)grads, exp_avg, exp_avg_sq, (__half *)shadow_f16, (size_t)n,
                                                              lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps, grad_scale, zero_grad, scaler_state, lr);
    else
        k_adam<__half><<<blocks, 256, 0, as_stream(stream)>>>(params, (__half *)grads, exp_avg, exp_avg_sq, (__half *)shadow_f16, (size_t)n,
                                                               lr_over_bc1, inv_sqrt_bc2, beta1, beta2, eps, grad_scale, zero_grad, scaler_state, lr);
    S3D_RETURN_LAST();
}

S3D_API int s3d_grad_scaler_check(const float *grads, uint64_t n, float *scaler_state, float beta1, float beta2, int advance_mlp, void *stream) {
    cudaStream_t st = as_stream(stream);
    if (n) k_found_inf<<<(unsigned)min((size_t)div_up((size_t)n, (size_t)1024), (size_t)2368), 256, 0, st>>>(grads, (size_t)n, scaler_state);
    k_scaler_prepare<<<1, 1, 0, st>>>(scaler_state, beta1, beta2, advance_mlp);
    S3D_RETURN_LAST();
}

S3D_API int s3d_grad_scaler_update(float *scaler_state, float growth_factor, float backoff_factor, uint32_t growth_interval, void *stream) {
    k_scaler_update<<<1, 1, 0, as_stream(stream)>>>(scaler_state, growth_factor, backoff_factor, (float)growth_interval);
    S3D_RETURN_LAST();
}

S3D_API int s3d_ema_update(float *shadow, const float *params, uint64_t n, float decay, void *stream) {
    if (n == 0) return 0;
    k_ema_update<<<(unsigned)min((size_t)div_up((size_t)n, (size_t)256), (size_t)4736), 256, 0, as_stream(stream)>>>(shadow, params, (size_t)n, 1.0f - decay);
    S3D_RETURN_LAST();
}

S3D_API int s3d_cast_f32_to_f16(const float *src, void *dst, uint64_t n, void *stream) {
    if (n == 0) return 0;
    k_f32_to_f16_arena<<<(unsigned)div_up((size_t)n, (size_t)512), 256, 0, as_stream(stream)>>>(src, (__half *)dst, (size_t)n);
    S3D_RETURN_LAST();
}

S3D_API int s3d_density_grid_ema(float *grid, const float *tmp_grid, uint32_t n, float decay, float *sum_out, void *stream) {
    if (n == 0) return 0;
    k_density_ema<<<min(div_up(n, 256u), 2048u), 256, 0, as_stream(stream)>>>(grid, tmp_grid, n, decay, sum_out);
    S3D_RETURN_LAST();
}


// nerf/renderer.py:379-443.  poses: device [B,4,4] cam2world; kx = cx/fx, ky = cy/fy; count_out (optional, int32 [C,H^3],
// morton order) receives the number of cameras that see each cell; density_grid[c, cell] = -1 where that number is 0.
S3D_API int s3d_mark_untrained_grid(float *density_grid, const float *poses, uint32_t B, float kx, float ky, uint32_t C, uint32_t H,
                                    float bound, int *count_out, void *stream) {
    if (C == 0 || H == 0) return S3D_EINVAL;
    if (B > 4000) return S3D_ENOTSUP;     // poses are staged in shared memory (48 B each)
    const uint32_t n = C * H * H * H;
    k_mark_untrained<<<div_up(n, 256u), 256, (size_t)B * 12 * sizeof(float), as_stream(stream)>>>(density_grid, poses, B, kx, ky, C, H, bound, count_out);
    S3D_RETURN_LAST();
}

// nerf/utils.py:53-140.  poses device [B,4,4]; inds device int64 [B or 1, N] (row * W + col; inds_rows == 1: shared by all
// views like the reference's expand) or NULL = all H*W pixels in order (then N must be H*W); rays_o / rays_d [B,N,3]
S3D_API int s3d_get_rays(const float *poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                         const long long *inds, uint32_t inds_rows, uint32_t N, float *rays_o, float *rays_d, void *stream) {
    if (B == 0 || N == 0) return 0;
    if (W == 0 || H == 0 || (inds == nullptr && N != H * W) || (inds && inds_rows != 1 && inds_rows != B)) return S3D_EINVAL;
    const uint64_t total = (uint64_t)B * N;
    k_get_rays<<<(unsigned)div_up(total, (uint64_t)256), 256, 0, as_stream(stream)>>>(poses, B, fx, fy, cx, cy, W, inds,
                                                                                       inds_rows == 1 ? 0u : N, N, rays_o, rays_d);
    S3D_RETURN_LAST();
}

