// TensoRF vector-matrix (VM) decomposition lookups for sm_100a (BASELINE config 5).
//
// The reference has no kernel here: tensoRF/network.py:99-151 calls F.grid_sample(bilinear, zeros, align_corners=True)
// twelve times per query batch (3 planes + 3 lines for sigma, the same for colour) on [1,R,H,W] images, i.e. channel-major
// planes where the 4 taps of every one of the R channels are four separate scalar gathers (16 / 48 channels -> 64 / 192
// scattered 4-byte loads per plane), plus ~10 stack / view / cat / sum launches around them.
//
// Here one kernel does all six lookups of a field, the plane x line products and (for sigma) the channel and plane sums:
//   * planes and lines are stored CHANNEL-LAST ([H,W,R] / [D,R]; the host keeps the nn.Parameters in torch's
//     channels_last memory format, so the state-dict shapes stay [1,R,H,W] / [1,R,D,1]): one tap of all R channels is
//     one contiguous 4R-byte run;
//   * a thread owns 4 channels of one sample (R/4 consecutive lanes per sample), so every tap is a 128-bit load and the
//     lanes of a sample read one contiguous 64 / 192-byte run; the sigma reduction over channels is a shuffle tree;
//   * the aabb normalisation of network.py:158 is evaluated in the kernel with the reference's operation order.
// The backward recomputes both factors and scatters with one 128-bit RED per (tap, 4 channels).
//
// Arithmetic follows ATen's grid_sampler_2d (align_corners: ix = ((x+1)/2)*(W-1); taps nw, ne, sw, se with weights
// (ix_se-ix)(iy_se-iy) ...; out-of-image taps contribute zero).  A line is the W = 1 image sampled at x = 0
// (network.py:106-107): only column 0 is in bounds, the weights reduce to (iy_se - iy) and (iy - iy_nw).
#include "common.cuh"

namespace {

struct VmArgs {
    const float *mat[3];
    const float *vec[3];
    int H[3], W[3], D[3];
};
struct VmGradArgs {
    float *mat[3];
    float *vec[3];
};

// tensoRF/network.py:37-38: plane i is indexed by (x[mat0] -> W, x[mat1] -> H), line i by x[vecid]
__device__ __constant__ int kMat0[3] = {0, 0, 1}, kMat1[3] = {1, 2, 2}, kVecId[3] = {2, 1, 0};

struct Taps {
    int x0, y0;
    float w[4];   // nw, ne, sw, se
    bool ok[4];
};

__device__ __forceinline__ void plane_taps(float gx, float gy, int H, int W, Taps &t) {
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const float ex = __fsub_rn(__fadd_rn(fx, 1.0f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.0f), iy);   // ix_se - ix, iy_se - iy
    const float dx = __fsub_rn(ix, fx), dy = __fsub_rn(iy, fy);
    // clamp before the int conversion so far-away points cannot overflow it; they are out of bounds either way
    t.x0 = (int)fminf(fmaxf(fx, -2.0f), (float)W);
    t.y0 = (int)fminf(fmaxf(fy, -2.0f), (float)H);
    t.w[0] = __fmul_rn(ex, ey); t.w[1] = __fmul_rn(dx, ey); t.w[2] = __fmul_rn(ex, dy); t.w[3] = __fmul_rn(dx, dy);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
        t.ok[k] = xx >= 0 && xx < W && yy >= 0 && yy < H;
    }
}

__device__ __forceinline__ float4 fma4(float4 v, float w, float4 a) {
    return make_float4(__fmaf_rn(v.x, w, a.x), __fmaf_rn(v.y, w, a.y), __fmaf_rn(v.z, w, a.z), __fmaf_rn(v.w, w, a.w));
}

// x normalised into the aabb: 2 * (x - lo) / (hi - lo) - 1 (network.py:158), elementwise float32 like torch
__device__ __forceinline__ void load_point(const float *__restrict__ xyz, uint32_t m, const float *__restrict__ aabb, float (&x)[3]) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float v = __ldg(xyz + (size_t)m * 3 + d);
        x[d] = aabb ? __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fsub_rn(v, __ldg(aabb + d))), __fsub_rn(__ldg(aabb + 3 + d), __ldg(aabb + d))), 1.0f) : v;
    }
}

// both factors of plane i for channels [4g, 4g+4)
__device__ __forceinline__ void factors(const VmArgs &a, int i, const float (&x)[3], uint32_t R, uint32_t g, Taps &tp, Taps &tl,
                                        float4 &fm, float4 &fv) {
    plane_taps(x[kMat0[i]], x[kMat1[i]], a.H[i], a.W[i], tp);
    plane_taps(0.0f, x[kVecId[i]], a.D[i], 1, tl);
    fm = make_float4(0.f, 0.f, 0.f, 0.f);
    fv = fm;
    float4 tm[4], tv[2];
    // issue all six loads, then accumulate in the reference's tap order
#pragma unroll
    for (int k = 0; k < 4; k++) {
        tm[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tp.ok[k]) tm[k] = __ldg(reinterpret_cast<const float4 *>(a.mat[i] + ((size_t)(tp.y0 + (k >> 1)) * a.W[i] + tp.x0 + (k & 1)) * R) + g);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        tv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tl.ok[2 * k]) tv[k] = __ldg(reinterpret_cast<const float4 *>(a.vec[i] + (size_t)(tl.y0 + k) * R) + g);
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (tp.ok[k]) fm = fma4(tm[k], tp.w[k], fm);
#pragma unroll
    for (int k = 0; k < 2; k++)
        if (tl.ok[2 * k]) fv = fma4(tv[k], tl.w[2 * k], fv);
}

// thread = (sample, 4-channel group); G = R / 4 lanes per sample.  REDUCE: out[M] (G a power of two <= 32), else out[M, 3R].
// T = float, or __half for the unreduced colour features of an fp16 step: the product is rounded to fp16 once on the way out --
// the same value the reference's autocast F.linear (basis_mat) makes of it -- instead of in a separate cast pass.
template <bool REDUCE, typename T = float>
__global__ void __launch_bounds__(256)
k_vm_forward(const float *__restrict__ xyz, uint32_t M, const float *__restrict__ aabb, VmArgs a, uint32_t R, uint32_t G,
             T *__restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // whole warps stay alive for the shuffles
    const bool live = t < (uint64_t)M * G;
    const uint32_t m = live ? (uint32_t)(t / G) : 0, g = live ? (uint32_t)(t % G) : 0;
    float x[3];
    load_point(xyz, m, aabb, x);
    float total = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Taps tp, tl;
        float4 fm, fv;
        factors(a, i, x, R, g, tp, tl, fm, fv);
        const float4 p = make_float4(fm.x * fv.x, fm.y * fv.y, fm.z * fv.z, fm.w * fv.w);
        if constexpr (REDUCE) {
            float s = (p.x + p.y) + (p.z + p.w);
            for (uint32_t o = 1; o < G; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            total += s;
        } else if (live) {
            if constexpr (sizeof(T) == 2) {
                const __half2 lo = __floats2half2_rn(p.x, p.y), hi = __floats2half2_rn(p.z, p.w);
                reinterpret_cast<uint2 *>(out + (size_t)m * 3 * R + (size_t)i * R)[g] =
                    make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
            } else {
                reinterpret_cast<float4 *>(out + (size_t)m * 3 * R + (size_t)i * R)[g] = p;
            }
        }
    }
    if constexpr (REDUCE) { if (live && g == 0) out[m] = total; }
}

template <bool REDUCE, typename T = float>
__global__ void __launch_bounds__(256)
k_vm_backward(const float *__restrict__ xyz, uint32_t M, const float *__restrict__ aabb, VmArgs a, uint32_t R, uint32_t G,
              const T *__restrict__ grad, VmGradArgs ga) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)M * G) return;
    const uint32_t m = (uint32_t)(t / G), g = (uint32_t)(t % G);
    float x[3];
    load_point(xyz, m, aabb, x);
    float4 go = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (REDUCE) { const float v = __ldg(grad + m); go = make_float4(v, v, v, v); }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        Taps tp, tl;
        float4 fm, fv;
        factors(a, i, x, R, g, tp, tl, fm, fv);
        if constexpr (!REDUCE) {
            if constexpr (sizeof(T) == 2) {
                const uint2 raw = __ldg(reinterpret_cast<const uint2 *>(grad + (size_t)m * 3 * R + (size_t)i * R) + g);
                const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
                go = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                go = __ldg(reinterpret_cast<const float4 *>(grad + (size_t)m * 3 * R + (size_t)i * R) + g);
            }
        }
        const float4 gm = make_float4(go.x * fv.x, go.y * fv.y, go.z * fv.z, go.w * fv.w);   // d/d(plane factor)
        const float4 gv = make_float4(go.x * fm.x, go.y * fm.y, go.z * fm.z, go.w * fm.w);   // d/d(line factor)
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (tp.ok[k])
                atomicAdd(reinterpret_cast<float4 *>(ga.mat[i] + ((size_t)(tp.y0 + (k >> 1)) * a.W[i] + tp.x0 + (k & 1)) * R) + g,
                          make_float4(gm.x * tp.w[k], gm.y * tp.w[k], gm.z * tp.w[k], gm.w * tp.w[k]));
#pragma unroll
        for (int k = 0; k < 2; k++)
            if (tl.ok[2 * k])
                atomicAdd(reinterpret_cast<float4 *>(ga.vec[i] + (size_t)(tl.y0 + k) * R) + g,
                          make_float4(gv.x * tl.w[2 * k], gv.y * tl.w[2 * k], gv.z * tl.w[2 * k], gv.w * tl.w[2 * k]));
    }
}

// bilinear resize of a channel-last image [H,W,R] -> [H2,W2,R], align_corners=True: what upsample_params
// (tensoRF/network.py:263-270) asks of F.interpolate.  ATen upsample_bilinear2d: src = dst * (in-1)/(out-1),
// i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1; out = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11).
__global__ void k_vm_resize(const float *__restrict__ src, int H, int W, float *__restrict__ dst, int H2, int W2, uint32_t R) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)H2 * W2 * R) return;
    const uint32_t r = (uint32_t)(t % R);
    const uint32_t px = (uint32_t)((t / R) % W2), py = (uint32_t)(t / ((uint64_t)R * W2));
    const float sh = H2 > 1 ? __fdiv_rn((float)(H - 1), (float)(H2 - 1)) : 0.0f, sw = W2 > 1 ? __fdiv_rn((float)(W - 1), (float)(W2 - 1)) : 0.0f;
    const float fy = __fmul_rn(sh, (float)py), fx = __fmul_rn(sw, (float)px);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly1 = __fsub_rn(fy, (float)y0), lx1 = __fsub_rn(fx, (float)x0);
    const float ly0 = __fsub_rn(1.0f, ly1), lx0 = __fsub_rn(1.0f, lx1);
    const float v00 = src[((size_t)y0 * W + x0) * R + r], v01 = src[((size_t)y0 * W + x1) * R + r];
    const float v10 = src[((size_t)y1 * W + x0) * R + r], v11 = src[((size_t)y1 * W + x1) * R + r];
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01)), bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    dst[t] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

int fill_args(VmArgs &a, const float *m0, const float *m1, const float *m2, const float *v0, const float *v1, const float *v2,
              const int *h_dims, uint32_t R, int reduce, uint32_t &G) {
    if (R == 0 || R % 4 != 0 || h_dims == nullptr) return S3D_EINVAL;
    G = R / 4;
    if (reduce && (G > 32 || (G & (G - 1)) != 0)) return S3D_ENOTSUP;   // the shuffle tree needs a power-of-two group inside one warp
    a.mat[0] = m0; a.mat[1] = m1; a.mat[2] = m2; a.vec[0] = v0; a.vec[1] = v1; a.vec[2] = v2;
    for (int i = 0; i < 3; i++) {
        a.H[i] = h_dims[i * 3]; a.W[i] = h_dims[i * 3 + 1]; a.D[i] = h_dims[i * 3 + 2];
        if (a.H[i] < 1 || a.W[i] < 1 || a.D[i] < 1) return S3D_EINVAL;
    }
    return 0;
}

}  // namespace

// tensoRF/network.py:99-123 get_sigma_feat (reduce = 1: out [M]) and :126-146 the mat_feat * vec_feat product of
// get_color_feat (reduce = 0: out [M, 3R], plane-major / channel-minor like the torch.cat at :141-142).
S3D_API int s3d_vm_forward(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                           const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, int reduce,
                           float *out, void *stream) {
    if (M == 0) return 0;
    VmArgs a;
    uint32_t G;
    if (int rc = fill_args(a, mat0, mat1, mat2, vec0, vec1, vec2, h_dims, R, reduce, G)) return rc;
    const unsigned blocks = (unsigned)div_up((uint64_t)M * G, (uint64_t)256);
    if (reduce) k_vm_forward<true><<<blocks, 256, 0, as_stream(stream)>>>(xyz, M, aabb, a, R, G, out);
    else k_vm_forward<false><<<blocks, 256, 0, as_stream(stream)>>>(xyz, M, aabb, a, R, G, out);
    S3D_RETURN_LAST();
}

// the unreduced lookup with fp16 output [M, 3R] / fp16 output gradient (the colour features of an fp16 step)
S3D_API int s3d_vm_forward_f16(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                               const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, void *out, void *stream) {
    if (M == 0) return 0;
    VmArgs a;
    uint32_t G;
    if (int rc = fill_args(a, mat0, mat1, mat2, vec0, vec1, vec2, h_dims, R, 0, G)) return rc;
    const unsigned blocks = (unsigned)div_up((uint64_t)M * G, (uint64_t)256);
    k_vm_forward<false, __half><<<blocks, 256, 0, as_stream(stream)>>>(xyz, M, aabb, a, R, G, (__half *)out);
    S3D_RETURN_LAST();
}

S3D_API int s3d_vm_backward_f16(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                                const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, const void *grad,
                                float *g_mat0, float *g_mat1, float *g_mat2, float *g_vec0, float *g_vec1, float *g_vec2, void *stream) {
    if (M == 0) return 0;
    VmArgs a;
    uint32_t G;
    if (int rc = fill_args(a, mat0, mat1, mat2, vec0, vec1, vec2, h_dims, R, 0, G)) return rc;
    VmGradArgs ga;
    ga.mat[0] = g_mat0; ga.mat[1] = g_mat1; ga.mat[2] = g_mat2; ga.vec[0] = g_vec0; ga.vec[1] = g_vec1; ga.vec[2] = g_vec2;
    const unsigned blocks = (unsigned)div_up((uint64_t)M * G, (uint64_t)256);
    k_vm_backward<false, __half><<<blocks, 256, 0, as_stream(stream)>>>(xyz, M, aabb, a, R, G, (const __half *)grad, ga);
    S3D_RETURN_LAST();
}

// gradients of the six factor images (channel-last, float32, ACCUMULATED into: the caller zero-fills) for d(out) = grad
S3D_API int s3d_vm_backward(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                            const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, int reduce,
                            const float *grad, float *g_mat0, float *g_mat1, float *g_mat2, float *g_vec0, float *g_vec1, float *g_vec2,
                            void *stream) {
    if (M == 0) return 0;
    VmArgs a;
    uint32_t G;
    if (int rc = fill_args(a, mat0, mat1, mat2, vec0, vec1, vec2, h_dims, R, reduce, G)) return rc;
    VmGradArgs ga;
    ga.mat[0] = g_mat0; ga.mat[1] = g_mat1; ga.mat[2] = g_mat2; ga.vec[0] = g_vec0; ga.vec[1] = g_vec1; ga.vec[2] = g_vec2;
    const unsigned blocks = (unsigned)div_up((uint64_t)M * G, (uint64_t)256);
    if (reduce) k_vm_backward<true><<<blocks, 256, 0, as_stream(stream)>>>(xyz, M, aabb, a, R, G, grad, ga);
    else k_vm_backward<false><<<blocks, 256, 0, as_stream(stream)>>>(xyz, M, aabb, a, R, G, grad, ga);
    S3D_RETURN_LAST();
}

// upsample_params (tensoRF/network.py:263-270): one channel-last image; a line is the W = W2 = 1 case
S3D_API int s3d_vm_resize(const float *src, uint32_t H, uint32_t W, float *dst, uint32_t H2, uint32_t W2, uint32_t R, void *stream) {
    if (H == 0 || W == 0 || R == 0) return S3D_EINVAL;
    const uint64_t n = (uint64_t)H2 * W2 * R;
    if (n == 0) return 0;
    k_vm_resize<<<(unsigned)div_up(n, (uint64_t)256), 256, 0, as_stream(stream)>>>(src, (int)H, (int)W, dst, (int)H2, (int)W2, R);
    S3D_RETURN_LAST();
}

// ---- TensoRF colour head: frequency encodings + concatenation + padding in one launch ------------------------------------
// tensoRF/network.py:170-172:  h = cat([encoder(color_feat), encoder_dir(d)])  with both encoders `frequency`, multires = deg
// (freqencoder layout: [x | sin(2^0 x) | cos(2^0 x) | sin(2^1 x) | cos(2^1 x) ...], freqencoder.cu:30-60).  feat arrives fp16
// from the tensor-core basis_mat, h leaves fp16, zero padded to K columns, ready to be the A operand of the colour MLP:
// replaces two encoder launches, a concatenation, a pad and two casts.
namespace {
// One WARP per row: lane d < Fd encodes feature d, lanes Fd .. Fd+2 the direction components, each writing its 1 + 2 deg values
// into a shared-memory image of the row; the warp then stores the row with coalesced 4-byte stores.  (One thread per output
// column pays two integer divisions and a sine per value: 1.2 ms for 1.36 M rows; this form 0.1 ms.)  Needs Fd + 3 <= 32.
constexpr uint32_t kHeadMaxK = 512;
__global__ void __launch_bounds__(256)
k_head_encode(const __half *__restrict__ feat, uint32_t Fd, uint32_t ld_feat, const float *__restrict__ dirs, uint32_t B, uint32_t deg, uint32_t K,
              __half *__restrict__ h) {
    __shared__ __align__(16) __half s_row[8][kHeadMaxK];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t C1 = Fd * (1 + 2 * deg), C2 = 3 * (1 + 2 * deg);
    __half *row = s_row[wib];
    for (uint32_t c = C1 + C2 + lane; c < K; c += 32) row[c] = __float2half_rn(0.0f);     // the zero padding, written once
    __syncwarp();
    for (uint32_t b = blockIdx.x * 8 + wib; b < B; b += gridDim.x * 8) {
        if (lane < Fd + 3) {
            const bool is_dir = lane >= Fd;
            const uint32_t D = is_dir ? 3u : Fd, d = is_dir ? lane - Fd : lane;
            const float x = is_dir ? __ldg(dirs + (size_t)b * 3 + d) : __half2float(feat[(size_t)b * ld_feat + d]);
            __half *o = row + (is_dir ? C1 : 0u) + d;
            o[0] = __float2half_rn(x);
            for (uint32_t f = 0; f < deg; f++) {
                const float xs = __fmul_rn(x, __int_as_float((127 + f) << 23));
                o[(1 + 2 * f) * D] = __float2half_rn(__sinf(xs));
                o[(2 + 2 * f) * D] = __float2half_rn(__sinf(__fadd_rn(xs, 1.5707963267948966f)));
            }
        }
        __syncwarp();
        const uint32_t *src = reinterpret_cast<const uint32_t *>(row);
        uint32_t *dst = reinterpret_cast<uint32_t *>(h + (size_t)b * K);
        for (uint32_t c = lane; c < K / 2; c += 32) dst[c] = src[c];
        __syncwarp();
    }
}
// d feat = g_x + sum_f 2^f (g_sin cos - g_cos sin), sin / cos taken from the stored encoding (freqencoder.cu:63-94)
__global__ void __launch_bounds__(256)
k_head_encode_bwd(const __half *__restrict__ gh, const __half *__restrict__ h, uint32_t Fd, uint32_t ld_feat, uint32_t B, uint32_t deg, uint32_t K,
                  __half *__restrict__ gfeat) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * ld_feat) return;
    const uint32_t b = (uint32_t)(t / ld_feat), d = (uint32_t)(t - (uint64_t)b * ld_feat);
    float r = 0.0f;
    if (d < Fd) {
        const __half *g = gh + (size_t)b * K, *o = h + (size_t)b * K;
        r = __half2float(g[d]);
        for (uint32_t f = 0; f < deg; f++) {
            const uint32_t is = (1 + 2 * f) * Fd + d, ic = (2 + 2 * f) * Fd + d;
            r += __int_as_float((127 + f) << 23) * (__half2float(g[is]) * __half2float(o[ic]) - __half2float(g[ic]) * __half2float(o[is]));
        }
    }
    gfeat[t] = __float2half_rn(r);           // columns Fd .. ld_feat-1 (padding of the feature rows) get zero
}
}  // namespace

// feat fp16 [B, ld_feat] (Fd valid columns), dirs fp32 [B,3] -> h fp16 [B,K] = [freq(feat) | freq(dirs) | 0 ...],  K even, >= (Fd + 3)(1 + 2 deg)
S3D_API int s3d_tensorf_head_encode(const void *feat, uint32_t Fd, uint32_t ld_feat, const float *dirs, uint32_t B, uint32_t deg, uint32_t K, void *h,
                                    void *stream) {
    if (B == 0) return 0;
    if (K < (Fd + 3) * (1 + 2 * deg) || (K & 1u) || deg > 16 || ld_feat < Fd) return S3D_EINVAL;
    if (Fd + 3 > 32 || K > kHeadMaxK) return S3D_ENOTSUP;
    k_head_encode<<<min(div_up(B, 8u), 148u * 16u), 256, 0, as_stream(stream)>>>((const __half *)feat, Fd, ld_feat, dirs, B, deg, K, (__half *)h);
    S3D_RETURN_LAST();
}
// grad_feat fp16 [B, ld_feat] (padding columns zeroed)
S3D_API int s3d_tensorf_head_encode_backward(const void *grad_h, const void *h, uint32_t Fd, uint32_t ld_feat, uint32_t B, uint32_t deg, uint32_t K,
                                             void *grad_feat, void *stream) {
    if (B == 0) return 0;
    if (K < (Fd + 3) * (1 + 2 * deg) || deg > 16 || ld_feat < Fd) return S3D_EINVAL;
    k_head_encode_bwd<<<(unsigned)div_up((uint64_t)B * ld_feat, (uint64_t)256), 256, 0, as_stream(stream)>>>((const __half *)grad_h, (const __half *)h, Fd, ld_feat, B, deg, K,
                                                                                                            (__half *)grad_feat);
    S3D_RETURN_LAST();
}
