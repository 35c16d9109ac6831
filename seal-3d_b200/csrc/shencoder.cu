// Real spherical-harmonics direction encoding (degree <= 8) and frequency encoding for sm_100a.
//
// Replaces shencoder/src/shencoder.cu (entry points shencoder.h:9-10) and
// freqencoder/src/freqencoder.cu (freqencoder.h:7,10) of the reference.
//
// The reference stores the 64 basis polynomials and their 192 partial derivatives as
// hand-expanded monomials (shencoder.cu:49-121, :130-350).  Here the same polynomials are produced
// by structure: every real SH of band l, order m factors as
//        Y_l^{+m} = N_l^m * Q_l^m(z) * c_m(x,y)      Y_l^{-m} = N_l^m * Q_l^m(z) * s_m(x,y)
// with c_m + i s_m = (x + i y)^m and Q_l^m(z) = P_l^m(z) / (1-z^2)^{m/2} a polynomial in z.  The
// normalised coefficients of Q_l^m are generated at COMPILE TIME by the associated-Legendre
// recurrence into a __constant__ table, so the kernel is a fully unrolled Horner evaluation with
// immediate constant-bank operands, and the derivatives follow from the factorisation
// (d c_m/dx = m c_{m-1}, d c_m/dy = -m s_{m-1}, d s_m/dx = m s_{m-1}, d s_m/dy = m c_{m-1}).
// As functions of (x,y,z) -- also off the unit sphere -- these are the reference's polynomials.
#include "common.cuh"

namespace {

constexpr int kMaxDeg = 8;

struct ShTable {
    // coef[l][m][k]: coefficient of z^k in N_l^m * Q_l^m(z)  (N includes the sqrt(2) for m > 0)
    float coef[kMaxDeg][kMaxDeg][kMaxDeg];
};

constexpr double c_sqrt(double x) {
    double r = x > 1 ? x : 1;
    for (int i = 0; i < 60; i++) r = 0.5 * (r + x / r);
    return r;
}

constexpr ShTable make_sh_table() {
    ShTable t{};
    constexpr double kPi = 3.14159265358979323846;
    for (int m = 0; m < kMaxDeg; m++) {
        double q2[kMaxDeg] = {}, q1[kMaxDeg] = {};  // Q_{l-2}^m, Q_{l-1}^m as coefficient arrays
        for (int l = m; l < kMaxDeg; l++) {
            double q[kMaxDeg] = {};
            if (l == m) {
                double v = 1;
                for (int k = 1; k <= m; k++) v *= -(2.0 * k - 1.0);  // (-1)^m (2m-1)!!
                q[0] = v;
            } else {
                // (l-m) Q_l = (2l-1) z Q_{l-1} - (l+m-1) Q_{l-2}
                for (int k = 0; k < kMaxDeg; k++) {
                    double a = (k > 0 ? (2.0 * l - 1.0) * q1[k - 1] : 0.0) - (l + m - 1.0) * q2[k];
                    q[k] = a / (double)(l - m);
                }
            }
            double norm = (2.0 * l + 1.0) / (4.0 * kPi);
            for (int k = l - m + 1; k <= l + m; k++) norm /= (double)k;
            norm = c_sqrt(norm) * (m > 0 ? c_sqrt(2.0) : 1.0);
            for (int k = 0; k < kMaxDeg; k++) {
                t.coef[l][m][k] = (float)(norm * q[k]);
                q2[k] = q1[k];
                q1[k] = q[k];
            }
        }
    }
    return t;
}

__constant__ ShTable kSh = make_sh_table();

template <int DEG, bool WITH_GRAD>
__device__ __forceinline__ void sh_eval(float x, float y, float z, float *__restrict__ out, float *__restrict__ gx,
                                        float *__restrict__ gy, float *__restrict__ gz) {
    float c[DEG], s[DEG];
    c[0] = 1.0f; s[0] = 0.0f;
#pragma unroll
    for (int m = 1; m < DEG; m++) {
        c[m] = x * c[m - 1] - y * s[m - 1];
        s[m] = x * s[m - 1] + y * c[m - 1];
    }
#pragma unroll
    for (int l = 0; l < DEG; l++) {
#pragma unroll
        for (int m = 0; m <= l; m++) {
            // Horner over the l-m+1 coefficients (alternate ones are exactly zero and fold away)
            float q = kSh.coef[l][m][l - m], dq = 0.0f;
#pragma unroll
            for (int k = l - m - 1; k >= 0; k--) {
                if (WITH_GRAD) dq = fmaf(dq, z, q);
                q = fmaf(q, z, kSh.coef[l][m][k]);
            }
            const int base = l * l + l;
            if (m == 0) {
                out[base] = q;
                if (WITH_GRAD) { gx[base] = 0.0f; gy[base] = 0.0f; gz[base] = dq; }
            } else {
                out[base + m] = q * c[m];
                out[base - m] = q * s[m];
                if (WITH_GRAD) {
                    const float qm = q * (float)m;
                    gx[base + m] = qm * c[m - 1];
                    gy[base + m] = -qm * s[m - 1];
                    gz[base + m] = dq * c[m];
                    gx[base - m] = qm * s[m - 1];
                    gy[base - m] = qm * c[m - 1];
                    gz[base - m] = dq * s[m];
                }
            }
        }
    }
}

template <int DEG, bool WITH_GRAD>
__global__ void __launch_bounds__(128)
k_sh_forward(const float *__restrict__ inputs, float *__restrict__ outputs, uint32_t B, uint32_t D, float *__restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    constexpr int C2 = DEG * DEG;
    const float x = inputs[(size_t)b * D], y = inputs[(size_t)b * D + 1], z = inputs[(size_t)b * D + 2];
    float out[C2], gx[WITH_GRAD ? C2 : 1], gy[WITH_GRAD ? C2 : 1], gz[WITH_GRAD ? C2 : 1];
    sh_eval<DEG, WITH_GRAD>(x, y, z, out, gx, gy, gz);
    float *o = outputs + (size_t)b * C2;
    if constexpr (C2 % 4 == 0) {
#pragma unroll
        for (int i = 0; i < C2 / 4; i++) reinterpret_cast<float4 *>(o)[i] = make_float4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < C2; i++) o[i] = out[i];
    }
    if constexpr (WITH_GRAD) {
        float *g = dy_dx + (size_t)b * D * C2;  // [3, C2]: d/dx, d/dy, d/dz   (shencoder.cu:126-128)
#pragma unroll
        for (int i = 0; i < C2; i++) { g[i] = gx[i]; g[C2 + i] = gy[i]; g[2 * C2 + i] = gz[i]; }
    }
}

// grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]      (shencoder.cu:359-382)
__global__ void k_sh_backward(const float *__restrict__ grad, uint32_t B, uint32_t D, uint32_t C2,
                              const float *__restrict__ dy_dx, float *__restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float *g = grad + (size_t)b * C2, *j = dy_dx + ((size_t)b * D + d) * C2;
    float acc = grad_inputs[t];
    for (uint32_t ch = 0; ch < C2; ch++) acc = fmaf(g[ch], j[ch], acc);
    grad_inputs[t] = acc;
}

template <int DEG>
int launch_sh(const float *inputs, float *outputs, uint32_t B, uint32_t D, float *dy_dx, cudaStream_t st) {
    if (dy_dx) k_sh_forward<DEG, true><<<div_up(B, 128u), 128, 0, st>>>(inputs, outputs, B, D, dy_dx);
    else k_sh_forward<DEG, false><<<div_up(B, 128u), 128, 0, st>>>(inputs, outputs, B, D, nullptr);
    return (int)cudaPeekAtLastError();
}

// ---- frequency encoding (freqencoder.cu:30-94) ---------------------------------------------

// one thread per INPUT element (b, d): x is read once, the 2*deg+1 outputs of that element are written as runs of D
// consecutive floats per (sample, column block) across the warp.  scalbnf(x, f) of the reference is a multiplication by
// the exact power of two 2^f (same bits for every finite x that does not overflow); sin.approx like the reference's
// -use_fast_math build (freqencoder/backend.py:9), cos as sin(x + pi/2) (freqencoder.cu:56-58).
__global__ void __launch_bounds__(256)
k_freq_forward(const float *__restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *__restrict__ outputs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (uint64_t)b * D);
    const float x = __ldg(inputs + t);
    float *o = outputs + (size_t)b * C + d;
    o[0] = x;
    for (uint32_t f = 0; f < deg; f++) {
        const float xs = f < 127 ? __fmul_rn(x, __int_as_float((127 + f) << 23)) : scalbnf(x, (int)f);
        o[(size_t)(1 + 2 * f) * D] = __sinf(xs);
        o[(size_t)(2 + 2 * f) * D] = __sinf(__fadd_rn(xs, 1.5707963267948966f));
    }
}

__global__ void k_freq_backward(const float *__restrict__ grad, const float *__restrict__ outputs, uint32_t B, uint32_t D,
                                uint32_t deg, uint32_t C, float *__restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float *g = grad + (size_t)b * C, *o = outputs + (size_t)b * C;
    float r = g[d];
    g += D; o += D;
    for (uint32_t f = 0; f < deg; f++) {
        r += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
        g += 2 * D; o += 2 * D;
    }
    grad_inputs[t] = r;
}

}  // namespace

// C = degree (1..8); outputs [B, C*C]; dy_dx [B, 3*C*C] or NULL.   shencoder.h:9
S3D_API int s3d_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C, float *dy_dx,
                                  void *stream) {
    if (B == 0) return 0;
    if (D != 3) return S3D_EINVAL;
    cudaStream_t st = as_stream(stream);
    switch (C) {
        case 1: return launch_sh<1>(inputs, outputs, B, D, dy_dx, st);
        case 2: return launch_sh<2>(inputs, outputs, B, D, dy_dx, st);
        case 3: return launch_sh<3>(inputs, outputs, B, D, dy_dx, st);
        case 4: return launch_sh<4>(inputs, outputs, B, D, dy_dx, st);
        case 5: return launch_sh<5>(inputs, outputs, B, D, dy_dx, st);
        case 6: return launch_sh<6>(inputs, outputs, B, D, dy_dx, st);
        case 7: return launch_sh<7>(inputs, outputs, B, D, dy_dx, st);
        case 8: return launch_sh<8>(inputs, outputs, B, D, dy_dx, st);
        default: return S3D_EINVAL;
    }
}

// shencoder.h:10 (inputs is unused by the reference kernel as well)
S3D_API int s3d_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C,
                                   const float *dy_dx, float *grad_inputs, void *stream) {
    (void)inputs;
    if (B == 0) return 0;
    k_sh_backward<<<div_up(B * D, 256u), 256, 0, as_stream(stream)>>>(grad, B, D, C * C, dy_dx, grad_inputs);
    S3D_RETURN_LAST();
}

// freqencoder.h:7
S3D_API int s3d_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *outputs,
                                    void *stream) {
    if (B == 0) return 0;
    if (C != D + D * 2 * deg) return S3D_EINVAL;
    k_freq_forward<<<(unsigned)div_up((uint64_t)B * D, (uint64_t)256), 256, 0, as_stream(stream)>>>(inputs, B, D, deg, C, outputs);
    S3D_RETURN_LAST();
}

// freqencoder.h:10
S3D_API int s3d_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg,
                                     uint32_t C, float *grad_inputs, void *stream) {
    if (B == 0) return 0;
    k_freq_backward<<<div_up(B * D, 256u), 256, 0, as_stream(stream)>>>(grad, outputs, B, D, deg, C, grad_inputs);
    S3D_RETURN_LAST();
}
