// Sampler-level bandwidth-bound kernels: DDRM / GMM / generic updates with in-kernel Philox noise,
// uint8 quantisation for the host codec, colour L1 loss.  All images are NCHW fp32 like the reference's.
#include <stdarg.h>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void ddpmir_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* ddpmir_last_error(void) { return g_err; }
extern "C" int ddpmir_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (definition pinned in ddpmir.h / oracle/restated.py::philox_normal)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ float u23(uint32_t w) {  // ((w >> 9) + 0.5) * 2^-23, exact in fp32
    return (float)((w >> 9) * 2u + 1u) * 5.9604644775390625e-08f;  // (2k+1) * 2^-24
}

__device__ __forceinline__ float4 philox_normal4(uint64_t group, uint32_t step, uint64_t seed) {
    uint4 c = make_uint4((uint32_t)group, step, (uint32_t)(group >> 32), 0u);
    uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 w = philox4x32_10(c, k);
    float r0 = sqrtf(-2.f * logf(u23(w.x)));
    float r1 = sqrtf(-2.f * logf(u23(w.z)));
    float s0, c0, s1, c1;
    sincospif(2.f * u23(w.y), &s0, &c0);
    sincospif(2.f * u23(w.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__global__ void philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint32_t step) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t e = g * 4;
    if (e >= n) return;
    float4 z = philox_normal4((uint64_t)g, step, seed);
    if (e + 3 < n && ((uintptr_t)(out + e) & 15) == 0) {
        *reinterpret_cast<float4*>(out + e) = z;
    } else {
        float zz[4] = {z.x, z.y, z.z, z.w};
        for (int i = 0; i < 4 && e + i < n; ++i) out[e + i] = zz[i];
    }
}

extern "C" int ddpmir_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t step, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(out && n > 0, "philox_normal: bad arguments");
    int64_t groups = (n + 3) / 4;
    philox_normal_kernel<<<ceil_div(groups, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, step);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// DDRM update.  One thread = 4 consecutive NCHW elements (one Philox call).  Unfused fp32 arithmetic in the
// reference's order (webp_inference.py:584-592) so the result is bit-identical for injected noise.
// ---------------------------------------------------------------------------------------------------------
template <bool U8>
__global__ void __launch_bounds__(256)
ddrm_update_kernel(const float* __restrict__ x_theta, const void* __restrict__ codec, const float* __restrict__ y,
                   const float* __restrict__ z, const float* __restrict__ t, float* __restrict__ out,
                   int64_t per_image, int C, int HW, int B, float sigma_scale, float eta, float eta_b,
                   float one_minus_eta_b, int last_step, uint64_t seed, uint32_t step, uint64_t group_offset) {
    const int64_t groups_per_image = per_image / 4;
    const int64_t total_groups = groups_per_image * B;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total_groups;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = g * 4;
        const int b = (int)(e / per_image);
        const float4 xt = *reinterpret_cast<const float4*>(x_theta + e);
        const float4 yy = *reinterpret_cast<const float4*>(y + e);
        float cc[4];
        if (U8) {
            // element e -> (c, hw) inside image b; the decoder's pixels are [HW, C] uint8
            const int64_t r = e - (int64_t)b * per_image;
            const int c = (int)(r / HW);
            const int hw = (int)(r - (int64_t)c * HW);
            const uint8_t* p = reinterpret_cast<const uint8_t*>(codec) + (int64_t)b * per_image;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = __fdiv_rn((float)p[(int64_t)(hw + i) * C + c], 255.f);   // ToTensor
                cc[i] = __fmul_rn(__fsub_rn(v, 0.5f), 2.0f);                      // .sub(0.5).mul(2.0)
            }
        } else {
            const float4 c4 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(codec) + e);
            cc[0] = c4.x; cc[1] = c4.y; cc[2] = c4.z; cc[3] = c4.w;
        }
        const float xv[4] = {xt.x, xt.y, xt.z, xt.w};
        const float yv[4] = {yy.x, yy.y, yy.z, yy.w};
        float o[4];
        if (last_step) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __fadd_rn(__fsub_rn(xv[i], cc[i]), yv[i]);
        } else {
            float4 zz;
            if (z) zz = *reinterpret_cast<const float4*>(z + e);
            else zz = philox_normal4(group_offset + (uint64_t)g, step, seed);
            const float zv[4] = {zz.x, zz.y, zz.z, zz.w};
            const float ns = __fmul_rn(t[b], sigma_scale);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xp = __fadd_rn(__fsub_rn(xv[i], cc[i]), yv[i]);
                const float noise = __fmul_rn(zv[i], ns);
                const float mix = __fadd_rn(__fmul_rn(eta_b, xp), __fmul_rn(one_minus_eta_b, xv[i]));
                o[i] = __fadd_rn(mix, __fmul_rn(eta, noise));
            }
        }
        *reinterpret_cast<float4*>(out + e) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

extern "C" int ddpmir_ddrm_update(const float* x_theta, const void* codec, int codec_u8_hwc, const float* y,
                                  const float* z, const float* t, float* out, int B, int C, int H, int W,
                                  double sigma_scale_d, double eta_d, double eta_b_d, int last_step, uint64_t seed,
                                  uint32_t step, uint64_t noise_offset, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x_theta && codec && y && t && out, "ddrm_update: null pointer");
    DDPMIR_CHECK_ARG(noise_offset % 4 == 0, "ddrm_update: noise_offset must be a multiple of 4");
    DDPMIR_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "ddrm_update: bad shape");
    const int64_t HW = (int64_t)H * W;
    DDPMIR_CHECK_ARG(HW % 4 == 0, "ddrm_update: H*W must be a multiple of 4 (got %lld)", (long long)HW);
    const int64_t per_image = HW * C;
    const int64_t groups = per_image / 4 * B;
    int grid = (int)((groups + 255) / 256);
    const int cap = 148 * 16;
    if (grid > cap) grid = cap;
    // Python-side scalars are doubles; the reference's tensor ops round each of them to fp32 once
    const float sigma_scale = (float)sigma_scale_d, eta = (float)eta_d, eta_b = (float)eta_b_d;
    const float om = (float)(1.0 - eta_b_d);
    if (codec_u8_hwc)
        ddrm_update_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x_theta, codec, y, z, t, out, per_image, C,
                                                                       (int)HW, B, sigma_scale, eta, eta_b, om,
                                                                       last_step, seed, step, noise_offset / 4);
    else
        ddrm_update_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x_theta, codec, y, z, t, out, per_image, C,
                                                                        (int)HW, B, sigma_scale, eta, eta_b, om,
                                                                        last_step, seed, step, noise_offset / 4);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// GMM update (0409_method.ipynb#c1:L411-447), reference operation order
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gmm_update_kernel(const float* __restrict__ x_t, const float* __restrict__ pred, const float* __restrict__ y,
                  const float* __restrict__ prior, float g, float one_minus_g, const float* __restrict__ z,
                  float* __restrict__ out, int64_t n4, int use_first, float noise_scale, int last_step,
                  uint64_t seed, uint32_t step) {
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < n4; gi += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = gi * 4;
        const float4 x4 = *reinterpret_cast<const float4*>(x_t + e);
        const float4 p4 = *reinterpret_cast<const float4*>(pred + e);
        float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        float pv[4] = {p4.x, p4.y, p4.z, p4.w};
        if (prior) {
            const float4 y4 = *reinterpret_cast<const float4*>(y + e);
            const float4 s4 = *reinterpret_cast<const float4*>(prior + e);
            const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                pv[i] = __fadd_rn(__fmul_rn(one_minus_g, pv[i]), __fmul_rn(g, __fsub_rn(yv[i], sv[i])));
        }
        float o[4];
        if (last_step) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __fadd_rn(xv[i], pv[i]);
        } else {
            float4 zz;
            if (z) zz = *reinterpret_cast<const float4*>(z + e);
            else zz = philox_normal4((uint64_t)gi, step, seed);
            const float zv[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x0 = __fadd_rn(xv[i], pv[i]);
                float mean;
                if (use_first) mean = __fadd_rn(__fmul_rn(x0, 0.9f), __fmul_rn(xv[i], 0.1f));
                else mean = __fsub_rn(__fmul_rn(x0, 1.1f), __fmul_rn(xv[i], 0.1f));
                o[i] = __fadd_rn(mean, __fmul_rn(noise_scale, zv[i]));
            }
        }
        *reinterpret_cast<float4*>(out + e) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

extern "C" int ddpmir_gmm_update(const float* x_t, const float* pred, const float* y, const float* svd_prior,
                                 float g, const float* z, float* out, int64_t n, int use_first, float noise_scale,
                                 int last_step, uint64_t seed, uint32_t step, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x_t && pred && out && n > 0 && n % 4 == 0, "gmm_update: bad arguments");
    DDPMIR_CHECK_ARG(!svd_prior || y, "gmm_update: svd_prior needs y");
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    gmm_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_t, pred, y, svd_prior, g, (float)(1.0 - (double)g), z,
                                                             out, n / 4, use_first, noise_scale, last_step, seed, step);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// generic linear combination (classical DDPM mean etc.)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lincomb_kernel(const float* __restrict__ a, float wa, const float* __restrict__ b, float wb,
               const float* __restrict__ z, float sigma, float* __restrict__ out, int64_t n4, uint64_t seed,
               uint32_t step) {
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < n4; gi += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = gi * 4;
        const float4 a4 = *reinterpret_cast<const float4*>(a + e);
        float o[4] = {wa * a4.x, wa * a4.y, wa * a4.z, wa * a4.w};
        if (b) {
            const float4 b4 = *reinterpret_cast<const float4*>(b + e);
            o[0] = fmaf(wb, b4.x, o[0]); o[1] = fmaf(wb, b4.y, o[1]); o[2] = fmaf(wb, b4.z, o[2]); o[3] = fmaf(wb, b4.w, o[3]);
        }
        if (sigma != 0.f) {
            float4 zz;
            if (z) zz = *reinterpret_cast<const float4*>(z + e);
            else zz = philox_normal4((uint64_t)gi, step, seed);
            o[0] = fmaf(sigma, zz.x, o[0]); o[1] = fmaf(sigma, zz.y, o[1]); o[2] = fmaf(sigma, zz.z, o[2]); o[3] = fmaf(sigma, zz.w, o[3]);
        }
        *reinterpret_cast<float4*>(out + e) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

extern "C" int ddpmir_lincomb(const float* a, float wa, const float* b, float wb, const float* z, float sigma,
                              float* out, int64_t n, uint64_t seed, uint32_t step, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(a && out && n > 0 && n % 4 == 0, "lincomb: bad arguments");
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    lincomb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, wa, b, wb, z, sigma, out, n / 4, seed, step);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// uint8 quantisation + NCHW -> HWC for the host codec (webp_inference.py:509)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quantize_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int C, int HW, int64_t total_px) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_px; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = p / HW;
        const int64_t hw = p - b * HW;
        const float* src = x + b * (int64_t)C * HW + hw;
        uint8_t* dst = out + p * C;
        for (int c = 0; c < C; ++c) {
            float v = __fadd_rn(__fmul_rn(src[(int64_t)c * HW], 127.5f), 127.5f);
            v = fminf(fmaxf(v, 0.f), 255.f);
            dst[c] = (uint8_t)v;  // truncation toward zero, like .to(torch.uint8)
        }
    }
}

extern "C" int ddpmir_quantize_u8_hwc(const float* x, uint8_t* out, int B, int C, int H, int W, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && B > 0 && C > 0 && H > 0 && W > 0, "quantize_u8: bad arguments");
    const int64_t total = (int64_t)B * H * W;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    quantize_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, C, H * W, total);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

__global__ void __launch_bounds__(256)
dequantize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int C, int HW, int64_t total) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / ((int64_t)C * HW);
        const int64_t r = e - b * (int64_t)C * HW;
        const int c = (int)(r / HW);
        const int hw = (int)(r - (int64_t)c * HW);
        const float v = __fdiv_rn((float)in[(b * HW + hw) * C + c], 255.f);
        out[e] = __fmul_rn(__fsub_rn(v, 0.5f), 2.0f);
    }
}

extern "C" int ddpmir_u8_hwc_to_nchw(const uint8_t* in, float* out, int B, int C, int H, int W, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0, "u8_hwc_to_nchw: bad arguments");
    const int64_t total = (int64_t)B * C * H * W;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    dequantize_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, C, H * W, total);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// channel-weighted L1 on clamped [0,1] images (0409_method.ipynb#c0:L66-76)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
color_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, int HW, int64_t total4,
                double* __restrict__ acc) {
    // total4 = B*3*HW/4 float4 groups; channel of a group = (e / HW) % 3
    float s[3] = {0.f, 0.f, 0.f};
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < total4; gi += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = gi * 4;
        const int c = (int)((e / HW) % 3);
        const float4 p = *reinterpret_cast<const float4*>(pred + e);
        const float4 q = *reinterpret_cast<const float4*>(target + e);
        const float pv[4] = {p.x, p.y, p.z, p.w}, qv[4] = {q.x, q.y, q.z, q.w};
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = fminf(fmaxf(__fadd_rn(__fmul_rn(pv[i], 0.5f), 0.5f), 0.f), 1.f);
            const float b = fminf(fmaxf(__fadd_rn(__fmul_rn(qv[i], 0.5f), 0.5f), 0.f), 1.f);
            d += fabsf(a - b);
        }
        s[0] += (c == 0) ? d : 0.f;
        s[1] += (c == 1) ? d : 0.f;
        s[2] += (c == 2) ? d : 0.f;
    }
    __shared__ float sh[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = warp_sum(s[c]);
        if (lane == 0) sh[c][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += (double)sh[threadIdx.x][w];
        atomicAdd(&acc[threadIdx.x], v);
    }
}

__global__ void color_l1_finalize(const double* __restrict__ acc, double inv_count, float* __restrict__ out) {
    const float r = (float)(acc[0] * inv_count), g = (float)(acc[1] * inv_count), b = (float)(acc[2] * inv_count);
    out[0] = __fadd_rn(__fadd_rn(__fmul_rn(0.25f, r), __fmul_rn(0.5f, g)), __fmul_rn(0.25f, b));
}

extern "C" int ddpmir_color_l1(const float* pred, const float* target, int B, int H, int W, float* out_scalar,
                               double* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(pred && target && out_scalar && ws && B > 0, "color_l1: bad arguments");
    const int HW = H * W;
    DDPMIR_CHECK_ARG(HW % 4 == 0, "color_l1: H*W must be a multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws, 0, 3 * sizeof(double), st);
    const int64_t total4 = (int64_t)B * 3 * HW / 4;
    int grid = (int)((total4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    color_l1_kernel<<<grid, 256, 0, st>>>(pred, target, HW, total4, ws);
    color_l1_finalize<<<1, 1, 0, st>>>(ws, 1.0 / ((double)B * HW), out_scalar);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
extern "C" int ddpmir_cast_f32_to_bf16(const float* in, void* out, int64_t n, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(in && out && n > 0, "cast: bad arguments");
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (bf16*)out, n);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
