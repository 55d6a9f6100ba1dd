// GroupNorm (statistics + fused normalise/affine/activation), the 3-channel input convolution with norm1
// folded in, and the out_conv + tanh tail.  All HBM-bound, vectorised 8 channels per access on NHWC.
#include "common.cuh"

namespace {

// ---- statistics -----------------------------------------------------------------------------------------
// NHWC: grid (chunks, B); 256 threads; thread owns channel vector v = tid % cv and strides over pixels.
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_nhwc_kernel(const T* __restrict__ x, int HW, int C, int G, int px_per_cta, double* __restrict__ ws) {
    const int cv = C >> 3;              // channel vectors per pixel
    const int ppi = 256 / cv;           // pixels per iteration (cv divides 256)
    const int b = blockIdx.y;
    const int v = threadIdx.x % cv;
    const int pl = threadIdx.x / cv;
    const int p0 = blockIdx.x * px_per_cta;
    const int p1 = min(HW, p0 + px_per_cta);
    const T* base = x + (long long)b * HW * C + v * 8;
    float s = 0.f, ss = 0.f;
    for (int p = p0 + pl; p < p1; p += ppi) {
        Vec8<T> a;
        a.load(base + (long long)p * C);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += a.v[i]; ss = fmaf(a.v[i], a.v[i], ss); }
    }
    __shared__ float sh[2][64];  // G <= 64
    if (threadIdx.x < 64) { sh[0][threadIdx.x] = 0.f; sh[1][threadIdx.x] = 0.f; }
    __syncthreads();
    const int g = (v * 8) / (C / G);
    atomicAdd(&sh[0][g], s);
    atomicAdd(&sh[1][g], ss);
    __syncthreads();
    if (threadIdx.x < G) {
        atomicAdd(&ws[((long long)b * G + threadIdx.x) * 2 + 0], (double)sh[0][threadIdx.x]);
        atomicAdd(&ws[((long long)b * G + threadIdx.x) * 2 + 1], (double)sh[1][threadIdx.x]);
    }
}

// NCHW fp32: group g of image b is the contiguous range [(b*G+g)*n, +n), n = (C/G)*HW.
__global__ void __launch_bounds__(256)
gn_stats_nchw_kernel(const float* __restrict__ x, long long n, int chunk, double* __restrict__ ws) {
    const long long bg = blockIdx.y;
    const long long e0 = (long long)blockIdx.x * chunk;
    const long long e1 = min(n, e0 + (long long)chunk);
    const float* base = x + bg * n;
    float s = 0.f, ss = 0.f;
    for (long long e = e0 + threadIdx.x; e < e1; e += 256) {
        const float a = base[e];
        s += a; ss = fmaf(a, a, ss);
    }
    s = warp_sum(s); ss = warp_sum(ss);
    __shared__ float sh[2][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wid] = s; sh[1][wid] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
        atomicAdd(&ws[bg * 2 + 0], a);
        atomicAdd(&ws[bg * 2 + 1], b);
    }
}

__global__ void gn_finalize_kernel(const double* __restrict__ ws, int count, double inv_n, float eps,
                                   float* __restrict__ mean_rstd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double mean = ws[2 * i] * inv_n;
    double var = ws[2 * i + 1] * inv_n - mean * mean;
    if (var < 0) var = 0;
    mean_rstd[2 * i] = (float)mean;
    mean_rstd[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---- apply ----------------------------------------------------------------------------------------------
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const T* __restrict__ x, TO* __restrict__ out, TO* __restrict__ raw_copy, long long total_vec, int HW,
                int C, int G, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                const float* __restrict__ beta, int act) {
    const int cv = C >> 3;
    const int cpg = C / G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        const long long px = i / cv;
        const int b = (int)(px / HW);
        const int c0 = v * 8;
        const int g = c0 / cpg;
        const float mean = mean_rstd[((long long)b * G + g) * 2], rstd = mean_rstd[((long long)b * G + g) * 2 + 1];
        Vec8<T> a;
        a.load(x + i * 8);
        Vec8<TO> o;
        if (raw_copy) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = a.v[k];
            o.store(raw_copy + i * 8);
        }
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + c0), g1 = *reinterpret_cast<const float4*>(gamma + c0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + c0), b1 = *reinterpret_cast<const float4*>(beta + c0 + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = act_apply(act, fmaf((a.v[k] - mean) * rstd, gg[k], bb[k]));
        o.store(out + i * 8);
    }
}

// ---- input conv (Cin <= 4, NCHW fp32 in, NHWC out), optional GroupNorm fold -----------------------------
template <typename T, int KS>
__global__ void __launch_bounds__(256)
conv_input_kernel(const float* __restrict__ x, int B, int Cin, int H, int W, const float* __restrict__ mean_rstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ w,
                  const float* __restrict__ bias, const float* __restrict__ row_bias, int N, T* __restrict__ out) {
    // smem weights as [tap][c][n] for broadcast-free reads: thread owns 16 consecutive n
    extern __shared__ float sw[];  // KS*KS*Cin*N
    const int KK = KS * KS;
    for (int i = threadIdx.x; i < KK * Cin * N; i += blockDim.x) {
        const int n = i % N, r = i / N;
        const int c = r % Cin, tap = r / Cin;
        sw[i] = w[((long long)n * Cin + c) * KK + tap];
    }
    __syncthreads();
    const int ngrp = N / 16;
    const long long total = (long long)B * H * W * ngrp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ng = (int)(i % ngrp);
        const long long px = i / ngrp;
        const int b = (int)(px / (H * W));
        const int rem = (int)(px - (long long)b * H * W);
        const int h = rem / W, ww = rem - h * W;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int n = ng * 16 + j;
            acc[j] = (bias ? bias[n] : 0.f) + (row_bias ? row_bias[(long long)b * N + n] : 0.f);
        }
        for (int c = 0; c < Cin; ++c) {
            float sc = 1.f, sh = 0.f;
            if (mean_rstd) {  // one group per channel for the 3-channel input (G == Cin)
                const float mean = mean_rstd[((long long)b * Cin + c) * 2], rstd = mean_rstd[((long long)b * Cin + c) * 2 + 1];
                sc = rstd * gamma[c];
                sh = beta[c] - mean * sc;
            }
            const float* plane = x + ((long long)b * Cin + c) * H * W;
#pragma unroll
            for (int tap = 0; tap < KK; ++tap) {
                const int hh = h + (KS == 3 ? tap / 3 - 1 : 0), w2 = ww + (KS == 3 ? tap % 3 - 1 : 0);
                if (hh < 0 || hh >= H || w2 < 0 || w2 >= W) continue;
                const float v = fmaf(plane[(long long)hh * W + w2], sc, sh);
                const float* wp = sw + ((long long)tap * Cin + c) * N + ng * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
            }
        }
        T* o = out + px * N + ng * 16;
        Vec8<T> v0, v1;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v0.v[j] = acc[j]; v1.v[j] = acc[8 + j]; }
        v0.store(o); v1.store(o + 8);
    }
}

// ---- out_conv + tanh: NHWC in, NCHW fp32 out, N <= 4 ----------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
out_conv_tanh_kernel(const T* __restrict__ x, int B, int H, int W, int Cin, const float* __restrict__ w,
                     const float* __restrict__ bias, int N, float* __restrict__ out) {
    extern __shared__ float sw[];  // [n][tap][c]
    for (int i = threadIdx.x; i < N * 9 * Cin; i += blockDim.x) {
        const int c = i % Cin, r = i / Cin;
        const int tap = r % 9, n = r / 9;
        sw[i] = w[((long long)n * Cin + c) * 9 + tap];
    }
    __syncthreads();
    // eight lanes per pixel, each lane 8 of every 64 channels: a warp reads four pixels' 128-byte channel rows per load (one
    // full line each) instead of 32 lines with 16 bytes used -- the thread-per-pixel version was bound by LSU wavefronts
    // (0.6 ms per 16 x 256 x 256 call for 134 MB); partial sums are folded with three shuffles
    const long long total = (long long)B * H * W;
    const int cl = threadIdx.x & 7;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 3);
    // the loop runs over the warp's first pixel so that all 32 lanes stay together (the shuffles below are warp-wide);
    // out-of-range pixels of the last warp recompute the last pixel and do not store
    for (long long base = (long long)blockIdx.x * (blockDim.x >> 3) + ((threadIdx.x >> 5) << 2); base < total; base += stride) {
        const long long px0 = base + ((threadIdx.x & 31) >> 3);
        const bool in = px0 < total;
        const long long px = in ? px0 : total - 1;
        const int b = (int)(px / (H * W));
        const int rem = (int)(px - (long long)b * H * W);
        const int h = rem / W, ww = rem - h * W;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int tap = 0; tap < 9; ++tap) {
            const int hh = h + tap / 3 - 1, w2 = ww + tap % 3 - 1;
            if (hh < 0 || hh >= H || w2 < 0 || w2 >= W) continue;
            const T* src = x + (((long long)b * H + hh) * W + w2) * Cin;
            for (int c = cl * 8; c < Cin; c += 64) {
                Vec8<T> a;
                a.load(src + c);
                for (int n = 0; n < N; ++n) {
                    const float* wp = sw + ((long long)n * 9 + tap) * Cin + c;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[n] = fmaf(a.v[k], wp[k], acc[n]);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 1);
            acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 2);
            acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], 4);
        }
        if (in && cl == 0)
            for (int n = 0; n < N; ++n)
                out[((long long)b * N + n) * H * W + rem] = tanhf(acc[n] + bias[n]);
    }
}

}  // namespace

extern "C" int ddpmir_groupnorm_stats(const void* x, int dtype, int nchw, int B, int HW, int C, int G, float eps,
                                      float* mean_rstd, double* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && mean_rstd && ws, "groupnorm_stats: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, "groupnorm_stats: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
    if (nchw) {
        DDPMIR_CHECK_ARG(dtype == DDPMIR_F32, "groupnorm_stats: NCHW input must be fp32");
        const long long n = (long long)(C / G) * HW;
        const int chunk = 1 << 14;
        dim3 grid(ceil_div(n, chunk), B * G);
        gn_stats_nchw_kernel<<<grid, 256, 0, st>>>((const float*)x, n, chunk, ws);
    } else {
        const int cv = C / 8;
        DDPMIR_CHECK_ARG(C % 8 == 0 && cv <= 256 && 256 % cv == 0 && (C / G) % 8 == 0 && G <= 64,
                         "groupnorm_stats: unsupported C=%d G=%d for NHWC", C, G);
        // aim for >= 4 waves of CTAs, each CTA at least 32 pixels
        int ppc = HW;
        const int target_ctas = 148 * 4;
        int chunks = (target_ctas + B - 1) / B;
        if (chunks < 1) chunks = 1;
        ppc = (HW + chunks - 1) / chunks;
        if (ppc < 32) ppc = 32;
        dim3 grid(ceil_div(HW, ppc), B);
        if (dtype == DDPMIR_F32) gn_stats_nhwc_kernel<float><<<grid, 256, 0, st>>>((const float*)x, HW, C, G, ppc, ws);
        else gn_stats_nhwc_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, HW, C, G, ppc, ws);
    }
    DDPMIR_LAUNCH_CHECK();
    gn_finalize_kernel<<<ceil_div(B * G, 128), 128, 0, st>>>(ws, B * G, 1.0 / ((double)(C / G) * HW), eps, mean_rstd);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_groupnorm_apply(const void* x, int dtype, int B, int HW, int C, int G, const float* mean_rstd,
                                      const float* gamma, const float* beta, int act, void* out, int out_dtype,
                                      void* raw_copy, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && mean_rstd && gamma && beta, "groupnorm_apply: null pointer");
    DDPMIR_CHECK_ARG(C % 8 == 0 && C % G == 0 && (C / G) % 8 == 0, "groupnorm_apply: unsupported C=%d G=%d", C, G);
    const long long total = (long long)B * HW * (C / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
#define GO(TI, TO) gn_apply_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)x, (TO*)out, (TO*)raw_copy, total, HW, C, G, mean_rstd, gamma, beta, act)
    if (dtype == DDPMIR_F32) { if (out_dtype == DDPMIR_F32) GO(float, float); else GO(float, bf16); }
    else { if (out_dtype == DDPMIR_F32) GO(bf16, float); else GO(bf16, bf16); }
#undef GO
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_conv_input(const float* x, int B, int Cin, int H, int W, const float* mean_rstd,
                                 const float* gamma, const float* beta, const float* w, const float* bias,
                                 const float* row_bias, int N, int ksize, int dtype, void* out,
                                 ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && w && out, "conv_input: null pointer");
    DDPMIR_CHECK_ARG(Cin >= 1 && Cin <= 4 && N % 16 == 0 && (ksize == 1 || ksize == 3), "conv_input: unsupported shape");
    DDPMIR_CHECK_ARG(!mean_rstd || (gamma && beta), "conv_input: norm fold needs gamma/beta");
    const size_t smem = sizeof(float) * ksize * ksize * Cin * N;
    DDPMIR_CHECK_ARG(smem <= 48 * 1024, "conv_input: weights do not fit shared memory");
    const long long total = (long long)B * H * W * (N / 16);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(T, KS) conv_input_kernel<T, KS><<<grid, 256, smem, st>>>(x, B, Cin, H, W, mean_rstd, gamma, beta, w, bias, row_bias, N, (T*)out)
    if (dtype == DDPMIR_F32) { if (ksize == 3) LAUNCH(float, 3); else LAUNCH(float, 1); }
    else { if (ksize == 3) LAUNCH(bf16, 3); else LAUNCH(bf16, 1); }
#undef LAUNCH
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_out_conv_tanh(const void* x, int dtype, int B, int H, int W, int Cin, const float* w,
                                    const float* bias, int N, float* out, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && w && bias && out, "out_conv_tanh: null pointer");
    DDPMIR_CHECK_ARG(N >= 1 && N <= 4 && Cin % 8 == 0, "out_conv_tanh: unsupported shape");
    const size_t smem = sizeof(float) * N * 9 * Cin;
    DDPMIR_CHECK_ARG(smem <= 48 * 1024, "out_conv_tanh: weights do not fit shared memory");
    const long long total = (long long)B * H * W;
    int grid = (int)((total + 15) / 16);            // 16 pixels per 128-thread CTA and iteration (eight lanes per pixel)
    if (grid > 148 * 16) grid = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_F32) out_conv_tanh_kernel<float><<<grid, 128, smem, st>>>((const float*)x, B, H, W, Cin, w, bias, N, out);
    else out_conv_tanh_kernel<bf16><<<grid, 128, smem, st>>>((const bf16*)x, B, H, W, Cin, w, bias, N, out);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
