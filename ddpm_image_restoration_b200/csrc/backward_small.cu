// Backward of the small fp32 layers of the training step: row-wise linears (time embedding, time_proj), element-wise
// activations, the 3-channel input convolution with its folded GroupNorm, and the out_conv + tanh tail.
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_grad_s(int act, float u) {
    switch (act) {
        case DDPMIR_ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752440f));
            return cdf + u * 0.39894228040143267794f * expf(-0.5f * u * u);
        }
        case DDPMIR_ACT_SILU: { const float s = 1.f / (1.f + expf(-u)); return s * (1.f + u * (1.f - s)); }
        case DDPMIR_ACT_RELU: return u > 0.f ? 1.f : 0.f;
        case DDPMIR_ACT_LRELU02: return u > 0.f ? 1.f : 0.2f;
        case DDPMIR_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-u)); return s * (1.f - s); }
        case DDPMIR_ACT_TANH: { const float t = tanhf(u); return 1.f - t * t; }
        default: return 1.f;
    }
}

__global__ void act_fwd_kernel(const float* __restrict__ x, int act, float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = act_apply(act, x[i]);
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ u, int act, float* __restrict__ dx, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = dy[i] * act_grad_s(act, u[i]);
}

// dx[r,k] = sum_n dy[r,n] W[n,k]
__global__ void linear_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int R, int K, int N,
                                 int accumulate) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (k >= K) return;
    float s = accumulate ? dx[(long long)r * K + k] : 0.f;
    for (int n = 0; n < N; ++n) s = fmaf(dy[(long long)r * N + n], w[(long long)n * K + k], s);
    dx[(long long)r * K + k] = s;
}
// dW[n,k] += sum_r dy[r,n] x[r,k];  db[n] += sum_r dy[r,n]
__global__ void linear_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                 float* __restrict__ db, int R, int K, int N) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (k >= K) return;
    float s = 0.f, sb = 0.f;
    for (int r = 0; r < R; ++r) { const float g = dy[(long long)r * N + n]; s = fmaf(g, x[(long long)r * K + k], s); sb += g; }
    dw[(long long)n * K + k] += s;
    if (db && k == 0) db[n] += sb;
}

// ---- input conv (Cin <= 4) backward: weight/bias gradients + gradients of the folded GroupNorm affine ---------------------
// dW[n,c,tap] += sum dh[b,h,w,n] * gn(x)[b,c,h+dh,w+dw];  bias handled by ddpmir_colsum.
template <int KS>
__global__ void __launch_bounds__(256)
conv_input_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dh, int B, int Cin, int H, int W, int N,
                        const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                        float* __restrict__ dw, int px_per_cta) {
    constexpr int KK = KS * KS;
    const long long total = (long long)B * H * W;
    const long long p0 = (long long)blockIdx.x * px_per_cta, p1 = min(total, p0 + px_per_cta);
    // thread (n lane = tid % 64, pixel lane = tid / 64); N handled per 64-wide slab.  ONE pass over the CTA's pixels with all
    // Cin * KK accumulators in registers: dh[p, n] is read once (coalesced over n), the <= 4 x 9 normalised input samples around p
    // are warp-uniform loads.  (The previous version made one pass per (c, tap): 27 passes, 1.2 ms at 32 x 64 x 64.)
    const int nl = threadIdx.x & 63, pl = threadIdx.x >> 6;
    __shared__ float red[4][64];
    const int hw = H * W;
    for (int n0 = 0; n0 < N; n0 += 64) {
        const int n = n0 + nl;
        float acc[4][KK];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int t = 0; t < KK; ++t) acc[c][t] = 0.f;
        if (n < N)
            for (long long p = p0 + pl; p < p1; p += 4) {
                const int b = (int)(p / hw);
                const int rem = (int)(p - (long long)b * hw);
                const int h0 = rem / W, w0 = rem - h0 * W;
                const float g = dh[p * N + n];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c >= Cin) break;
                    float sc = 1.f, sh = 0.f;
                    if (mean_rstd) {      // gn(x) = (x - mean) * rstd * gamma + beta = x * sc + sh
                        const float mean = mean_rstd[((long long)b * Cin + c) * 2], rstd = mean_rstd[((long long)b * Cin + c) * 2 + 1];
                        sc = rstd * gamma[c]; sh = fmaf(-mean, sc, beta[c]);
                    }
                    const float* xp = x + ((long long)b * Cin + c) * hw;
#pragma unroll
                    for (int t = 0; t < KK; ++t) {
                        const int h = h0 + (KS == 3 ? t / 3 - 1 : 0), w = w0 + (KS == 3 ? t % 3 - 1 : 0);
                        if (h < 0 || h >= H || w < 0 || w >= W) continue;
                        acc[c][t] = fmaf(g, fmaf(xp[h * W + w], sc, sh), acc[c][t]);
                    }
                }
            }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= Cin) break;
#pragma unroll
            for (int t = 0; t < KK; ++t) {
                red[pl][nl] = acc[c][t];
                __syncthreads();
                if (pl == 0 && n < N) atomicAdd(&dw[((long long)n * Cin + c) * KK + t], red[0][nl] + red[1][nl] + red[2][nl] + red[3][nl]);
                __syncthreads();
            }
        }
    }
}

// gradient wrt the normalised 3-channel input, reduced on the fly to dgamma[c] = sum g*xhat, dbeta[c] = sum g
__global__ void __launch_bounds__(256)
conv_input_affine_grad_kernel(const float* __restrict__ x, const float* __restrict__ dh, int B, int Cin, int H, int W, int N,
                              const float* __restrict__ w, const float* __restrict__ mean_rstd, float* __restrict__ dgamma,
                              float* __restrict__ dbeta) {
    extern __shared__ float sw[];   // [tap][c][n]
    for (int i = threadIdx.x; i < 9 * Cin * N; i += blockDim.x) {
        const int n = i % N, r = i / N;
        const int c = r % Cin, tap = r / Cin;
        sw[i] = w[((long long)n * Cin + c) * 9 + tap];
    }
    __syncthreads();
    float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
    const long long total = (long long)B * H * W;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(p / (H * W));
        const int rem = (int)(p - (long long)b * H * W);
        const int h = rem / W, ww = rem - h * W;
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        for (int tap = 0; tap < 9; ++tap) {
            // output pixel (h - dh, w - dw) read input pixel (h, w) through this tap
            const int ho = h - (tap / 3 - 1), wo = ww - (tap % 3 - 1);
            if (ho < 0 || ho >= H || wo < 0 || wo >= W) continue;
            const float* d = dh + (((long long)b * H + ho) * W + wo) * N;
            for (int n = 0; n < N; ++n) {
                const float dv = d[n];
                for (int c = 0; c < Cin; ++c) g[c] = fmaf(dv, sw[((long long)tap * Cin + c) * N + n], g[c]);
            }
        }
        for (int c = 0; c < Cin; ++c) {
            const float mean = mean_rstd[((long long)b * Cin + c) * 2], rstd = mean_rstd[((long long)b * Cin + c) * 2 + 1];
            const float xh = (x[(((long long)b * Cin + c) * H + h) * W + ww] - mean) * rstd;
            dg[c] = fmaf(g[c], xh, dg[c]); db[c] += g[c];
        }
    }
    for (int c = 0; c < Cin; ++c) {
        const float a = warp_sum(dg[c]), bsum = warp_sum(db[c]);
        if ((threadIdx.x & 31) == 0) { atomicAdd(&dgamma[c], a); atomicAdd(&dbeta[c], bsum); }
    }
}

// ---- out_conv + tanh backward ---------------------------------------------------------------------------------------
// dz = dy * (1 - y^2) (NCHW, N <= 4);  da[b,h,w,c] = sum_{n,tap} dz[b,n,h-dh,w-dw] W[n,c,tap];  dW, dbias accumulated.
__global__ void __launch_bounds__(256)
out_conv_bwd_data_kernel(const float* __restrict__ y, const float* __restrict__ dy, int B, int H, int W, int Cin, int N,
                         const float* __restrict__ w, float* __restrict__ da) {
    extern __shared__ float sw[];   // [n][tap][c]
    for (int i = threadIdx.x; i < N * 9 * Cin; i += blockDim.x) {
        const int c = i % Cin, r = i / Cin;
        const int tap = r % 9, n = r / 9;
        sw[i] = w[((long long)n * Cin + c) * 9 + tap];
    }
    __syncthreads();
    const long long total = (long long)B * H * W * Cin;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        const long long p = i / Cin;
        const int b = (int)(p / (H * W));
        const int rem = (int)(p - (long long)b * H * W);
        const int h = rem / W, ww = rem - h * W;
        float s = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int ho = h - (tap / 3 - 1), wo = ww - (tap % 3 - 1);
            if (ho < 0 || ho >= H || wo < 0 || wo >= W) continue;
            for (int n = 0; n < N; ++n) {
                const long long j = (((long long)b * N + n) * H + ho) * W + wo;
                const float yv = y[j];
                s = fmaf(dy[j] * (1.f - yv * yv), sw[((long long)n * 9 + tap) * Cin + c], s);
            }
        }
        da[i] = s;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
out_conv_bwd_weight_kernel(const T* __restrict__ a, const float* __restrict__ y, const float* __restrict__ dy, int B, int H, int W,
                           int Cin, int N, float* __restrict__ dw, float* __restrict__ dbias, int px_per_cta) {
    // thread = (c lane 0..63, pixel lane 0..3).  ONE pass over the CTA's output pixels with all N * 9 accumulators in registers:
    // dz[n] = dy (1 - y^2) at the pixel is warp-uniform, the nine shifted activation rows are coalesced over c.  (The previous
    // version made one pass per (n, tap): 27 passes, 1.3 ms at 32 x 64 x 64.)
    const int cl = threadIdx.x & 63, pl = threadIdx.x >> 6;
    const long long total = (long long)B * H * W;
    const long long p0 = (long long)blockIdx.x * px_per_cta, p1 = min(total, p0 + px_per_cta);
    __shared__ float red[4][64];
    const int hw = H * W;
    for (int c0 = 0; c0 < Cin; c0 += 64) {
        const int c = c0 + cl;
        float acc[4][9];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[n][t] = 0.f;
        if (c < Cin)
            for (long long p = p0 + pl; p < p1; p += 4) {
                const int b = (int)(p / hw);
                const int rem = (int)(p - (long long)b * hw);
                const int h = rem / W, ww = rem - h * W;
                float dz[4];
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    dz[n] = 0.f;
                    if (n < N) {
                        const long long j = ((long long)b * N + n) * hw + rem;
                        const float yv = y[j];
                        dz[n] = dy[j] * (1.f - yv * yv);
                    }
                }
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int hi = h + t / 3 - 1, wi = ww + t % 3 - 1;
                    if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                    const float av = to_f(a[(((long long)b * H + hi) * W + wi) * Cin + c]);
#pragma unroll
                    for (int n = 0; n < 4; ++n) acc[n][t] = fmaf(dz[n], av, acc[n][t]);
                }
            }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (n >= N) break;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                red[pl][cl] = acc[n][t];
                __syncthreads();
                if (pl == 0 && c < Cin) atomicAdd(&dw[((long long)n * Cin + c) * 9 + t], red[0][cl] + red[1][cl] + red[2][cl] + red[3][cl]);
                __syncthreads();
            }
        }
    }
    if (blockIdx.x * blockDim.x + threadIdx.x < N) { /* bias handled below by the first CTAs' threads */ }
    // bias gradient: every CTA reduces its own pixel range
    for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (long long p = p0 + threadIdx.x; p < p1; p += 256) {
            const int b = (int)(p / (H * W));
            const int rem = (int)(p - (long long)b * H * W);
            const long long j = ((long long)b * N + n) * H * W + rem;
            const float yv = y[j];
            s += dy[j] * (1.f - yv * yv);
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) atomicAdd(&dbias[n], s);
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148 * 16;
    return (int)(g > cap ? cap : g);
}

}  // namespace

extern "C" int ddpmir_act_forward(const float* x, int act, float* out, int64_t n, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && n > 0, "act_forward: bad arguments");
    act_fwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, act, out, n);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
extern "C" int ddpmir_act_backward(const float* dy, const float* u, int act, float* dx, int64_t n, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dy && u && dx && n > 0, "act_backward: bad arguments");
    act_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, u, act, dx, n);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_linear_rows_backward(const float* dy, const float* x, const float* w, int rows, int K, int N, float* dx,
                                           int accumulate_dx, float* dw, float* db, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dy && x && w && rows > 0 && K > 0 && N > 0 && rows <= 65535 && N <= 65535, "linear_rows_backward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) linear_dx_kernel<<<dim3(ceil_div(K, 128), rows), 128, 0, st>>>(dy, w, dx, rows, K, N, accumulate_dx);
    if (dw) linear_dw_kernel<<<dim3(ceil_div(K, 128), N), 128, 0, st>>>(dy, x, dw, db, rows, K, N);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_conv_input_backward(const float* x, const float* dh, int B, int Cin, int H, int W, int N, int ksize,
                                          const float* w, const float* mean_rstd, const float* gamma, const float* beta, float* dw,
                                          float* dgamma, float* dbeta, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && dh && dw && Cin >= 1 && Cin <= 4 && (ksize == 1 || ksize == 3), "conv_input_backward: bad arguments");
    DDPMIR_CHECK_ARG(!mean_rstd || (gamma && beta && dgamma && dbeta && w && ksize == 3), "conv_input_backward: norm fold needs its tensors");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)B * H * W;
    int ppc = (int)((total + 148 * 2 - 1) / (148 * 2));
    if (ppc < 64) ppc = 64;
    if (ksize == 3) conv_input_wgrad_kernel<3><<<ceil_div(total, ppc), 256, 0, st>>>(x, dh, B, Cin, H, W, N, mean_rstd, gamma, beta, dw, ppc);
    else conv_input_wgrad_kernel<1><<<ceil_div(total, ppc), 256, 0, st>>>(x, dh, B, Cin, H, W, N, mean_rstd, gamma, beta, dw, ppc);
    DDPMIR_LAUNCH_CHECK();
    if (mean_rstd) {
        const size_t smem = sizeof(float) * 9 * Cin * N;
        DDPMIR_CHECK_ARG(smem <= 48 * 1024, "conv_input_backward: weights do not fit shared memory");
        conv_input_affine_grad_kernel<<<grid_for(total, 256), 256, smem, st>>>(x, dh, B, Cin, H, W, N, w, mean_rstd, dgamma, dbeta);
        DDPMIR_LAUNCH_CHECK();
    }
    return DDPMIR_OK;
}

extern "C" int ddpmir_out_conv_tanh_backward(const void* a, int dtype, const float* y, const float* dy, int B, int H, int W, int Cin,
                                             int N, const float* w, float* da, float* dw, float* dbias, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(a && y && dy && w && da && dw && dbias && N >= 1 && N <= 4, "out_conv_tanh_backward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * N * 9 * Cin;
    DDPMIR_CHECK_ARG(smem <= 48 * 1024, "out_conv_tanh_backward: weights do not fit shared memory");
    out_conv_bwd_data_kernel<<<grid_for((long long)B * H * W * Cin, 256), 256, smem, st>>>(y, dy, B, H, W, Cin, N, w, da);
    DDPMIR_LAUNCH_CHECK();
    const long long total = (long long)B * H * W;
    int ppc = (int)((total + 148 * 2 - 1) / (148 * 2));
    if (ppc < 64) ppc = 64;
    if (dtype == DDPMIR_F32) out_conv_bwd_weight_kernel<float><<<ceil_div(total, ppc), 256, 0, st>>>((const float*)a, y, dy, B, H, W, Cin, N, dw, dbias, ppc);
    else out_conv_bwd_weight_kernel<bf16><<<ceil_div(total, ppc), 256, 0, st>>>((const bf16*)a, y, dy, B, H, W, Cin, N, dw, dbias, ppc);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
