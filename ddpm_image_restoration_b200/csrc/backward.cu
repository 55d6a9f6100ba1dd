// Backward kernels of the training step (webp_training.py:476-537): weight gradients of the GEMM-shaped layers, bias /
// per-image column sums, GroupNorm backward, the element-wise backward pieces of the frequency gate and dropout,
// max-pool / up-sample backward.  Gradients live in fp32 (the "stream" class of DESIGN.md); activation operands may be
// bf16.  Data gradients of convolutions reuse the forward implicit-GEMM kernels with transposed, tap-flipped weights.
#include "epilogue.cuh"

namespace {

// ---- weight gradient:  dW[n, k] += sum_m dY[m, n] * im2col(X)[m, k]  (k = tap*Cin + c) ------------------------------
// Output sub-block [n_begin, n_begin+n_count) x [k_begin, k_begin+k_count) is written to out with leading dimension
// out_ld; layout 0 = row-major [n][k - k_begin], layout 1 = OIHW for 3x3 convs: ((n*Cin + c)*9 + tap).
// grid: (k tiles, n tiles, M splits); 64x64 output tile per CTA, 16 pixels per step; atomicAdd across splits.
__global__ void __launch_bounds__(256)
wgrad_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype, float* __restrict__ out,
             long long M, int H, int W, int Cin, int N, int taps, int n_begin, int n_count, int k_begin, int k_count,
             int out_ld, int layout, long long m_per_split) {
    __shared__ float Ys[16][64 + 1];
    __shared__ float Xs[16][64 + 1];
    const int tid = threadIdx.x;
    const int k0 = k_begin + blockIdx.x * 64, n0 = n_begin + blockIdx.y * 64;
    const int k_end = k_begin + k_count, n_end = n_begin + n_count;
    const long long m_lo = (long long)blockIdx.z * m_per_split;
    const long long m_hi = min(M, m_lo + m_per_split);
    const int ty = tid >> 4, tx = tid & 15;   // 4x4 outputs: n = n0 + ty*4.., k = k0 + tx*4..
    float acc[4][4] = {};
    const int lp = tid >> 4;       // loader: pixel row 0..15
    const int lc = (tid & 15) * 4; // 4 consecutive columns
    const int hw = H * W;
    for (long long mb = m_lo; mb < m_hi; mb += 16) {
        const long long m = mb + lp;
        float yv[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < m_hi) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + lc + j;
                if (n < n_end) yv[j] = ld_any(dy, dy_dtype, m * N + n);
            }
            const int b = (int)(m / hw);
            const int rem = (int)(m - (long long)b * hw);
            const int h = rem / W, w = rem - h * W;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + lc + j;
                if (k < k_end) {
                    const int tap = k / Cin, c = k - tap * Cin;
                    int hh = h, ww = w;
                    if (taps == 9) { hh += tap / 3 - 1; ww += tap % 3 - 1; }
                    if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                        xv[j] = ld_any(x, x_dtype, (((long long)b * H + hh) * W + ww) * Cin + c);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) { Ys[lp][lc + j] = yv[j]; Xs[lp][lc + j] = xv[j]; }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ys[p][ty * 4 + i]; bb[i] = Xs[p][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= n_end) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= k_end) continue;
            long long idx;
            if (layout == 1) { const int tap = k / Cin, c = k - tap * Cin; idx = ((long long)(n - n_begin) * Cin + c) * 9 + tap; }
            else idx = (long long)(n - n_begin) * out_ld + (k - k_begin);
            atomicAdd(&out[idx], acc[i][j]);
        }
    }
}

// ---- column sums of dY: bias gradients, per-image row-bias gradients, low/high class-split bias gradients -----------
// out_total[n] += sum_m sel(m) dY[m,n];  out_img[b,n] += the same per image.  cls: -1 all rows, 0 high-frequency rows
// only, 1 low-frequency rows only (is_low_freq of the pixel).
// Threads = CW column lanes (power of two >= n_count, <= 256) x 256 / CW row lanes, four independent partial sums per thread: a
// 64-column gradient keeps 16 rows in flight per CTA instead of one (the one-row-at-a-time version was latency-bound: 4.2 ms of
// a 35 ms training step).  Row lanes are folded in shared memory, then one atomic per (column, CTA, image).
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ dy, int dtype, long long M, int N, int H, int W, int cls, int bs, int low, int n_begin,
              int n_count, float* __restrict__ out_total, float* __restrict__ out_img, int rows_per_cta) {
    __shared__ float red[256];
    const int hw = H * W;
    const long long m0 = (long long)blockIdx.x * rows_per_cta;
    const long long m1 = min(M, m0 + rows_per_cta);
    dy = dtype == DDPMIR_F32 ? (const void*)((const float*)dy + n_begin) : (const void*)((const bf16*)dy + n_begin);
    const int ldn = N;
    int CW = 1;
    while (CW < n_count && CW < 256) CW <<= 1;
    const int RL = 256 / CW;
    const int cn = threadIdx.x & (CW - 1), rl = threadIdx.x / CW;
    auto take = [&](long long m, int b) -> bool {
        if (cls < 0) return true;
        const int rem = (int)(m - (long long)b * hw);
        const int h = rem / W, w = rem - h * W;
        return (int)is_low_freq(h, w, H, W, bs, low) == cls;
    };
    for (long long seg = m0; seg < m1;) {                      // one segment per image touched by this CTA (usually one)
        const int b = (int)(seg / hw);
        const long long seg_end = min(m1, (long long)(b + 1) * hw);
        for (int nb = 0; nb < n_count; nb += CW) {
            const int n = nb + cn;
            float s = 0.f;
            if (n < n_count) {
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                long long m = seg + rl;
                for (; m + 3 * RL < seg_end; m += 4 * RL) {
                    const float v0 = ld_any(dy, dtype, m * ldn + n), v1 = ld_any(dy, dtype, (m + RL) * ldn + n);
                    const float v2 = ld_any(dy, dtype, (m + 2 * RL) * ldn + n), v3 = ld_any(dy, dtype, (m + 3 * RL) * ldn + n);
                    s0 += take(m, b) ? v0 : 0.f; s1 += take(m + RL, b) ? v1 : 0.f;
                    s2 += take(m + 2 * RL, b) ? v2 : 0.f; s3 += take(m + 3 * RL, b) ? v3 : 0.f;
                }
                for (; m < seg_end; m += RL)
                    if (take(m, b)) s0 += ld_any(dy, dtype, m * ldn + n);
                s = (s0 + s1) + (s2 + s3);
            }
            if (RL > 1) {
                red[threadIdx.x] = s;
                __syncthreads();
                if (rl == 0)
                    for (int l = 1; l < RL; ++l) s += red[l * CW + cn];
                __syncthreads();
            }
            if (rl == 0 && n < n_count) {
                if (out_img) atomicAdd(&out_img[(long long)b * n_count + n], s);
                if (out_total) atomicAdd(&out_total[n], s);
            }
        }
        seg = seg_end;
    }
}

// ---- GroupNorm backward --------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(int act, float u) {
    switch (act) {
        case DDPMIR_ACT_GELU: {
            const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752440f));
            const float pdf = 0.39894228040143267794f * expf(-0.5f * u * u);
            return cdf + u * pdf;
        }
        case DDPMIR_ACT_SILU: {
            const float s = 1.f / (1.f + expf(-u));
            return s * (1.f + u * (1.f - s));
        }
        case DDPMIR_ACT_RELU: return u > 0.f ? 1.f : 0.f;
        case DDPMIR_ACT_LRELU02: return u > 0.f ? 1.f : 0.2f;
        case DDPMIR_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-u)); return s * (1.f - s); }
        case DDPMIR_ACT_TANH: { const float t = tanhf(u); return 1.f - t * t; }
        default: return 1.f;
    }
}

// pass 1: per (b, group) sums of dxhat and dxhat*xhat (double), per-channel dgamma / dbeta (float atomics)
__global__ void __launch_bounds__(256)
gn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, int HW, int C, int G, int act,
                    const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    double* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, int px_per_cta) {
    // threads = CW channel lanes (power of two >= C, <= 256) x 256 / CW pixel lanes: a 64-channel tensor keeps four pixel rows in
    // flight per CTA instead of one (same restructuring as colsum_kernel); pixel lanes are folded in shared memory
    __shared__ float red[4][256];
    const int b = blockIdx.y;
    const int cpg = C / G;
    const int p0 = blockIdx.x * px_per_cta, p1 = min(HW, p0 + px_per_cta);
    int CW = 1;
    while (CW < C && CW < 256) CW <<= 1;
    const int RL = 256 / CW;
    const int cn = threadIdx.x & (CW - 1), rl = threadIdx.x / CW;
    for (int c0 = 0; c0 < C; c0 += CW) {
        const int c = c0 + cn;
        const bool ok = c < C;
        const int g = ok ? c / cpg : 0;
        float s1 = 0.f, s2 = 0.f, dg = 0.f, db = 0.f;
        if (ok) {
            const float mean = mean_rstd[((long long)b * G + g) * 2], rstd = mean_rstd[((long long)b * G + g) * 2 + 1];
            const float ga = gamma[c], be = beta[c];
            for (int p = p0 + rl; p < p1; p += RL) {
                const long long i = ((long long)b * HW + p) * C + c;
                const float xh = (x[i] - mean) * rstd;
                const float dyh = dy[i] * act_grad(act, fmaf(ga, xh, be));
                dg = fmaf(dyh, xh, dg); db += dyh;
                const float dxh = dyh * ga;
                s1 += dxh; s2 = fmaf(dxh, xh, s2);
            }
        }
        if (RL > 1) {
            red[0][threadIdx.x] = s1; red[1][threadIdx.x] = s2; red[2][threadIdx.x] = dg; red[3][threadIdx.x] = db;
            __syncthreads();
            if (rl == 0)
                for (int l = 1; l < RL; ++l) {
                    s1 += red[0][l * CW + cn]; s2 += red[1][l * CW + cn]; dg += red[2][l * CW + cn]; db += red[3][l * CW + cn];
                }
            __syncthreads();
        }
        if (rl == 0 && ok) {
            atomicAdd(&dgamma[c], dg);
            atomicAdd(&dbeta[c], db);
            atomicAdd(&ws[((long long)b * G + g) * 2], (double)s1);
            atomicAdd(&ws[((long long)b * G + g) * 2 + 1], (double)s2);
        }
    }
}

// pass 2: dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat))
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long long total,
                    int HW, int C, int G, int act, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const double* __restrict__ ws, int accumulate) {
    const int cpg = C / G;
    const double inv_n = 1.0 / ((double)cpg * HW);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int b = (int)(i / ((long long)HW * C));
        const int g = c / cpg;
        const float mean = mean_rstd[((long long)b * G + g) * 2], rstd = mean_rstd[((long long)b * G + g) * 2 + 1];
        const float xh = (x[i] - mean) * rstd;
        const float dxh = dy[i] * act_grad(act, fmaf(gamma[c], xh, beta[c])) * gamma[c];
        const float m1 = (float)(ws[((long long)b * G + g) * 2] * inv_n), m2 = (float)(ws[((long long)b * G + g) * 2 + 1] * inv_n);
        const float v = rstd * (dxh - m1 - xh * m2);
        dx[i] = accumulate ? dx[i] + v : v;
    }
}

// ---- frequency gate backward (element-wise) ----------------------------------------------------------------------------
// forward: e = h3 + g * s * d, g = sigmoid(z), s = is_low ? 1 : boost[b].  Given de: dz = de*s*d*g*(1-g),
// dd = de*g*s (written), dh3 handled by the caller (= de).
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ de, const void* __restrict__ g, const void* __restrict__ d, int op_dtype,
                const float* __restrict__ boost, float* __restrict__ dz, float* __restrict__ dd, long long total, int H,
                int W, int C, int bs, int low) {
    const int hw = H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / C;
        const int b = (int)(m / hw);
        const int rem = (int)(m - (long long)b * hw);
        const int h = rem / W, w = rem - h * W;
        const float s = is_low_freq(h, w, H, W, bs, low) ? 1.f : boost[b];
        const float gv = ld_any(g, op_dtype, i), dv = ld_any(d, op_dtype, i), e = de[i];
        dz[i] = e * s * dv * gv * (1.f - gv);
        dd[i] = e * gv * s;
    }
}

// hidden layer of the stacked gate MLP: g1 = mask_class(lrelu(pre)); dpre = dg1 * mask * lrelu'(pre), sign(pre) = sign(g1)
__global__ void __launch_bounds__(256)
lrelu_mask_bwd_kernel(const float* __restrict__ dg1, const void* __restrict__ g1, int op_dtype, float* __restrict__ dpre,
                      long long total, int H, int W, int N, int bs, int low) {
    const int hw = H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / N;
        const int n = (int)(i - m * N);
        const int rem = (int)(m % hw);
        const int h = rem / W, w = rem - h * W;
        const bool lowp = is_low_freq(h, w, H, W, bs, low);
        const float v = ld_any(g1, op_dtype, i);
        float r = 0.f;
        if ((n < (N >> 1)) == lowp) r = dg1[i] * (v > 0.f ? 1.f : 0.2f);
        dpre[i] = r;
    }
}

// ---- dropout (Philox-free counter hash; same mask in forward and backward) -----------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (uint32_t)x;
}
__global__ void __launch_bounds__(256)
dropout_kernel(const void* __restrict__ in, int in_dtype, void* __restrict__ out, int out_dtype, long long total, float p,
               uint64_t seed) {
    const float keep_scale = 1.f / (1.f - p);
    const uint32_t thr = (uint32_t)((double)p * 4294967296.0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const bool keep = hash32(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)i) >= thr;
        st_any(out, out_dtype, i, keep ? ld_any(in, in_dtype, i) * keep_scale : 0.f);
    }
}

// ---- max-pool / up-sample backward -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C) {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)B * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long p = i / C;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        const long long base = (((long long)b * H + 2 * ho) * W + 2 * wo) * C + c;
        const long long o[4] = {base, base + C, base + (long long)W * C, base + (long long)W * C + C};
        int arg = 0; float best = x[o[0]];
#pragma unroll
        for (int k = 1; k < 4; ++k) { const float v = x[o[k]]; if (v > best) { best = v; arg = k; } }
        const float g = dy[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) dx[o[k]] = (k == arg) ? g : 0.f;
    }
}

__device__ __forceinline__ void up2_taps(int h, int H, int (&r)[4], float (&w)[4]) {
    // hi rows that read lo row h in F.interpolate(scale_factor=2, bilinear, align_corners=False), with their weights
    r[0] = 2 * h - 1; w[0] = h >= 1 ? 0.25f : 0.f;
    r[1] = 2 * h;     w[1] = h >= 1 ? 0.75f : 1.f;
    r[2] = 2 * h + 1; w[2] = h < H - 1 ? 0.75f : 1.f;
    r[3] = 2 * h + 2; w[3] = h < H - 1 ? 0.25f : 0.f;
}
__global__ void __launch_bounds__(256)
upsample2_concat_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dlo, float* __restrict__ dskip, int B, int H,
                            int W, int C1, int C2) {
    const int Ct = C1 + C2, Ho = 2 * H, Wo = 2 * W;
    const long long n_lo = (long long)B * H * W * C1, n_sk = (long long)B * Ho * Wo * C2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_lo + n_sk; i += (long long)gridDim.x * blockDim.x) {
        if (i < n_lo) {
            const int c = (int)(i % C1);
            long long p = i / C1;
            const int w = (int)(p % W); p /= W;
            const int h = (int)(p % H);
            const int b = (int)(p / H);
            int rr[4], cc[4]; float wr[4], wc[4];
            up2_taps(h, H, rr, wr); up2_taps(w, W, cc, wc);
            float s = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (wr[a] == 0.f) continue;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (wc[q] == 0.f) continue;
                    s = fmaf(wr[a] * wc[q], dy[(((long long)b * Ho + rr[a]) * Wo + cc[q]) * Ct + c], s);
                }
            }
            dlo[i] = s;
        } else {
            const long long j = i - n_lo;
            const int c = (int)(j % C2);
            const long long px = j / C2;
            dskip[j] = dy[px * Ct + C1 + c];
        }
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148 * 16;
    return (int)(g > cap ? cap : g);
}

}  // namespace

int ddpmir_wgrad_mma(const void* dy, const void* x, float* out, int B, int H, int W, int Cin, int N, int taps, int n_begin, int n_count,
                     int k_begin, int k_count, int out_ld, int oihw, cudaStream_t st);

extern "C" int ddpmir_wgrad(const void* dy, int dy_dtype, const void* x, int x_dtype, float* out, int B, int H, int W, int Cin,
                            int N, int taps, int n_begin, int n_count, int k_begin, int k_count, int out_ld, int oihw,
                            ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dy && x && out, "wgrad: null pointer");
    DDPMIR_CHECK_ARG((taps == 1 || taps == 9) && n_count > 0 && k_count > 0 && n_begin >= 0 && k_begin >= 0 &&
                     n_begin + n_count <= N && k_begin + k_count <= taps * Cin, "wgrad: bad sub-block");
    if (dy_dtype == DDPMIR_BF16 && x_dtype == DDPMIR_BF16) {      // tensor-core kernel when both operands are bf16
        const int rc = ddpmir_wgrad_mma(dy, x, out, B, H, W, Cin, N, taps, n_begin, n_count, k_begin, k_count, out_ld, oihw,
                                        (cudaStream_t)stream);
        if (rc != DDPMIR_ERR_UNSUPPORTED) return rc;
    }
    const long long M = (long long)B * H * W;
    const int tiles = ceil_div(k_count, 64) * ceil_div(n_count, 64);
    int splits = (148 * 4 + tiles - 1) / tiles;
    const long long max_splits = (M + 63) / 64;
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    long long per = (M + splits - 1) / splits;
    per = (per + 15) / 16 * 16;
    splits = (int)((M + per - 1) / per);
    dim3 grid(ceil_div(k_count, 64), ceil_div(n_count, 64), splits);
    wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, dy_dtype, x, x_dtype, out, M, H, W, Cin, N, taps, n_begin, n_count,
                                                        k_begin, k_count, out_ld, oihw, per);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_colsum(const void* dy, int dtype, int B, int H, int W, int N, int cls, int bs, int low, int n_begin,
                             int n_count, float* out_total, float* out_img, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dy && (out_total || out_img), "colsum: null pointer");
    DDPMIR_CHECK_ARG(n_begin >= 0 && n_count > 0 && n_begin + n_count <= N, "colsum: bad column range");
    DDPMIR_CHECK_ARG(cls < 0 || (bs > 0 && low > 0), "colsum: class split needs bs/low");
    const long long M = (long long)B * H * W;
    int rows = (int)((M + 148 * 4 - 1) / (148 * 4));
    if (rows < 8) rows = 8;
    colsum_kernel<<<ceil_div(M, rows), 256, 0, (cudaStream_t)stream>>>(dy, dtype, M, N, H, W, cls, bs, low, n_begin, n_count, out_total,
                                                                     out_img, rows);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_groupnorm_backward(const float* x, const float* dy, int B, int HW, int C, int G, int act,
                                         const float* mean_rstd, const float* gamma, const float* beta, float* dx,
                                         int accumulate_dx, float* dgamma, float* dbeta, double* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && dy && mean_rstd && gamma && beta && dx && dgamma && dbeta && ws, "groupnorm_backward: null pointer");
    DDPMIR_CHECK_ARG(C % G == 0, "groupnorm_backward: bad groups");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws, 0, sizeof(double) * 2 * B * G, st);
    int chunks = (148 * 4 + B - 1) / B;
    int ppc = (HW + chunks - 1) / chunks;
    if (ppc < 4) ppc = 4;
    gn_bwd_stats_kernel<<<dim3(ceil_div(HW, ppc), B), 256, 0, st>>>(x, dy, HW, C, G, act, mean_rstd, gamma, beta, ws, dgamma, dbeta, ppc);
    DDPMIR_LAUNCH_CHECK();
    const long long total = (long long)B * HW * C;
    gn_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, dy, dx, total, HW, C, G, act, mean_rstd, gamma, beta, ws, accumulate_dx);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_gate_backward(const float* de, const void* g, const void* d, int op_dtype, const float* boost, float* dz,
                                    float* dd, int B, int H, int W, int C, int bs, int low, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(de && g && d && boost && dz && dd, "gate_backward: null pointer");
    const long long total = (long long)B * H * W * C;
    gate_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(de, g, d, op_dtype, boost, dz, dd, total, H, W, C, bs, low);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_lrelu_mask_backward(const float* dg1, const void* g1, int op_dtype, float* dpre, int B, int H, int W, int N,
                                          int bs, int low, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dg1 && g1 && dpre, "lrelu_mask_backward: null pointer");
    const long long total = (long long)B * H * W * N;
    lrelu_mask_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dg1, g1, op_dtype, dpre, total, H, W, N, bs, low);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_dropout(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, float p, uint64_t seed,
                              ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(in && out && n > 0 && p >= 0.f && p < 1.f, "dropout: bad arguments");
    dropout_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, in_dtype, out, out_dtype, n, p, seed);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_maxpool2_backward(const float* x, const float* dy, float* dx, int B, int H, int W, int C, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && dy && dx && H % 2 == 0 && W % 2 == 0, "maxpool2_backward: bad arguments");
    const long long total = (long long)B * (H / 2) * (W / 2) * C;
    maxpool2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, B, H, W, C);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_upsample2_concat_backward(const float* dy, float* dlo, float* dskip, int B, int H, int W, int C1, int C2,
                                                ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dy && dlo && dskip, "upsample2_concat_backward: null pointer");
    const long long total = (long long)B * H * W * C1 + (long long)B * 4 * H * W * C2;
    upsample2_concat_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, dlo, dskip, B, H, W, C1, C2);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
