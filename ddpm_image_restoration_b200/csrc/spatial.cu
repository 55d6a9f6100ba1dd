// Coalesced NHWC spatial kernels: blockwise per-channel transforms (DCT / learned AVIF transform), max-pool,
// bilinear up-sample fused with the skip concat, the adaptive average-pool pyramid and the AVIF gate combine.
#include "common.cuh"

namespace {

// ---- blockwise transform  out = alpha*x + beta*(T X T^T) -------------------------------------------------
// CTA = 64 channels (threadIdx.x) x 4 spatial blocks (threadIdx.y).  A warp touches 32 consecutive channels of
// one pixel per access.  The per-channel matrices are staged transposed in shared memory ([elem][channel]).
template <typename T, typename TO, int BS>
__global__ void __launch_bounds__(256)
block_transform_kernel(const T* __restrict__ x, TO* __restrict__ out, int B, int H, int W, int C,
                       const float* __restrict__ Tm, int per_channel, float alpha, float beta) {
    __shared__ float Ts[BS * BS][64];
    const int c = blockIdx.y * 64 + threadIdx.x;
    const int tid = threadIdx.y * 64 + threadIdx.x;
    for (int i = tid; i < BS * BS * 64; i += 256) {
        const int ch = i & 63, e = i >> 6;
        const int cc = blockIdx.y * 64 + ch;
        Ts[e][ch] = per_channel ? (cc < C ? Tm[(long long)cc * BS * BS + e] : 0.f) : Tm[e];
    }
    __syncthreads();
    const int nbh = (H + BS - 1) / BS, nbw = (W + BS - 1) / BS;
    const long long nblk = (long long)B * nbh * nbw;
    const long long blk = (long long)blockIdx.x * 4 + threadIdx.y;
    if (blk >= nblk || c >= C) return;
    const int b = (int)(blk / (nbh * nbw));
    const int r = (int)(blk - (long long)b * nbh * nbw);
    const int h0 = (r / nbw) * BS, w0 = (r % nbw) * BS;
    const T* src = x + (long long)b * H * W * C + c;
    float X[BS][BS];
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
        for (int j = 0; j < BS; ++j) {
            const int h = h0 + i, w = w0 + j;
            X[i][j] = (h < H && w < W) ? to_f(src[((long long)h * W + w) * C]) : 0.f;  // zero padding (F.pad)
        }
    // Y = T X  (rows), then Z = Y T^T
    float Y[BS][BS];
#pragma unroll
    for (int u = 0; u < BS; ++u)
#pragma unroll
        for (int j = 0; j < BS; ++j) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < BS; ++i) a = fmaf(Ts[u * BS + i][threadIdx.x], X[i][j], a);
            Y[u][j] = a;
        }
    TO* dst = out + (long long)b * H * W * C + c;
#pragma unroll
    for (int u = 0; u < BS; ++u)
#pragma unroll
        for (int v = 0; v < BS; ++v) {
            const int h = h0 + u, w = w0 + v;
            if (h < H && w < W) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < BS; ++j) a = fmaf(Y[u][j], Ts[v * BS + j][threadIdx.x], a);
                dst[((long long)h * W + w) * C] = from_f<TO>(alpha * X[u][v] + beta * a);
            }
        }
}

// ---- MaxPool2d(2) ----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C) {
    const int cv = C >> 3, Ho = H >> 1, Wo = W >> 1;
    const long long total = (long long)B * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long p = i / cv;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        const T* s = x + (((long long)b * H + 2 * ho) * W + 2 * wo) * C + v * 8;
        Vec8<T> a, b2, c2, d;
        a.load(s); b2.load(s + C); c2.load(s + (long long)W * C); d.load(s + (long long)W * C + C);
#pragma unroll
        for (int k = 0; k < 8; ++k) a.v[k] = fmaxf(fmaxf(a.v[k], b2.v[k]), fmaxf(c2.v[k], d.v[k]));
        a.store(out + i * 8);
    }
}

// ---- bilinear x2 (align_corners=False) fused with the channel concat -------------------------------------
__device__ __forceinline__ void up2_src(int d, int n, int& i0, int& i1, float& lam) {
    // src = (d + 0.5) / 2 - 0.5, clamped at 0 (PyTorch area_pixel_compute_source_index, align_corners=False)
    float s = (d + 0.5f) * 0.5f - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + (i0 < n - 1 ? 1 : 0);
    lam = s - (float)i0;
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2_concat_kernel(const T* __restrict__ lo, const T* __restrict__ skip, T* __restrict__ out, int B, int H, int W,
                        int C1, int C2) {
    const int Ct = C1 + C2, cv = Ct >> 3, cv1 = C1 >> 3;
    const int Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * Ho * Wo * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long p = i / cv;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        Vec8<T> r;
        if (v < cv1) {
            int h0, h1, w0, w1; float lh, lw;
            up2_src(ho, H, h0, h1, lh);
            up2_src(wo, W, w0, w1, lw);
            const T* base = lo + (long long)b * H * W * C1 + v * 8;
            Vec8<T> a00, a01, a10, a11;
            a00.load(base + ((long long)h0 * W + w0) * C1);
            a01.load(base + ((long long)h0 * W + w1) * C1);
            a10.load(base + ((long long)h1 * W + w0) * C1);
            a11.load(base + ((long long)h1 * W + w1) * C1);
            const float hl0 = 1.f - lh, wl0 = 1.f - lw;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                r.v[k] = hl0 * (wl0 * a00.v[k] + lw * a01.v[k]) + lh * (wl0 * a10.v[k] + lw * a11.v[k]);
        } else {
            r.load(skip + (((long long)b * Ho + ho) * Wo + wo) * C2 + (v - cv1) * 8);
        }
        r.store(out + i * 8);
    }
}

// ---- adaptive average pool pyramid (s = 1, 2, 4, 8) ------------------------------------------------------
// generic adaptive pooling: cell i of s covers [floor(i*n/s), ceil((i+1)*n/s)).  One CTA per (cell, b);
// threads stride channels (coalesced) and split the window rows between y-lanes.
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_pyramid_kernel(const T* __restrict__ x, float* __restrict__ out, int B, int H, int W, int C, int cell_lo) {
    const int cell = cell_lo + blockIdx.x, b = blockIdx.y;
    int s, ci, cj;
    pyramid_cell(cell, s, ci, cj);
    const int h0 = (ci * H) / s, h1 = ((ci + 1) * H + s - 1) / s;
    const int w0 = (cj * W) / s, w1 = ((cj + 1) * W + s - 1) / s;
    const int npx = (h1 - h0) * (w1 - w0);
    const int ww = w1 - w0;
    __shared__ float red[256];
    // thread (tx = channel lane 0..63, ty = pixel lane 0..3)
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int c0 = 0; c0 < C; c0 += 64) {
        const int c = c0 + tx;
        float a = 0.f;
        if (c < C)
            for (int p = ty; p < npx; p += 4) {
                const int h = h0 + p / ww, w = w0 + p % ww;
                a += to_f(x[(((long long)b * H + h) * W + w) * C + c]);
            }
        red[threadIdx.x] = a;
        __syncthreads();
        if (ty == 0 && c < C)
            out[((long long)cell * B + b) * C + c] = (red[tx] + red[64 + tx] + red[128 + tx] + red[192 + tx]) / (float)npx;
        __syncthreads();
    }
}

// coarse levels from the s=8 level when H, W are multiples of 8 (each coarse cell = mean of equal-size cells)
__global__ void avgpool_coarsen_kernel(float* __restrict__ out, int C, int B) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (c >= C) return;
    float* o = out + (long long)b * C + c;
    const long long CS = (long long)B * C;  // cell stride
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float a = 0.f;
            for (int u = 0; u < 2; ++u) for (int v = 0; v < 2; ++v) a += o[(long long)(21 + (2 * i + u) * 8 + 2 * j + v) * CS];
            o[(long long)(5 + i * 4 + j) * CS] = a * 0.25f;
        }
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
            float a = 0.f;
            for (int u = 0; u < 2; ++u) for (int v = 0; v < 2; ++v) a += o[(long long)(5 + (2 * i + u) * 4 + 2 * j + v) * CS];
            o[(long long)(1 + i * 2 + j) * CS] = a * 0.25f;
        }
    float a = 0.f;
    for (int k = 0; k < 4; ++k) a += o[(long long)(1 + k) * CS];
    o[0] = a * 0.25f;
}

// ---- AVIF combine ----------------------------------------------------------------------------------------
template <typename TH, typename T>
__global__ void __launch_bounds__(256)
avif_combine_kernel(const TH* __restrict__ hsrc, const T* __restrict__ xt, const float* __restrict__ gates,
                    const T* __restrict__ color, const T* __restrict__ edge, T* __restrict__ out, int B, int H, int W,
                    int C) {
    const int cv = C >> 3;
    const long long total = (long long)B * H * W * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long p = i / cv;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        const float* gb = gates + (long long)b * C + v * 8;
        const long long CS = (long long)B * C;  // gates are [85, B, C]
        float attn[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) attn[k] = gb[k];  // s = 1: constant map
        const int base[3] = {1, 5, 21};
#pragma unroll
        for (int li = 0; li < 3; ++li) {
            const int s = 2 << li;
            if (s == H && s == W) {  // same shape: the reference skips interpolate
                const float* g = gb + (long long)(base[li] + h * s + w) * CS;
#pragma unroll
                for (int k = 0; k < 8; ++k) attn[k] += g[k];
                continue;
            }
            int h0, h1, w0, w1; float lh, lw;
            bilin_src(h, s, H, h0, h1, lh);
            bilin_src(w, s, W, w0, w1, lw);
            const float* g00 = gb + (long long)(base[li] + h0 * s + w0) * CS;
            const float* g01 = gb + (long long)(base[li] + h0 * s + w1) * CS;
            const float* g10 = gb + (long long)(base[li] + h1 * s + w0) * CS;
            const float* g11 = gb + (long long)(base[li] + h1 * s + w1) * CS;
            const float hl0 = 1.f - lh, wl0 = 1.f - lw;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                attn[k] += hl0 * (wl0 * g00[k] + lw * g01[k]) + lh * (wl0 * g10[k] + lw * g11[k]);
        }
        Vec8<TH> hv;
        Vec8<T> xv, cvv, ev;
        hv.load(hsrc + i * 8); xv.load(xt + i * 8); cvv.load(color + i * 8); ev.load(edge + i * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) xv.v[k] = hv.v[k] + xv.v[k] * (attn[k] * 0.25f) * cvv.v[k] * ev.v[k];
        xv.store(out + i * 8);
    }
}

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148 * 16;
    return (int)(g > cap ? cap : g);
}

}  // namespace

extern "C" int ddpmir_block_transform(const void* x, int dtype, int B, int H, int W, int C, const float* T, int bs,
                                      int per_channel, float alpha, float beta, void* out, int out_dtype,
                                      ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && T, "block_transform: null pointer");
    DDPMIR_CHECK_ARG(bs == 4 || bs == 8, "block_transform: block size %d", bs);
    const long long nblk = (long long)B * ((H + bs - 1) / bs) * ((W + bs - 1) / bs);
    dim3 grid(ceil_div(nblk, 4), ceil_div(C, 64)), block(64, 4);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(TT, TO, BS) block_transform_kernel<TT, TO, BS><<<grid, block, 0, st>>>((const TT*)x, (TO*)out, B, H, W, C, T, per_channel, alpha, beta)
#define LAUNCH_BS(TT, TO) do { if (bs == 4) LAUNCH(TT, TO, 4); else LAUNCH(TT, TO, 8); } while (0)
    if (dtype == DDPMIR_F32) { if (out_dtype == DDPMIR_F32) LAUNCH_BS(float, float); else LAUNCH_BS(float, bf16); }
    else { if (out_dtype == DDPMIR_F32) LAUNCH_BS(bf16, float); else LAUNCH_BS(bf16, bf16); }
#undef LAUNCH_BS
#undef LAUNCH
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_maxpool2(const void* x, int dtype, int B, int H, int W, int C, void* out, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2: bad arguments");
    const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_F32) maxpool2_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)x, (float*)out, B, H, W, C);
    else maxpool2_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)x, (bf16*)out, B, H, W, C);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_upsample2_concat(const void* lo, const void* skip, int dtype, int B, int H, int W, int C1,
                                       int C2, void* out, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(lo && skip && out && C1 % 8 == 0 && C2 % 8 == 0, "upsample2_concat: bad arguments");
    const long long total = (long long)B * 4 * H * W * ((C1 + C2) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_F32)
        upsample2_concat_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)lo, (const float*)skip, (float*)out, B, H, W, C1, C2);
    else
        upsample2_concat_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)lo, (const bf16*)skip, (bf16*)out, B, H, W, C1, C2);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_avgpool_pyramid(const void* x, int dtype, int B, int H, int W, int C, float* out,
                                      ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && C > 0, "avgpool_pyramid: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool uniform = (H % 8 == 0) && (W % 8 == 0);
    const int cell_lo = uniform ? 21 : 0;
    dim3 grid(85 - cell_lo, B);
    if (dtype == DDPMIR_F32) avgpool_pyramid_kernel<float><<<grid, 256, 0, st>>>((const float*)x, out, B, H, W, C, cell_lo);
    else avgpool_pyramid_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, out, B, H, W, C, cell_lo);
    DDPMIR_LAUNCH_CHECK();
    if (uniform) {
        avgpool_coarsen_kernel<<<dim3(ceil_div(C, 128), B), 128, 0, st>>>(out, C, B);
        DDPMIR_LAUNCH_CHECK();
    }
    return DDPMIR_OK;
}

// out = x + y * s[b, c]  (FrequencyAwareBlock of the 0409 UNet, 0409_method.ipynb#c0:L256-263: x + x_freq * attn with a
// per-image, per-channel squeeze-excite gate); x and y fp32 NHWC, optional second copy of the result in the operand dtype
template <typename T2>
__global__ void __launch_bounds__(256)
channel_scale_add_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ s, float* __restrict__ out,
                         T2* __restrict__ out2, long long HW, int C, long long total) {
    const int cv = C >> 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        const long long b = i / cv / HW;
        const float* sp = s + b * C + v * 8;
        Vec8<float> xv, yv;
        xv.load(x + i * 8); yv.load(y + i * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) xv.v[k] = fmaf(yv.v[k], sp[k], xv.v[k]);
        xv.store(out + i * 8);
        if (out2) {
            Vec8<T2> ov;
#pragma unroll
            for (int k = 0; k < 8; ++k) ov.v[k] = xv.v[k];
            ov.store(out2 + i * 8);
        }
    }
}

extern "C" int ddpmir_avif_combine(const void* h, int h_dtype, const void* xt, const float* gates, const void* color,
                                   const void* edge, int dtype, int B, int H, int W, int C, void* out,
                                   ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(h && xt && gates && color && edge && out && C % 8 == 0, "avif_combine: bad arguments");
    const long long total = (long long)B * H * W * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
#define GO(TH, T) avif_combine_kernel<TH, T><<<grid_for(total, 256), 256, 0, st>>>((const TH*)h, (const T*)xt, gates, (const T*)color, (const T*)edge, (T*)out, B, H, W, C)
    if (dtype == DDPMIR_F32) { DDPMIR_CHECK_ARG(h_dtype == DDPMIR_F32, "avif_combine: fp32 mode needs fp32 h"); GO(float, float); }
    else { if (h_dtype == DDPMIR_F32) GO(float, bf16); else GO(bf16, bf16); }
#undef GO
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_channel_scale_add(const float* x, const float* y, const float* s, int B, long long HW, int C, float* out,
                                        void* out2, int out2_dtype, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && y && s && out && B > 0 && HW > 0 && C > 0 && C % 8 == 0, "channel_scale_add: bad arguments");
    const long long total = (long long)B * HW * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (out2 && out2_dtype == DDPMIR_BF16)
        channel_scale_add_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>(x, y, s, out, (bf16*)out2, HW, C, total);
    else
        channel_scale_add_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(x, y, s, out, (float*)out2, HW, C, total);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
