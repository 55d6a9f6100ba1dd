// Backward kernels of the AVIF family's frequency block (AVIFFreqAwareBlock / AVIFAdaptiveTransform, avif.py:185-321) for the
// training step train_epoch_ddrm_avif (avif.py:528-590).  Forward (spatial.cu, conv_tc.cu epilogues):
//     tr    = T_c X T_c^T  per channel and 8x8 block            (learned transform, a PARAMETER: needs dT)
//     xt    = tr * sigmoid(q2(relu(q0(tr))))
//     A     = 1/4 sum_s bilinear_up(gate_s),  gate_s = sigmoid(W3 relu(W1 adaptive_avg_pool_s(h)))   (s = 1, 2, 4, 8)
//     color = boost_c[b] * sigmoid(c2(relu(c0(h)))),  edge = boost_e[b] * sigmoid(e2(relu(e0(h))))    (1x1 / 3x3 convs)
//     e     = h + xt * A * color * edge
// The GEMM-shaped pieces reuse the forward implicit-GEMM kernels (data gradients) and ddpmir_wgrad / ddpmir_colsum; this file
// holds the element-wise product rule, the two reductions of the gate pyramid (bilinear up-sampling and adaptive pooling are
// linear maps: their backward is the transposed map), the transform's weight gradient and the ReLU mask.
// All gradients are fp32 ("stream" class); saved forward activations may be bf16 operands.
#include "epilogue.cuh"

namespace {

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148 * 16;
    return (int)(g > cap ? cap : g);
}

// ---- product rule of  e = h + xt * A * color * edge ---------------------------------------------------------------------
// thread = 8 consecutive channels of one pixel.  Writes dxt, the PRE-sigmoid gradients of the colour and edge gates
// (color = boost * sig  =>  dz = dcolor * color * (1 - color / boost)) and dA (already times 1/4, the mean over scales).
template <typename T>
__global__ void __launch_bounds__(256)
avif_combine_bwd_kernel(const float* __restrict__ de, const T* __restrict__ xt, const float* __restrict__ gates,
                        const T* __restrict__ color, const T* __restrict__ edge, const float* __restrict__ boost_c,
                        const float* __restrict__ boost_e, float* __restrict__ dxt, float* __restrict__ dzc,
                        float* __restrict__ dze, float* __restrict__ dA, int B, int H, int W, int C) {
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
        for (int k = 0; k < 8; ++k) attn[k] = gb[k];
        const int base[3] = {1, 5, 21};
#pragma unroll
        for (int li = 0; li < 3; ++li) {
            const int s = 2 << li;
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
        Vec8<float> g, o1, o2, o3, o4;
        Vec8<T> xv, cvv, ev;
        g.load(de + i * 8); xv.load(xt + i * 8); cvv.load(color + i * 8); ev.load(edge + i * 8);
        const float bc = boost_c[b], be = boost_e[b];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a = attn[k] * 0.25f, x = xv.v[k], c = cvv.v[k], e = ev.v[k], d = g.v[k];
            o1.v[k] = d * a * c * e;                                  // d xt
            o2.v[k] = d * x * a * e * c * (1.f - c / bc);             // d (pre-sigmoid colour logit)
            o3.v[k] = d * x * a * c * e * (1.f - e / be);             // d (pre-sigmoid edge logit)
            o4.v[k] = d * x * c * e * 0.25f;                          // d (sum of the up-sampled gate maps)
        }
        o1.store(dxt + i * 8); o2.store(dzc + i * 8); o3.store(dze + i * 8); o4.store(dA + i * 8);
    }
}

// ---- transposed bilinear up-sampling: dgate[cell, b, c] = sum_pixels weight(pixel, cell) * dA[b, pixel, c] -------------------
// One CTA per (cell, b), 64 channel lanes x 4 pixel lanes.  A cell (ci, cj) of the s x s map feeds exactly the pixels whose
// bilinear source rows/columns include ci / cj; the window below is a superset and the weight is evaluated per pixel with the
// forward's own index arithmetic (bilin_src), so the two can not disagree at the clamped borders.
__device__ __forceinline__ float bilin_weight(int d, int n_in, int n_out, int cell) {
    int i0, i1; float lam;
    bilin_src(d, n_in, n_out, i0, i1, lam);
    return (i0 == cell ? 1.f - lam : 0.f) + (i1 == cell ? lam : 0.f);
}
__device__ __forceinline__ void bilin_window(int cell, int s, int n, int& lo, int& hi) {
    const float r = (float)n / (float)s;                       // pixels per cell
    int a = (int)floorf(((float)cell - 0.5f) * r - 0.5f) - 1;  // source coordinate (d + .5) / r - .5 in [cell - 1, cell + 1]
    int b = (int)ceilf(((float)cell + 1.5f) * r - 0.5f) + 1;
    if (cell == 0) a = 0;                                      // the source coordinate is clamped at 0 ...
    if (cell == s - 1) b = n;                                  // ... and the index at s - 1
    lo = a < 0 ? 0 : a;
    hi = b > n ? n : b;
}

__global__ void __launch_bounds__(256)
avif_gates_bwd_kernel(const float* __restrict__ dA, float* __restrict__ dgates, int B, int H, int W, int C) {
    const int cell = blockIdx.x, b = blockIdx.y;
    int s, ci, cj;
    pyramid_cell(cell, s, ci, cj);
    int h_lo, h_hi, w_lo, w_hi;
    bilin_window(ci, s, H, h_lo, h_hi);
    bilin_window(cj, s, W, w_lo, w_hi);
    const int ww = w_hi - w_lo, npx = (h_hi - h_lo) * ww;
    __shared__ float red[256];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int c0 = 0; c0 < C; c0 += 64) {
        const int c = c0 + tx;
        float a = 0.f;
        if (c < C)
            for (int p = ty; p < npx; p += 4) {
                const int h = h_lo + p / ww, w = w_lo + p % ww;
                const float wt = bilin_weight(h, s, H, ci) * bilin_weight(w, s, W, cj);
                if (wt != 0.f) a = fmaf(wt, dA[(((long long)b * H + h) * W + w) * C + c], a);
            }
        red[threadIdx.x] = a;
        __syncthreads();
        if (ty == 0 && c < C)
            dgates[((long long)cell * B + b) * C + c] = red[tx] + red[64 + tx] + red[128 + tx] + red[192 + tx];
        __syncthreads();
    }
}

// ---- transposed adaptive average pooling: dx[b, h, w, c] (+)= sum over cells containing (h, w) of dpooled[cell, b, c] / |cell| --
// cell i of s covers [floor(i*n/s), ceil((i+1)*n/s)) (the forward's windows; they overlap when s does not divide n and
// replicate pixels when s > n), so every pixel simply scans the <= 8 cells per axis of each scale.
__global__ void __launch_bounds__(256)
avgpool_pyramid_bwd_kernel(const float* __restrict__ dpooled, float* __restrict__ dx, int B, int H, int W, int C, int accumulate) {
    const int cv = C >> 3;
    const long long total = (long long)B * H * W * cv;
    const long long CS = (long long)B * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % cv);
        long long p = i / cv;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        const float* gb = dpooled + (long long)b * C + v * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int base[4] = {0, 1, 5, 21};
#pragma unroll
        for (int li = 0; li < 4; ++li) {
            const int s = 1 << li;
            for (int ci = 0; ci < s; ++ci) {
                const int h0 = (ci * H) / s, h1 = ((ci + 1) * H + s - 1) / s;
                if (h < h0 || h >= h1) continue;
                for (int cj = 0; cj < s; ++cj) {
                    const int w0 = (cj * W) / s, w1 = ((cj + 1) * W + s - 1) / s;
                    if (w < w0 || w >= w1) continue;
                    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
                    const float* g = gb + (long long)(base[li] + ci * s + cj) * CS;
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(g[k], inv, acc[k]);
                }
            }
        }
        Vec8<float> o;
        if (accumulate) {
            o.load(dx + i * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] += acc[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = acc[k];
        }
        o.store(dx + i * 8);
    }
}

// ---- weight gradient of the learned per-channel transform  Z = T X T^T ----------------------------------------------------
// Per block and channel, with Y = T X:   dT += dZ^T Y + (dZ T) X^T.   CTA = 64 channels x 4 block lanes (the forward's layout);
// every lane walks its share of the spatial blocks with the 8 x 8 accumulator in registers, one row u of Y and dY at a time;
// lanes are summed in shared memory and added to dT with one atomic per (channel, element) and CTA.
template <int BS>
__global__ void __launch_bounds__(256)
block_transform_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dz, const float* __restrict__ Tm,
                             float* __restrict__ dT, int B, int H, int W, int C) {
    __shared__ float Ts[BS * BS][64];
    __shared__ float red[BS * BS][64];
    const int c = blockIdx.y * 64 + threadIdx.x;
    const int tid = threadIdx.y * 64 + threadIdx.x;
    for (int i = tid; i < BS * BS * 64; i += 256) {
        const int ch = i & 63, e = i >> 6;
        const int cc = blockIdx.y * 64 + ch;
        Ts[e][ch] = cc < C ? Tm[(long long)cc * BS * BS + e] : 0.f;
    }
    __syncthreads();
    const int nbh = (H + BS - 1) / BS, nbw = (W + BS - 1) / BS;
    const long long nblk = (long long)B * nbh * nbw;
    float acc[BS][BS];
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
        for (int j = 0; j < BS; ++j) acc[i][j] = 0.f;
    if (c < C)
        for (long long blk = (long long)blockIdx.x * 4 + threadIdx.y; blk < nblk; blk += (long long)gridDim.x * 4) {
            const int b = (int)(blk / (nbh * nbw));
            const int r = (int)(blk - (long long)b * nbh * nbw);
            const int h0 = (r / nbw) * BS, w0 = (r % nbw) * BS;
            const long long img = (long long)b * H * W * C + c;
            float X[BS][BS];
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
                for (int j = 0; j < BS; ++j) {
                    const int h = h0 + i, w = w0 + j;
                    X[i][j] = (h < H && w < W) ? x[img + ((long long)h * W + w) * C] : 0.f;
                }
#pragma unroll
            for (int u = 0; u < BS; ++u) {
                float Yu[BS], dZu[BS], dYu[BS];
#pragma unroll
                for (int j = 0; j < BS; ++j) {
                    float a = 0.f;
#pragma unroll
                    for (int i = 0; i < BS; ++i) a = fmaf(Ts[u * BS + i][threadIdx.x], X[i][j], a);
                    Yu[j] = a;
                    const int h = h0 + u, w = w0 + j;
                    dZu[j] = (h < H && w < W) ? dz[img + ((long long)h * W + w) * C] : 0.f;   // the crop's gradient is zero
                }
#pragma unroll
                for (int j = 0; j < BS; ++j) {
                    float a = 0.f;
#pragma unroll
                    for (int v = 0; v < BS; ++v) a = fmaf(dZu[v], Ts[v * BS + j][threadIdx.x], a);
                    dYu[j] = a;
                }
#pragma unroll
                for (int v = 0; v < BS; ++v)
#pragma unroll
                    for (int j = 0; j < BS; ++j) acc[v][j] = fmaf(dZu[v], Yu[j], acc[v][j]);      // Z = Y T^T: dT[v][j] += dZ[u][v] Y[u][j]
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    float a = 0.f;
#pragma unroll
                    for (int j = 0; j < BS; ++j) a = fmaf(dYu[j], X[i][j], a);
                    acc[u][i] += a;                                                               // Y = T X:  dT[u][i] += dY[u][j] X[i][j]
                }
            }
        }
    for (int lane = 1; lane < 4; ++lane) {       // fold the block lanes into lane 0, one at a time (16 KB of shared memory)
        if (threadIdx.y == lane) {
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
                for (int j = 0; j < BS; ++j) red[i * BS + j][threadIdx.x] = acc[i][j];
        }
        __syncthreads();
        if (threadIdx.y == 0) {
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
                for (int j = 0; j < BS; ++j) acc[i][j] += red[i * BS + j][threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int j = 0; j < BS; ++j) atomicAdd(dT + (long long)c * BS * BS + i * BS + j, acc[i][j]);
    }
}

// ---- ReLU mask: dpre = dy * [y > 0]  (y = relu(pre): same sign test as on the pre-activation) --------------------------------
__global__ void __launch_bounds__(256)
relu_mask_bwd_kernel(const float* __restrict__ dy, const void* __restrict__ y, int y_dtype, float* __restrict__ dpre, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        dpre[i] = ld_any(y, y_dtype, i) > 0.f ? dy[i] : 0.f;
}

}  // namespace

extern "C" int ddpmir_avif_combine_backward(const float* de, const void* xt, const float* gates, const void* color, const void* edge,
                                            int dtype, const float* boost_color, const float* boost_edge, int B, int H, int W,
                                            int C, float* dxt, float* dz_color, float* dz_edge, float* dattn, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(de && xt && gates && color && edge && boost_color && boost_edge && dxt && dz_color && dz_edge && dattn,
                     "avif_combine_backward: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "avif_combine_backward: bad shape");
    const long long total = (long long)B * H * W * (C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_F32)
        avif_combine_bwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(de, (const float*)xt, gates, (const float*)color,
                                                                            (const float*)edge, boost_color, boost_edge, dxt, dz_color,
                                                                            dz_edge, dattn, B, H, W, C);
    else if (dtype == DDPMIR_BF16)
        avif_combine_bwd_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>(de, (const bf16*)xt, gates, (const bf16*)color,
                                                                           (const bf16*)edge, boost_color, boost_edge, dxt, dz_color,
                                                                           dz_edge, dattn, B, H, W, C);
    else DDPMIR_CHECK_ARG(false, "avif_combine_backward: dtype %d", dtype);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_avif_gates_backward(const float* dattn, int B, int H, int W, int C, float* dgates, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dattn && dgates && B > 0 && H > 0 && W > 0 && C > 0, "avif_gates_backward: bad arguments");
    avif_gates_bwd_kernel<<<dim3(85, B), 256, 0, (cudaStream_t)stream>>>(dattn, dgates, B, H, W, C);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_avgpool_pyramid_backward(const float* dpooled, int B, int H, int W, int C, float* dx, int accumulate,
                                               ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dpooled && dx && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "avgpool_pyramid_backward: bad arguments");
    const long long total = (long long)B * H * W * (C / 8);
    avgpool_pyramid_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dpooled, dx, B, H, W, C, accumulate);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_block_transform_wgrad(const float* x, const float* dz, const float* T, int bs, int B, int H, int W, int C,
                                            float* dT, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && dz && T && dT && B > 0 && H > 0 && W > 0 && C > 0, "block_transform_wgrad: bad arguments");
    DDPMIR_CHECK_ARG(bs == 8, "block_transform_wgrad: block size %d (the learned transform is 8x8)", bs);
    const long long nblk = (long long)B * ((H + 7) / 8) * ((W + 7) / 8);
    long long gx = (nblk + 3) / 4;
    if (gx > 148 * 2) gx = 148 * 2;
    block_transform_wgrad_kernel<8><<<dim3((unsigned)gx, ceil_div(C, 64)), dim3(64, 4), 0, (cudaStream_t)stream>>>(x, dz, T, dT, B, H, W, C);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_relu_mask_backward(const float* dy, const void* y, int y_dtype, float* dpre, int64_t n, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(dy && y && dpre && n >= 0, "relu_mask_backward: bad arguments");
    DDPMIR_CHECK_ARG(y_dtype == DDPMIR_F32 || y_dtype == DDPMIR_BF16, "relu_mask_backward: dtype %d", y_dtype);
    if (n == 0) return DDPMIR_OK;
    relu_mask_bwd_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, y, y_dtype, dpre, n);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
