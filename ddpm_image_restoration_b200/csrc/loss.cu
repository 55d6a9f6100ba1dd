// Loss reductions (forward values): mean squared error and SSIM (pytorch_msssim.ssim semantics: gaussian window 11,
// sigma 1.5, separable 'valid' filtering, K = (0.01, 0.03), mean over all outputs) on NCHW fp32 images.
// Call sites: frequency_aware_loss webp_training.py:108,129; color_preservation_loss 0409_method.ipynb#c0:L79.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sq_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4, double* __restrict__ acc) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    s = warp_sum(s);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += (double)red[w];
        atomicAdd(acc, v);
    }
}

constexpr int WIN = 11, TILE = 16, IN = TILE + WIN - 1;   // 16x16 outputs need 26x26 inputs

struct Gauss { float w[WIN]; };

__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int clamp01, Gauss g,
            double* __restrict__ acc) {
    __shared__ float sx[IN][IN + 1], sy[IN][IN + 1];
    __shared__ float hf[5][IN][TILE + 1];
    const int plane = blockIdx.z;
    const int oh0 = blockIdx.y * TILE, ow0 = blockIdx.x * TILE;
    const int OH = H - WIN + 1, OW = W - WIN + 1;
    const float* px = x + (long long)plane * H * W;
    const float* py = y + (long long)plane * H * W;
    for (int i = threadIdx.x; i < IN * IN; i += 256) {
        const int r = i / IN, c = i - r * IN;
        const int h = oh0 + r, w = ow0 + c;
        float a = 0.f, b = 0.f;
        if (h < H && w < W) {
            a = __fadd_rn(__fmul_rn(px[(long long)h * W + w], 0.5f), 0.5f);
            b = __fadd_rn(__fmul_rn(py[(long long)h * W + w], 0.5f), 0.5f);
            if (clamp01) { a = fminf(fmaxf(a, 0.f), 1.f); b = fminf(fmaxf(b, 0.f), 1.f); }
        }
        sx[r][c] = a; sy[r][c] = b;
    }
    __syncthreads();
    // pytorch_msssim filters dim 2 (H) first, then dim 3 (W); the separable result is the same up to rounding.
    for (int i = threadIdx.x; i < IN * TILE; i += 256) {
        const int r = i / TILE, c = i - r * TILE;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const float a = sx[r][c + k], b = sy[r][c + k], wk = g.w[k];
            m1 = fmaf(wk, a, m1); m2 = fmaf(wk, b, m2);
            s11 = fmaf(wk, a * a, s11); s22 = fmaf(wk, b * b, s22); s12 = fmaf(wk, a * b, s12);
        }
        hf[0][r][c] = m1; hf[1][r][c] = m2; hf[2][r][c] = s11; hf[3][r][c] = s22; hf[4][r][c] = s12;
    }
    __syncthreads();
    float v = 0.f;
    {
        const int r = threadIdx.x / TILE, c = threadIdx.x % TILE;
        if (oh0 + r < OH && ow0 + c < OW) {
            float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
            for (int k = 0; k < WIN; ++k) {
                const float wk = g.w[k];
                m1 = fmaf(wk, hf[0][r + k][c], m1); m2 = fmaf(wk, hf[1][r + k][c], m2);
                s11 = fmaf(wk, hf[2][r + k][c], s11); s22 = fmaf(wk, hf[3][r + k][c], s22); s12 = fmaf(wk, hf[4][r + k][c], s12);
            }
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float v11 = s11 - m1 * m1, v22 = s22 - m2 * m2, v12 = s12 - m1 * m2;
            const float cs = (2.f * v12 + C2) / (v11 + v22 + C2);
            v = ((2.f * m1 * m2 + C1) / (m1 * m1 + m2 * m2 + C1)) * cs;
        }
    }
    v = warp_sum(v);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += (double)red[w];
        atomicAdd(acc, t);
    }
}

// ---- SSIM backward (d mean-SSIM / dX, Y constant) -----------------------------------------------------------------------
// pass 1: per output position the partials wrt the three filtered maps that depend on X: mu1, E[X^2], E[XY]
__global__ void __launch_bounds__(256)
ssim_bwd_partials_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int clamp01, Gauss g,
                         float* __restrict__ gmaps /* [planes][3][OH][OW] */) {
    __shared__ float sx[IN][IN + 1], sy[IN][IN + 1];
    __shared__ float hf[5][IN][TILE + 1];
    const int plane = blockIdx.z;
    const int oh0 = blockIdx.y * TILE, ow0 = blockIdx.x * TILE;
    const int OH = H - WIN + 1, OW = W - WIN + 1;
    const float* px = x + (long long)plane * H * W;
    const float* py = y + (long long)plane * H * W;
    for (int i = threadIdx.x; i < IN * IN; i += 256) {
        const int r = i / IN, c = i - r * IN;
        const int h = oh0 + r, w = ow0 + c;
        float a = 0.f, b = 0.f;
        if (h < H && w < W) {
            a = __fadd_rn(__fmul_rn(px[(long long)h * W + w], 0.5f), 0.5f);
            b = __fadd_rn(__fmul_rn(py[(long long)h * W + w], 0.5f), 0.5f);
            if (clamp01) { a = fminf(fmaxf(a, 0.f), 1.f); b = fminf(fmaxf(b, 0.f), 1.f); }
        }
        sx[r][c] = a; sy[r][c] = b;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IN * TILE; i += 256) {
        const int r = i / TILE, c = i - r * TILE;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const float a = sx[r][c + k], b = sy[r][c + k], wk = g.w[k];
            m1 = fmaf(wk, a, m1); m2 = fmaf(wk, b, m2);
            s11 = fmaf(wk, a * a, s11); s22 = fmaf(wk, b * b, s22); s12 = fmaf(wk, a * b, s12);
        }
        hf[0][r][c] = m1; hf[1][r][c] = m2; hf[2][r][c] = s11; hf[3][r][c] = s22; hf[4][r][c] = s12;
    }
    __syncthreads();
    const int r = threadIdx.x / TILE, c = threadIdx.x % TILE;
    if (oh0 + r < OH && ow0 + c < OW) {
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const float wk = g.w[k];
            m1 = fmaf(wk, hf[0][r + k][c], m1); m2 = fmaf(wk, hf[1][r + k][c], m2);
            e11 = fmaf(wk, hf[2][r + k][c], e11); e22 = fmaf(wk, hf[3][r + k][c], e22); e12 = fmaf(wk, hf[4][r + k][c], e12);
        }
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * (e12 - m1 * m2) + C2;
        const float B1 = m1 * m1 + m2 * m2 + C1, B2 = (e11 - m1 * m1) + (e22 - m2 * m2) + C2;
        const float inv = 1.f / (B1 * B2);
        const float g_e12 = 2.f * A1 * inv;
        const float g_e11 = -A1 * A2 * inv / B2;
        // mu1 enters A1 (2 mu2), A2 (-2 mu2), B1 (2 mu1), B2 (-2 mu1)
        const float g_mu = (2.f * m2 * A2 - 2.f * m2 * A1) * inv - A1 * A2 * inv * inv * (2.f * m1 * B2 - 2.f * m1 * B1);
        float* o = gmaps + ((long long)plane * 3) * OH * OW + (long long)(oh0 + r) * OW + (ow0 + c);
        o[0] = g_mu; o[(long long)OH * OW] = g_e11; o[2LL * OH * OW] = g_e12;
    }
}

// pass 2: dX[h,w] += coef * sum_{i,j} g(h-i) g(w-j) [ gmu(i,j) + 2 X gE11(i,j) + Y gE12(i,j) ]  (transpose of the valid filter)
__global__ void __launch_bounds__(256)
ssim_bwd_scatter_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gmaps, int planes, int H,
                        int W, int clamp01, Gauss g, float coef, float* __restrict__ dx) {
    const int OH = H - WIN + 1, OW = W - WIN + 1;
    const long long total = (long long)planes * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int plane = (int)(i / ((long long)H * W));
        const int rem = (int)(i - (long long)plane * H * W);
        const int h = rem / W, w = rem - h * W;
        float a = __fadd_rn(__fmul_rn(x[i], 0.5f), 0.5f), b = __fadd_rn(__fmul_rn(y[i], 0.5f), 0.5f);
        float pass = 1.f;
        if (clamp01) { if (a <= 0.f || a >= 1.f) pass = 0.f; a = fminf(fmaxf(a, 0.f), 1.f); b = fminf(fmaxf(b, 0.f), 1.f); }
        const float* gm = gmaps + ((long long)plane * 3) * OH * OW;
        float s_mu = 0.f, s_11 = 0.f, s_12 = 0.f;
        for (int u = 0; u < WIN; ++u) {
            const int oi = h - u;
            if (oi < 0 || oi >= OH) continue;
            for (int v = 0; v < WIN; ++v) {
                const int oj = w - v;
                if (oj < 0 || oj >= OW) continue;
                const float wv = g.w[u] * g.w[v];
                const long long o = (long long)oi * OW + oj;
                s_mu = fmaf(wv, gm[o], s_mu);
                s_11 = fmaf(wv, gm[(long long)OH * OW + o], s_11);
                s_12 = fmaf(wv, gm[2LL * OH * OW + o], s_12);
            }
        }
        dx[i] += coef * pass * (s_mu + 2.f * a * s_11 + b * s_12);
    }
}

__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float coef, float* __restrict__ da, long long n, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = coef * (a[i] - b[i]);
        da[i] = accumulate ? da[i] + v : v;
    }
}

// Huber loss (nn.HuberLoss(reduction='mean', delta), 0409_method.ipynb#c0:L438, 567): 0.5 d^2 if |d| <= delta else delta (|d| - 0.5 delta)
__global__ void __launch_bounds__(256)
huber_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4, float delta, double* __restrict__ acc) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
        const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ad = fabsf(d[k]);
            s += ad <= delta ? 0.5f * d[k] * d[k] : delta * (ad - 0.5f * delta);
        }
    }
    s = warp_sum(s);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0;
        for (int w = 0; w < 8; ++w) v += (double)red[w];
        atomicAdd(acc, v);
    }
}
__global__ void __launch_bounds__(256)
huber_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float delta, float coef, float* __restrict__ da, long long n,
                 int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        const float v = coef * fminf(fmaxf(d, -delta), delta);       // d inside the quadratic zone, +-delta outside
        da[i] = accumulate ? da[i] + v : v;
    }
}
// colour L1 (0409_method.ipynb#c0:L67-76): d/dpred of 0.25 L1_R + 0.5 L1_G + 0.25 L1_B on x01 = clamp(0.5 x + 0.5, 0, 1), each L1 a
// mean over B*H*W.  The clamp passes gradient strictly inside (0, 1) (torch.clamp's backward masks with min <= x <= max; on
// the boundary the difference of two clamped values has measure-zero support, so the convention does not show in practice).
__global__ void __launch_bounds__(256)
color_l1_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, int HW, long long n, float coef,
                    float* __restrict__ dpred, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / HW) % 3);
        const float p01 = __fadd_rn(__fmul_rn(pred[i], 0.5f), 0.5f), t01 = __fadd_rn(__fmul_rn(target[i], 0.5f), 0.5f);
        const float a = fminf(fmaxf(p01, 0.f), 1.f), b = fminf(fmaxf(t01, 0.f), 1.f);
        const float pass = (p01 >= 0.f && p01 <= 1.f) ? 1.f : 0.f;
        const float sgn = a > b ? 1.f : (a < b ? -1.f : 0.f);
        const float v = coef * (c == 1 ? 0.5f : 0.25f) * pass * sgn;
        dpred[i] = accumulate ? dpred[i] + v : v;
    }
}

__global__ void scale_kernel(const double* __restrict__ acc, double scale, float* __restrict__ out) { out[0] = (float)(acc[0] * scale); }

// gradient_loss of avif_frequency_aware_loss (avif.py:136-144) on the [0,1] images x*0.5+0.5: sums of
// (|x_i - x_down| - |y_i - y_down|)^2 (acc[0], over H-1 rows) and (|x_i - x_right| - |y_i - y_right|)^2 (acc[1], over W-1 columns)
__global__ void __launch_bounds__(256)
edge_loss_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, long long n, double* __restrict__ acc) {
    float sv = 0.f, sh = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const float xi = x[i] * 0.5f, yi = y[i] * 0.5f;            // the +0.5 cancels in every difference
        if (h + 1 < H) { const float d = fabsf(xi - x[i + W] * 0.5f) - fabsf(yi - y[i + W] * 0.5f); sv = fmaf(d, d, sv); }
        if (w + 1 < W) { const float d = fabsf(xi - x[i + 1] * 0.5f) - fabsf(yi - y[i + 1] * 0.5f); sh = fmaf(d, d, sh); }
    }
    sv = warp_sum(sv); sh = warp_sum(sh);
    __shared__ float red[2][8];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sh; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0;
        for (int k = 0; k < 8; ++k) v += (double)red[threadIdx.x][k];
        atomicAdd(&acc[threadIdx.x], v);
    }
}
// dx (+)= wv d acc[0]/dx + wh d acc[1]/dx: every pixel collects the (up to four) difference terms it takes part in.
// d |a - b| / da = sign(a - b) (0 at a == b, torch.abs's convention).
__global__ void __launch_bounds__(256)
edge_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, long long n, float wv, float wh,
                     float* __restrict__ dx, int accumulate) {
    auto term = [](float xa, float xb, float ya, float yb) {      // d/d xa of (|xa - xb| - |ya - yb|)^2
        const float dxv = xa - xb;
        const float sgn = dxv > 0.f ? 1.f : (dxv < 0.f ? -1.f : 0.f);
        return 2.f * (fabsf(dxv) - fabsf(ya - yb)) * sgn;
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % W), h = (int)((i / W) % H);
        const float xi = x[i] * 0.5f, yi = y[i] * 0.5f;
        float g = 0.f;
        if (h + 1 < H) g += wv * term(xi, x[i + W] * 0.5f, yi, y[i + W] * 0.5f);
        if (h > 0) g += wv * term(xi, x[i - W] * 0.5f, yi, y[i - W] * 0.5f);
        if (w + 1 < W) g += wh * term(xi, x[i + 1] * 0.5f, yi, y[i + 1] * 0.5f);
        if (w > 0) g += wh * term(xi, x[i - 1] * 0.5f, yi, y[i - 1] * 0.5f);
        g *= 0.5f;                                                 // chain rule of x01 = 0.5 x + 0.5
        dx[i] = accumulate ? dx[i] + g : g;
    }
}

}  // namespace

extern "C" int ddpmir_mse(const float* a, const float* b, int64_t n, float* out_scalar, double* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(a && b && out_scalar && ws && n > 0 && n % 4 == 0, "mse: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws, 0, sizeof(double), st);
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    sq_diff_kernel<<<grid, 256, 0, st>>>(a, b, n / 4, ws);
    scale_kernel<<<1, 1, 0, st>>>(ws, 1.0 / (double)n, out_scalar);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_ssim(const float* x, const float* y, int planes, int H, int W, int clamp01, float* out_scalar,
                           double* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && y && out_scalar && ws && planes > 0, "ssim: bad arguments");
    DDPMIR_CHECK_ARG(H >= WIN && W >= WIN, "ssim: images must be at least 11x11 (got %d x %d)", H, W);
    DDPMIR_CHECK_ARG(planes <= 65535, "ssim: too many planes");
    Gauss g;
    double sum = 0;
    for (int k = 0; k < WIN; ++k) { const double c = k - WIN / 2; g.w[k] = (float)exp(-(c * c) / (2.0 * 1.5 * 1.5)); sum += g.w[k]; }
    for (int k = 0; k < WIN; ++k) g.w[k] = (float)(g.w[k] / sum);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws, 0, sizeof(double), st);
    const int OH = H - WIN + 1, OW = W - WIN + 1;
    dim3 grid(ceil_div(OW, TILE), ceil_div(OH, TILE), planes);
    ssim_kernel<<<grid, 256, 0, st>>>(x, y, H, W, clamp01, g, ws);
    scale_kernel<<<1, 1, 0, st>>>(ws, 1.0 / ((double)planes * OH * OW), out_scalar);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// da (+)= weight * d mse(a,b)/da = weight * 2 (a - b) / n
extern "C" int ddpmir_mse_backward(const float* a, const float* b, int64_t n, float weight, float* da, int accumulate,
                                   ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(a && b && da && n > 0, "mse_backward: bad arguments");
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    mse_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, (float)(2.0 * weight / (double)n), da, n, accumulate);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_huber(const float* a, const float* b, int64_t n, float delta, float* out_scalar, double* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(a && b && out_scalar && ws && n > 0 && n % 4 == 0 && delta > 0.f, "huber: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws, 0, sizeof(double), st);
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    huber_kernel<<<grid, 256, 0, st>>>(a, b, n / 4, delta, ws);
    scale_kernel<<<1, 1, 0, st>>>(ws, 1.0 / (double)n, out_scalar);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_huber_backward(const float* a, const float* b, int64_t n, float delta, float weight, float* da, int accumulate,
                                     ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(a && b && da && n > 0 && delta > 0.f, "huber_backward: bad arguments");
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    huber_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, delta, (float)((double)weight / (double)n), da, n, accumulate);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_color_l1_backward(const float* pred, const float* target, int B, int H, int W, float weight, float* dpred,
                                        int accumulate, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(pred && target && dpred && B > 0 && H > 0 && W > 0, "color_l1_backward: bad arguments");
    const long long n = (long long)B * 3 * H * W;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    // chain rule of x01 = 0.5 x + 0.5 and the mean over B*H*W per channel
    color_l1_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, H * W, n, (float)(0.5 * (double)weight / ((double)B * H * W)),
                                                                  dpred, accumulate);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// dx += weight * d ssim(x*0.5+0.5, y*0.5+0.5) / dx.   ws: planes * 3 * (H-10) * (W-10) floats.
extern "C" int ddpmir_ssim_backward(const float* x, const float* y, int planes, int H, int W, int clamp01, float weight, float* dx,
                                    float* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && y && dx && ws && planes > 0 && planes <= 65535, "ssim_backward: bad arguments");
    DDPMIR_CHECK_ARG(H >= WIN && W >= WIN, "ssim_backward: images must be at least 11x11");
    Gauss g;
    double sum = 0;
    for (int k = 0; k < WIN; ++k) { const double c = k - WIN / 2; g.w[k] = (float)exp(-(c * c) / (2.0 * 1.5 * 1.5)); sum += g.w[k]; }
    for (int k = 0; k < WIN; ++k) g.w[k] = (float)(g.w[k] / sum);
    cudaStream_t st = (cudaStream_t)stream;
    const int OH = H - WIN + 1, OW = W - WIN + 1;
    dim3 grid(ceil_div(OW, TILE), ceil_div(OH, TILE), planes);
    ssim_bwd_partials_kernel<<<grid, 256, 0, st>>>(x, y, H, W, clamp01, g, ws);
    const long long total = (long long)planes * H * W;
    int g2 = (int)((total + 255) / 256);
    if (g2 > 148 * 16) g2 = 148 * 16;
    // mean over planes*OH*OW outputs; chain rule of x01 = 0.5 x + 0.5
    const float coef = (float)(0.5 * (double)weight / ((double)planes * OH * OW));
    ssim_bwd_scatter_kernel<<<g2, 256, 0, st>>>(x, y, ws, planes, H, W, clamp01, g, coef, dx);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// gradient_loss of avif_frequency_aware_loss (avif.py:136-146): acc2[0] / acc2[1] = the two sums of squares over [planes, H, W]
// images (see edge_loss_kernel); the caller divides by planes*(H-1)*W and planes*H*(W-1) (F.mse_loss means).
extern "C" int ddpmir_edge_loss(const float* pred, const float* target, int planes, int H, int W, double* acc2, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(pred && target && acc2 && planes > 0 && H > 1 && W > 1, "edge_loss: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st);
    const long long n = (long long)planes * H * W;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    edge_loss_kernel<<<grid, 256, 0, st>>>(pred, target, H, W, n, acc2);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
// dpred (+)= w_v d acc2[0]/dpred + w_h d acc2[1]/dpred
extern "C" int ddpmir_edge_loss_backward(const float* pred, const float* target, int planes, int H, int W, float w_v, float w_h,
                                         float* dpred, int accumulate, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(pred && target && dpred && planes > 0 && H > 1 && W > 1, "edge_loss_backward: bad arguments");
    const long long n = (long long)planes * H * W;
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    edge_loss_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, H, W, n, w_v, w_h, dpred, accumulate);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
