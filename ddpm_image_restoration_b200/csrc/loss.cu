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

__global__ void scale_kernel(const double* __restrict__ acc, double scale, float* __restrict__ out) { out[0] = (float)(acc[0] * scale); }

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
