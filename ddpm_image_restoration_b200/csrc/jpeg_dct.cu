// DCT-domain JPEG projection on the GPU: the reference's pure-torch JPEG simulator
// DCTProcessor.jpeg_compress (experiments/code/dct.ipynb cell 2, L43-139: scalar Python loops over every 8x8 block) as one
// HBM-bound kernel -- per channel and 8x8 block: orthonormal DCT-II of (block - 128), round(c / Q) * Q with the luma table for
// channel 0 and the chroma table for the others (tables scaled by quality, L105-112), inverse DCT, + 128.
// One read and one write per sample; a CTA owns an 8-row strip of 256 columns in shared memory: column pass in registers
// (thread = column), row pass + quantisation (thread = (coefficient row, block)), inverse row pass, inverse column pass.
#include "common.cuh"
#include <math.h>

namespace {

struct DctParams {
    float D[64];       // D[u * 8 + x] = 0.5 * c_u * cos((2x + 1) u pi / 16)
    float Q[2][64];    // [0] luma, [1] chroma; Q[u * 8 + v]
};

constexpr int TW = 256;                 // columns per CTA
constexpr int PITCH = TW + TW / 8;      // element (u, col) lives at u * PITCH + col + (col >> 3): 8-float groups are 9 apart

__global__ void __launch_bounds__(TW)
jpeg_dct_project_kernel(const float* __restrict__ x, float* __restrict__ out, int C, int H, int W, float in_scale, float in_offset,
                        const __grid_constant__ DctParams P) {
    __shared__ float tile[8 * PITCH];
    const int t = threadIdx.x;
    const int col = blockIdx.x * TW + t;
    const int row0 = blockIdx.y * 8;
    const int plane = blockIdx.z;
    const int table = (plane % C) == 0 ? 0 : 1;
    const float* xp = x + ((long long)plane * H + row0) * W;
    float* op = out + ((long long)plane * H + row0) * W;
    const bool ok = col < W;

    // ---- forward column pass (thread = column) ----
    float v[8], c[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = ok ? fmaf(xp[(long long)r * W + col], in_scale, in_offset) - 128.f : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) a = fmaf(P.D[u * 8 + r], v[r], a);
        tile[u * PITCH + t + (t >> 3)] = a;
    }
    __syncthreads();
    // ---- forward row pass, quantise / dequantise, inverse row pass (thread = coefficient row u of block j) ----
    {
        const int u = t >> 5, j = t & 31;
        float* rp = tile + u * PITCH + j * 9;
        float r[8], d[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = rp[k];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) a = fmaf(P.D[w * 8 + k], r[k], a);
            const float q = P.Q[table][u * 8 + w];
            d[w] = rintf(__fdiv_rn(a, q)) * q;            // torch.round is round-half-to-even
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) a = fmaf(P.D[w * 8 + k], d[w], a);
            rp[k] = a;
        }
    }
    __syncthreads();
    // ---- inverse column pass (thread = column) ----
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = tile[u * PITCH + t + (t >> 3)];
    const float inv = 1.f / in_scale;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        float a = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) a = fmaf(P.D[u * 8 + r], c[u], a);
        if (ok) op[(long long)r * W + col] = (a + 128.f - in_offset) * inv;
    }
}

}  // namespace

extern "C" int ddpmir_jpeg_dct_project(const float* x, float* out, int B, int C, int H, int W, float quality, float in_scale,
                                       float in_offset, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out, "jpeg_dct_project: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, "jpeg_dct_project: H and W must be multiples of 8");
    DDPMIR_CHECK_ARG(quality > 0.f && quality <= 100.f && in_scale != 0.f, "jpeg_dct_project: bad quality or scale");
    DDPMIR_CHECK_ARG((long long)B * C <= 65535 && H / 8 <= 65535, "jpeg_dct_project: grid too large");
    static const float QY[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                                 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                                 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const float QC[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                                 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    DctParams P;
    const double pi = 3.14159265358979323846;
    for (int u = 0; u < 8; ++u)
        for (int k = 0; k < 8; ++k) P.D[u * 8 + k] = (float)(0.5 * (u == 0 ? 1.0 / sqrt(2.0) : 1.0) * cos((2 * k + 1) * u * pi / 16.0));
    // dct.ipynb#c2:L105-112 in fp32 as torch does: scale = 50/q (q < 50) or 2 - q/50; Q = max(round_half_even(table * scale), 1)
    const float scale = quality < 50.f ? 50.f / quality : 2.f - quality / 50.f;
    for (int i = 0; i < 64; ++i) {
        P.Q[0][i] = fmaxf(nearbyintf(QY[i] * scale), 1.f);
        P.Q[1][i] = fmaxf(nearbyintf(QC[i] * scale), 1.f);
    }
    dim3 grid(ceil_div(W, TW), H / 8, B * C);
    jpeg_dct_project_kernel<<<grid, TW, 0, (cudaStream_t)stream>>>(x, out, C, H, W, in_scale, in_offset, P);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
