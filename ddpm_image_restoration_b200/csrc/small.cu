// Tiny fp32 kernels that run once per UNet forward: sinusoidal time features and row-wise linears
// (TimeEmbedding webp_inference.py:145-151, time_proj :308, pooled multi-scale gate MLPs avif_inference.py:193-201).
#include "common.cuh"

namespace {

__global__ void time_features_kernel(const float* __restrict__ t, int B, int dim, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = dim / 2;
    if (i >= B * half) return;
    const int b = i / half, k = i % half;
    // emb = exp(arange(half) * -(ln(1e4) / (half - 1)))  evaluated in fp32 like the reference
    const float coef = -(float)(9.210340371976184 / (double)(half - 1));
    const float f = expf((float)k * coef);
    const float a = t[b] * f;
    out[(long long)b * dim + k] = sinf(a);
    out[(long long)b * dim + half + k] = cosf(a);
}

// one warp per (row, n)
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ in, int rows, int K, const float* __restrict__ w,
                   const float* __restrict__ bias, int N, int act, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = blockIdx.x * 8 + wid;
    const int r = blockIdx.y;
    if (n >= N) return;
    const float* a = in + (long long)r * K;
    const float* ww = w + (long long)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(a[k], ww[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[(long long)r * N + n] = act_apply(act, acc + (bias ? bias[n] : 0.f));
}

}  // namespace

extern "C" int ddpmir_linear_rows(const float* in, int rows, int K, const float* w, const float* bias, int N, int act,
                                  float* out, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(in && w && out && rows > 0 && K > 0 && N > 0, "linear_rows: bad arguments");
    DDPMIR_CHECK_ARG(rows <= 65535, "linear_rows: too many rows (%d)", rows);
    linear_rows_kernel<<<dim3(ceil_div(N, 8), rows), 256, 0, (cudaStream_t)stream>>>(in, rows, K, w, bias, N, act, out);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_time_features(const float* t, int B, int dim, float* out, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(t && out && B > 0 && dim >= 4 && dim % 2 == 0, "time_features: bad arguments");
    time_features_kernel<<<ceil_div((long long)B * dim / 2, 128), 128, 0, (cudaStream_t)stream>>>(t, B, dim, out);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_time_embed(const float* t, int B, int dim, const float* w0, const float* b0, const float* w1,
                                 const float* b1, float* ws, float* out, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(t && w0 && b0 && w1 && b1 && ws && out && B > 0 && dim >= 4 && dim % 2 == 0, "time_embed: bad arguments");
    // ws: [B, dim] features followed by [B, 4*dim] hidden
    float* feat = ws;
    float* hid = ws + (long long)B * dim;
    time_features_kernel<<<ceil_div((long long)B * dim / 2, 128), 128, 0, (cudaStream_t)stream>>>(t, B, dim, feat);
    DDPMIR_LAUNCH_CHECK();
    int rc = ddpmir_linear_rows(feat, B, dim, w0, b0, 4 * dim, DDPMIR_ACT_SILU, hid, stream);
    if (rc) return rc;
    return ddpmir_linear_rows(hid, B, 4 * dim, w1, b1, dim, DDPMIR_ACT_NONE, out, stream);
}
