// Optimiser step of the training loop (webp_training.py:521-524, 775): clip_grad_norm_(1.0) + AdamW, fused per tensor.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fmaf(x[i], x[i], s);
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

// torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6)); torch.optim.AdamW (decoupled decay)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
             float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt, const double* __restrict__ grad_sumsq,
             float max_norm) {
    float coef = 1.f;
    if (grad_sumsq) {
        const float total = (float)sqrt(*grad_sumsq);
        coef = fminf(1.f, max_norm / (total + 1e-6f));
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

// the same update over MANY tensors in one launch: the gradient and both moments live in flat buffers (Trainer's layout); a chunk
// table maps CTA -> (offset into the flat buffers, length, pointer to the matching parameter elements).  342 per-tensor launches
// of ~5 us each were 1.7 ms of the training step for 0.5 ms of memory traffic.
__global__ void __launch_bounds__(256)
adamw_multi_kernel(const long long* __restrict__ chunk_start, const int* __restrict__ chunk_len, float* const* __restrict__ chunk_param,
                   const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float lr, float beta1, float beta2,
                   float eps, float wd, float bc1, float bc2_sqrt, const double* __restrict__ grad_sumsq, float max_norm) {
    float coef = 1.f;
    if (grad_sumsq) {
        const float total = (float)sqrt(*grad_sumsq);
        coef = fminf(1.f, max_norm / (total + 1e-6f));
    }
    const long long s = chunk_start[blockIdx.x];
    const int n = chunk_len[blockIdx.x];
    float* __restrict__ p = chunk_param[blockIdx.x];
    for (int i = threadIdx.x; i < n; i += 256) {
        const float gi = g[s + i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[s + i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[s + i] + (1.f - beta2) * gi * gi;
        m[s + i] = mi; v[s + i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

}  // namespace

extern "C" int ddpmir_adamw_multi(const int64_t* chunk_start, const int32_t* chunk_len, float* const* chunk_param, int n_chunks,
                                  const float* g_flat, float* m_flat, float* v_flat, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, int step, const double* grad_sumsq, float max_norm, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(chunk_start && chunk_len && chunk_param && g_flat && m_flat && v_flat && n_chunks > 0 && step >= 1,
                     "adamw_multi: bad arguments");
    const float bc1 = (float)(1.0 - pow((double)beta1, step));
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, step));
    adamw_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const long long*)chunk_start, chunk_len, chunk_param, g_flat, m_flat,
                                                                    v_flat, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_sumsq,
                                                                    max_norm);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_sumsq(const float* x, int64_t n, double* acc, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && acc && n > 0, "sumsq: bad arguments");
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 4) grid = 148 * 4;
    sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, acc);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, const double* grad_sumsq, float max_norm,
                                 ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "adamw_step: bad arguments");
    const float bc1 = (float)(1.0 - pow((double)beta1, step));
    const float bc2s = (float)sqrt(1.0 - pow((double)beta2, step));
    int grid = (int)((n + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, grad_sumsq, max_norm);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
