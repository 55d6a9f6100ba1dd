// Generic flash-style self-attention on CUDA cores (fp32 math, any head_dim that is a multiple of 8, any L).
// This is the fp32 check-mode kernel and the fallback for shapes the tensor-core kernel (attn_mma.cu) does not
// take (ragged L, head_dim 256).  One query is owned by G consecutive lanes, each holding DPT dims of q and o.
#include "common.cuh"

namespace {

template <typename T, int DPT, int G>
__global__ void __launch_bounds__(128)
attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int L, int C, int heads, float scale) {
    constexpr int HD = DPT * G;
    constexpr int QPB = 128 / G;                        // queries per CTA
    constexpr int KT = (4096 / HD) > 64 ? 64 : (4096 / HD);  // keys per tile
    __shared__ float Ks[KT * HD];
    __shared__ float Vs[KT * HD];
    const int b = blockIdx.z, h = blockIdx.y;
    const int g = threadIdx.x % G;
    const int qi = blockIdx.x * QPB + threadIdx.x / G;
    const bool q_ok = qi < L;
    const long long row = 3LL * C;
    const T* base = qkv + (long long)b * L * row + (long long)h * HD;
    float q[DPT], o[DPT];
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        q[d] = q_ok ? to_f(base[(long long)qi * row + g * DPT + d]) * scale : 0.f;
        o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < L; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * HD; i += 128) {
            const int j = i / HD, d = i - j * HD;
            const int kj = k0 + j;
            float kv = 0.f, vv = 0.f;
            if (kj < L) {
                kv = to_f(base[(long long)kj * row + C + d]);
                vv = to_f(base[(long long)kj * row + 2 * C + d]);
            }
            Ks[i] = kv; Vs[i] = vv;
        }
        __syncthreads();
        const int kn = min(KT, L - k0);
        for (int j0 = 0; j0 < kn; j0 += 4) {
            float s[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float a = 0.f;
                const float* kp = Ks + (j0 + u) * HD + g * DPT;
#pragma unroll
                for (int d = 0; d < DPT; ++d) a = fmaf(q[d], kp[d], a);
#pragma unroll
                for (int off = G >> 1; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
                s[u] = (j0 + u < kn) ? a : -INFINITY;
            }
            const float mx = fmaxf(fmaxf(m, fmaxf(s[0], s[1])), fmaxf(s[2], s[3]));
            const float corr = expf(m - mx);  // m = -inf on the first chunk -> 0
            l *= corr;
#pragma unroll
            for (int d = 0; d < DPT; ++d) o[d] *= corr;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float p = expf(s[u] - mx);
                l += p;
                const float* vp = Vs + (j0 + u) * HD + g * DPT;
#pragma unroll
                for (int d = 0; d < DPT; ++d) o[d] = fmaf(p, vp[d], o[d]);
            }
            m = mx;
        }
    }
    if (q_ok) {
        const float inv = 1.f / l;
        T* op = out + ((long long)b * L + qi) * C + (long long)h * HD + g * DPT;
#pragma unroll
        for (int d = 0; d < DPT; ++d) op[d] = from_f<T>(o[d] * inv);
    }
}

template <typename T, int DPT, int G>
void launch(const void* qkv, void* out, int B, int L, int C, int heads, cudaStream_t st) {
    constexpr int QPB = 128 / G;
    const float scale = 1.f / sqrtf((float)(DPT * G));
    dim3 grid(ceil_div(L, QPB), heads, B);
    attn_simt_kernel<T, DPT, G><<<grid, 128, 0, st>>>((const T*)qkv, (T*)out, L, C, heads, scale);
}

template <typename T>
int dispatch(const void* qkv, void* out, int B, int L, int C, int heads, cudaStream_t st) {
    const int hd = C / heads;
    switch (hd) {
        case 8: launch<T, 8, 1>(qkv, out, B, L, C, heads, st); break;
        case 16: launch<T, 16, 1>(qkv, out, B, L, C, heads, st); break;
        case 32: launch<T, 16, 2>(qkv, out, B, L, C, heads, st); break;
        case 64: launch<T, 16, 4>(qkv, out, B, L, C, heads, st); break;
        case 128: launch<T, 16, 8>(qkv, out, B, L, C, heads, st); break;
        case 256: launch<T, 16, 16>(qkv, out, B, L, C, heads, st); break;
        default:
            ddpmir_set_error("attention: unsupported head_dim %d", hd);
            return DDPMIR_ERR_UNSUPPORTED;
    }
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

}  // namespace

int ddpmir_attention_simt(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, cudaStream_t st) {
    if (dtype == DDPMIR_F32) return dispatch<float>(qkv, out, B, L, C, heads, st);
    return dispatch<bf16>(qkv, out, B, L, C, heads, st);
}
