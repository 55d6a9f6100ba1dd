// Generic flash-style self-attention on CUDA cores (fp32 math, any head_dim that is a multiple of 8, any L).
// This is the fp32 check-mode kernel and the fallback for shapes the tensor-core kernel (attn_mma.cu) does not
// take (ragged L, head_dim 256).  One query is owned by G consecutive lanes, each holding DPT dims of q and o.
#include "common.cuh"

namespace {

template <typename T, int DPT, int G>
__global__ void __launch_bounds__(128)
attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int L, int C, int heads, float scale) {
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
        if (lse && g == 0) lse[((long long)b * heads + h) * L + qi] = m + logf(l);   // log-sum-exp of the scaled scores
    }
}

// ---- backward (training step): recompute P = exp(s - lse) from Q, K and the saved log-sum-exp ----------------------------
// delta_i = sum_d dO_i O_i
template <typename T>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const T* __restrict__ o, const float* __restrict__ dout, float* __restrict__ delta, int L, int C, int heads) {
    const int hd = C / heads;
    const long long total = (long long)gridDim.y * heads * L;   // gridDim.y = B
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)heads * L; i += (long long)gridDim.x * blockDim.x) {
        const int b = blockIdx.y;
        const int h = (int)(i / L), q = (int)(i - (long long)h * L);
        const long long base = ((long long)b * L + q) * C + (long long)h * hd;
        float s = 0.f;
        for (int d = 0; d < hd; ++d) s = fmaf(dout[base + d], to_f(o[base + d]), s);
        delta[((long long)b * heads + h) * L + q] = s;
    }
    (void)total;
}

// dQ: one query per G lanes, loop over keys.  dqkv is [B, L, 3C] fp32; this kernel writes the q third.
template <typename T, int DPT, int G>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const T* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta, float* __restrict__ dqkv, int L, int C, int heads, float scale) {
    constexpr int HD = DPT * G;
    constexpr int QPB = 128 / G;
    constexpr int KT = (4096 / HD) > 64 ? 64 : (4096 / HD);
    __shared__ float Ks[KT * HD];
    __shared__ float Vs[KT * HD];
    const int b = blockIdx.z, h = blockIdx.y;
    const int g = threadIdx.x % G;
    const int qi = blockIdx.x * QPB + threadIdx.x / G;
    const bool q_ok = qi < L;
    const long long row = 3LL * C;
    const T* base = qkv + (long long)b * L * row + (long long)h * HD;
    float q[DPT], dO[DPT], dq[DPT];
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        q[d] = q_ok ? to_f(base[(long long)qi * row + g * DPT + d]) : 0.f;
        dO[d] = q_ok ? dout[((long long)b * L + qi) * C + (long long)h * HD + g * DPT + d] : 0.f;
        dq[d] = 0.f;
    }
    const float lse_i = q_ok ? lse[((long long)b * heads + h) * L + qi] : 0.f;
    const float del_i = q_ok ? delta[((long long)b * heads + h) * L + qi] : 0.f;
    for (int k0 = 0; k0 < L; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * HD; i += 128) {
            const int j = i / HD, d = i - j * HD;
            const int kj = k0 + j;
            Ks[i] = kj < L ? to_f(base[(long long)kj * row + C + d]) : 0.f;
            Vs[i] = kj < L ? to_f(base[(long long)kj * row + 2 * C + d]) : 0.f;
        }
        __syncthreads();
        const int kn = min(KT, L - k0);
        for (int j = 0; j < kn; ++j) {
            float s = 0.f, dp = 0.f;
            const float* kp = Ks + j * HD + g * DPT;
            const float* vp = Vs + j * HD + g * DPT;
#pragma unroll
            for (int d = 0; d < DPT; ++d) { s = fmaf(q[d], kp[d], s); dp = fmaf(dO[d], vp[d], dp); }
#pragma unroll
            for (int off = G >> 1; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); dp += __shfl_xor_sync(0xffffffffu, dp, off); }
            const float p = expf(s * scale - lse_i);
            const float ds = p * (dp - del_i) * scale;
#pragma unroll
            for (int d = 0; d < DPT; ++d) dq[d] = fmaf(ds, kp[d], dq[d]);
        }
    }
    if (q_ok) {
        float* o = dqkv + ((long long)b * L + qi) * row + (long long)h * HD + g * DPT;
#pragma unroll
        for (int d = 0; d < DPT; ++d) o[d] = dq[d];
    }
}

// dK, dV: one key per G lanes, loop over queries.
template <typename T, int DPT, int G>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const T* __restrict__ qkv, const float* __restrict__ dout, const float* __restrict__ lse,
                    const float* __restrict__ delta, float* __restrict__ dqkv, int L, int C, int heads, float scale) {
    constexpr int HD = DPT * G;
    constexpr int KPB = 128 / G;
    constexpr int QT = (2048 / HD) > 32 ? 32 : (2048 / HD);
    __shared__ float Qs[QT * HD];
    __shared__ float Os[QT * HD];
    __shared__ float Ls[QT], Ds[QT];
    const int b = blockIdx.z, h = blockIdx.y;
    const int g = threadIdx.x % G;
    const int kj = blockIdx.x * KPB + threadIdx.x / G;
    const bool k_ok = kj < L;
    const long long row = 3LL * C;
    const T* base = qkv + (long long)b * L * row + (long long)h * HD;
    float k[DPT], v[DPT], dk[DPT], dv[DPT];
#pragma unroll
    for (int d = 0; d < DPT; ++d) {
        k[d] = k_ok ? to_f(base[(long long)kj * row + C + g * DPT + d]) : 0.f;
        v[d] = k_ok ? to_f(base[(long long)kj * row + 2 * C + g * DPT + d]) : 0.f;
        dk[d] = 0.f; dv[d] = 0.f;
    }
    for (int q0 = 0; q0 < L; q0 += QT) {
        __syncthreads();
        for (int i = threadIdx.x; i < QT * HD; i += 128) {
            const int j = i / HD, d = i - j * HD;
            const int qi = q0 + j;
            Qs[i] = qi < L ? to_f(base[(long long)qi * row + d]) : 0.f;
            Os[i] = qi < L ? dout[((long long)b * L + qi) * C + (long long)h * HD + d] : 0.f;
        }
        for (int j = threadIdx.x; j < QT; j += 128) {
            const int qi = q0 + j;
            Ls[j] = qi < L ? lse[((long long)b * heads + h) * L + qi] : 0.f;
            Ds[j] = qi < L ? delta[((long long)b * heads + h) * L + qi] : 0.f;
        }
        __syncthreads();
        const int qn = min(QT, L - q0);
        for (int j = 0; j < qn; ++j) {
            float s = 0.f, dp = 0.f;
            const float* qp = Qs + j * HD + g * DPT;
            const float* op = Os + j * HD + g * DPT;
#pragma unroll
            for (int d = 0; d < DPT; ++d) { s = fmaf(qp[d], k[d], s); dp = fmaf(op[d], v[d], dp); }
#pragma unroll
            for (int off = G >> 1; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); dp += __shfl_xor_sync(0xffffffffu, dp, off); }
            const float p = expf(s * scale - Ls[j]);
            const float ds = p * (dp - Ds[j]) * scale;
#pragma unroll
            for (int d = 0; d < DPT; ++d) { dv[d] = fmaf(p, op[d], dv[d]); dk[d] = fmaf(ds, qp[d], dk[d]); }
        }
    }
    if (k_ok) {
        float* o = dqkv + ((long long)b * L + kj) * row + (long long)h * HD + g * DPT;
#pragma unroll
        for (int d = 0; d < DPT; ++d) { o[C + d] = dk[d]; o[2 * C + d] = dv[d]; }
    }
}

template <typename T, int DPT, int G>
void launch(const void* qkv, void* out, float* lse, int B, int L, int C, int heads, float scale, cudaStream_t st) {
    constexpr int QPB = 128 / G;
    dim3 grid(ceil_div(L, QPB), heads, B);
    attn_simt_kernel<T, DPT, G><<<grid, 128, 0, st>>>((const T*)qkv, (T*)out, lse, L, C, heads, scale);
}

template <typename T, int DPT, int G>
void launch_bwd(const void* qkv, const float* dout, const float* lse, const float* delta, float* dqkv, int B, int L, int C,
                int heads, cudaStream_t st) {
    constexpr int QPB = 128 / G;
    const float scale = 1.f / sqrtf((float)(DPT * G));
    dim3 grid(ceil_div(L, QPB), heads, B);
    attn_bwd_dq_kernel<T, DPT, G><<<grid, 128, 0, st>>>((const T*)qkv, dout, lse, delta, dqkv, L, C, heads, scale);
    attn_bwd_dkv_kernel<T, DPT, G><<<grid, 128, 0, st>>>((const T*)qkv, dout, lse, delta, dqkv, L, C, heads, scale);
}

template <typename T>
int dispatch_bwd(const void* qkv, const void* o, const float* dout, const float* lse, float* delta, float* dqkv, int B, int L,
                 int C, int heads, cudaStream_t st) {
    const int hd = C / heads;
    attn_delta_kernel<T><<<dim3(ceil_div((long long)heads * L, 256), B), 256, 0, st>>>((const T*)o, dout, delta, L, C, heads);
    switch (hd) {
        case 8: launch_bwd<T, 8, 1>(qkv, dout, lse, delta, dqkv, B, L, C, heads, st); break;
        case 16: launch_bwd<T, 16, 1>(qkv, dout, lse, delta, dqkv, B, L, C, heads, st); break;
        case 32: launch_bwd<T, 16, 2>(qkv, dout, lse, delta, dqkv, B, L, C, heads, st); break;
        case 64: launch_bwd<T, 16, 4>(qkv, dout, lse, delta, dqkv, B, L, C, heads, st); break;
        case 128: launch_bwd<T, 16, 8>(qkv, dout, lse, delta, dqkv, B, L, C, heads, st); break;
        case 256: launch_bwd<T, 16, 16>(qkv, dout, lse, delta, dqkv, B, L, C, heads, st); break;
        default:
            ddpmir_set_error("attention_backward: unsupported head_dim %d", hd);
            return DDPMIR_ERR_UNSUPPORTED;
    }
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

template <typename T>
int dispatch(const void* qkv, void* out, float* lse, int B, int L, int C, int heads, float scale, cudaStream_t st) {
    const int hd = C / heads;
    if (scale <= 0.f) scale = 1.f / sqrtf((float)hd);   // default: the softmax scale; otherwise the factor still owed to q
    switch (hd) {
        case 8: launch<T, 8, 1>(qkv, out, lse, B, L, C, heads, scale, st); break;
        case 16: launch<T, 16, 1>(qkv, out, lse, B, L, C, heads, scale, st); break;
        case 32: launch<T, 16, 2>(qkv, out, lse, B, L, C, heads, scale, st); break;
        case 64: launch<T, 16, 4>(qkv, out, lse, B, L, C, heads, scale, st); break;
        case 128: launch<T, 16, 8>(qkv, out, lse, B, L, C, heads, scale, st); break;
        case 256: launch<T, 16, 16>(qkv, out, lse, B, L, C, heads, scale, st); break;
        default:
            ddpmir_set_error("attention: unsupported head_dim %d", hd);
            return DDPMIR_ERR_UNSUPPORTED;
    }
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

}  // namespace

int ddpmir_attention_simt(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, float qscale, cudaStream_t st) {
    if (dtype == DDPMIR_F32) return dispatch<float>(qkv, out, nullptr, B, L, C, heads, qscale, st);
    return dispatch<bf16>(qkv, out, nullptr, B, L, C, heads, qscale, st);
}

int ddpmir_attention_mma(const void* qkv, int B, int L, int C, int heads, void* out, float* lse, cudaStream_t st);

// training forward: same kernel, also writes the per-row log-sum-exp [B, heads, L] the backward needs
extern "C" int ddpmir_attention_train_forward(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, float* lse,
                                              ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(qkv && out && lse, "attention_train_forward: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && L > 0 && heads > 0 && C % heads == 0 && (C / heads) % 8 == 0, "attention_train_forward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_F32) return dispatch<float>(qkv, out, lse, B, L, C, heads, 0.f, st);
    const int rc = ddpmir_attention_mma(qkv, B, L, C, heads, out, lse, st);     // tensor-core kernel when the shape fits
    if (rc != DDPMIR_ERR_UNSUPPORTED) return rc;
    return dispatch<bf16>(qkv, out, lse, B, L, C, heads, 0.f, st);
}

// dqkv [B, L, 3C] fp32 from qkv, o (forward output), dout [B, L, C] fp32 and lse; delta: workspace [B, heads, L] fp32
int ddpmir_attention_backward_mma(const void* qkv, const void* dout_bf16, const float* lse, const float* delta, float* dqkv, int B,
                                  int L, int C, int heads, cudaStream_t st);

extern "C" int ddpmir_attention_backward(const void* qkv, const void* o, int dtype, const float* dout, const void* dout_op,
                                         const float* lse, float* delta, float* dqkv, int B, int L, int C, int heads,
                                         ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(qkv && o && dout && lse && delta && dqkv, "attention_backward: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && B <= 65535 && L > 0 && heads > 0 && C % heads == 0 && (C / heads) % 8 == 0, "attention_backward: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_BF16 && dout_op) {
        // tensor-core path: delta from the fp32 gradient, the matrix products on bf16 operands
        attn_delta_kernel<bf16><<<dim3(ceil_div((long long)heads * L, 256), B), 256, 0, st>>>((const bf16*)o, dout, delta, L, C, heads);
        const int rc = ddpmir_attention_backward_mma(qkv, dout_op, lse, delta, dqkv, B, L, C, heads, st);
        if (rc != DDPMIR_ERR_UNSUPPORTED) return rc;
    }
    if (dtype == DDPMIR_F32) return dispatch_bwd<float>(qkv, o, dout, lse, delta, dqkv, B, L, C, heads, st);
    return dispatch_bwd<bf16>(qkv, o, dout, lse, delta, dqkv, B, L, C, heads, st);
}
