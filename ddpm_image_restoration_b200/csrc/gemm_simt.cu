// Generic fp32-FMA implicit GEMM for 3x3 convolutions (TAPS = 9) and 1x1 convolutions / linears (TAPS = 1) on
// NHWC activations.  This is the fp32 check-mode kernel (DDPMIR_F32) and the any-shape fallback of the bf16
// path; the tcgen05 kernels in conv_tc.cu take over whenever their shape constraints hold.
#include "epilogue.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T> struct Ld4;
template <> struct Ld4<float> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
};
template <> struct Ld4<bf16> {
    static __device__ __forceinline__ void ld(const bf16* p, float (&v)[4]) {
        uint2 r = *reinterpret_cast<const uint2*>(p);
        float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.x));
        float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};

template <typename T, int TAPS>
__global__ void __launch_bounds__(256)
igemm_simt_kernel(const T* __restrict__ x, const T* __restrict__ w, void* __restrict__ out, EpiDev ep, long long M,
                  int H, int W, int Cin, int N) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = TAPS * Cin;

    // loader role: one row of the A tile and one row of the B tile, 4 consecutive k each
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const long long lm = m0 + lr;
    int pb = 0, ph = 0, pw = 0;
    const bool m_ok = lm < M;
    if (m_ok) {
        const int hw = H * W;
        pb = (int)(lm / hw);
        const int rem = (int)(lm - (long long)pb * hw);
        ph = rem / W; pw = rem - ph * W;
    }
    const int ln = n0 + lr;
    const bool n_ok = ln < N;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const int tap = k0 / Cin;
            const int c = k0 - tap * Cin + lk;
            int hh = ph, ww = pw;
            if (TAPS == 9) { hh += tap / 3 - 1; ww += tap % 3 - 1; }
            if (m_ok && hh >= 0 && hh < H && ww >= 0 && ww < W)
                Ld4<T>::ld(x + (((long long)pb * H + hh) * W + ww) * Cin + c, av);
            if (n_ok) Ld4<T>::ld(w + (long long)ln * K + k0 + lk, bv);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) { As[lk + i][lr] = av[i]; Bs[lk + i][lr] = bv[i]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        const EpiRow row = epi_row(ep, m);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) epi_store(ep, out, m, n, epi_apply(ep, row, acc[i][j], m, n));
        }
    }
}

template <int TAPS>
int launch(const void* x, int dtype, int B, int H, int W, int Cin, const void* w, int N, const ddpmir_epilogue_t* epi,
           void* out, cudaStream_t st) {
    const long long M = (long long)B * H * W;
    EpiDev ep = make_epi(epi, H, W, N, dtype);
    dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
    if (dtype == DDPMIR_F32)
        igemm_simt_kernel<float, TAPS><<<grid, 256, 0, st>>>((const float*)x, (const float*)w, out, ep, M, H, W, Cin, N);
    else
        igemm_simt_kernel<bf16, TAPS><<<grid, 256, 0, st>>>((const bf16*)x, (const bf16*)w, out, ep, M, H, W, Cin, N);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

}  // namespace

int ddpmir_conv3x3_simt(const void* x, int dtype, int B, int H, int W, int Cin, const void* w, int N,
                        const ddpmir_epilogue_t* epi, void* out, cudaStream_t st) {
    return launch<9>(x, dtype, B, H, W, Cin, w, N, epi, out, st);
}
int ddpmir_gemm_simt(const void* a, int dtype, int B, int H, int W, int K, const void* w, int N,
                     const ddpmir_epilogue_t* epi, void* out, cudaStream_t st) {
    return launch<1>(a, dtype, B, H, W, K, w, N, epi, out, st);
}
