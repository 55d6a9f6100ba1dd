// Microbenchmark (one-off): how many softmax elements/clk/SM can a "one thread = one score row" loop sustain with only
// NW warps per SM (the tcgen05 attention design: scores arrive in TMEM, no shuffles, no ldmatrix), for several
// MUFU : polynomial splits?  Work per element: FFMA (scale, subtract max) + exp2 + 1/2 pack + 1/2 FMNMX3.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float M = 12582912.f;
    const float t = x + M, n = t - M, f = x - n;
    const float p = fmaf(fmaf(fmaf(0.05517166f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack_trunc(float lo, float hi) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi))); return r; }

template <int POLY_OF_8>
__global__ void __launch_bounds__(256, 1) k(float* out, const float* in, int iters) {
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) s[i] = in[(threadIdx.x * 32 + i) & 1023];
    float m = -1e30f; uint32_t acc = 0;
    const float sc = 0.51f;
    for (int it = 0; it < iters; ++it) {
        float mx = m;
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(s[i], s[i + 1]));
        m = mx * 0.999f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float x0 = fmaf(s[i], sc, -m), x1 = fmaf(s[i + 1], sc, -m);
            float e0, e1;
            if (((i >> 1) & 7) < POLY_OF_8) { e0 = ex2_poly(x0); e1 = ex2_poly(x1); } else { e0 = ex2f(x0); e1 = ex2f(x1); }
            acc ^= pack_trunc(e0, e1);
            s[i] = e0 - 3.f; s[i + 1] = e1 - 3.f;     // feed back so nothing is hoisted
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = m + __uint_as_float(acc);
}
template <int P> void run(int warps_per_sm) {
    float *out, *in; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    const int iters = 20000, threads = warps_per_sm * 32;
    k<P><<<148, threads>>>(out, in, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<P><<<148, threads>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double elems = 148.0 * threads * iters * 32.0;
    printf("poly %d/8, %2d warps/SM: %7.3f ms  %6.2f elem/clk/SM @1965MHz\n", P, warps_per_sm, ms, elems / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out); cudaFree(in);
}
int main() {
    for (int w : {4, 8}) { run<0>(w); run<2>(w); run<3>(w); run<4>(w); }
    return 0;
}
