// Microbenchmark (one-off measurement, not part of the library): per-SM throughput of the exp2 variants the attention
// softmax can use.  Prints elements/clk/SM for: ex2.approx.ftz.f32, ex2.approx.ftz.bf16x2, ex2.approx.f16x2, an
// FMA-pipe polynomial exp2 (Cody-Waite + degree-3), and mixes.   nvcc -arch=sm_100a -O3 -o exp_mb exp_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float poly_ex2(float x) {
    // x <= 0.  round-to-nearest split with the 1.5*2^23 magic constant (FMA/ALU pipes only, no F2I/FRND):
    // t = x + M  -> low mantissa bits of t hold n = rint(x);  f = x - n in [-0.5, 0.5];  2^f by a degree-3 polynomial
    x = fmaxf(x, -125.f);
    const float M = 12582912.f;
    float t = x + M;
    float n = t - M;
    float f = x - n;
    float p = fmaf(fmaf(fmaf(0.05550411f, f, 0.24022651f), f, 0.69314718f), f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t cvt_pack(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t prmt_pack(float lo, float hi) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi))); return r; }

template <int MODE>
__global__ void k(float* out, int iters) {
    float a0 = threadIdx.x * -1e-3f, a1 = a0 - 0.1f, a2 = a0 - 0.2f, a3 = a0 - 0.3f;
    uint32_t u0 = 0xBF80BF00u + threadIdx.x, u1 = u0 + 7, u2 = u0 + 11, u3 = u0 + 13;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { a0 = ex2f(a0) - 1.5f; a1 = ex2f(a1) - 1.5f; a2 = ex2f(a2) - 1.5f; a3 = ex2f(a3) - 1.5f; }
        if (MODE == 1) { u0 = ex2_bf16x2(u0) ^ 0x80008000u; u1 = ex2_bf16x2(u1) ^ 0x80008000u; u2 = ex2_bf16x2(u2) ^ 0x80008000u; u3 = ex2_bf16x2(u3) ^ 0x80008000u; }
        if (MODE == 2) { u0 = ex2_f16x2(u0) ^ 0x80008000u; u1 = ex2_f16x2(u1) ^ 0x80008000u; u2 = ex2_f16x2(u2) ^ 0x80008000u; u3 = ex2_f16x2(u3) ^ 0x80008000u; }
        if (MODE == 3) { a0 = poly_ex2(a0) - 1.5f; a1 = poly_ex2(a1) - 1.5f; a2 = poly_ex2(a2) - 1.5f; a3 = poly_ex2(a3) - 1.5f; }
        if (MODE == 4) { a0 = ex2f(a0) - 1.5f; a1 = ex2f(a1) - 1.5f; a2 = ex2f(a2) - 1.5f; a3 = poly_ex2(a3) - 1.5f; }   // 3:1 mix
        if (MODE == 5) { a0 = ex2f(a0) - 1.5f; a1 = poly_ex2(a1) - 1.5f; a2 = ex2f(a2) - 1.5f; a3 = poly_ex2(a3) - 1.5f; } // 1:1 mix
        if (MODE == 6) { u0 = cvt_pack(a0, __uint_as_float(u0)); u1 = cvt_pack(a1, __uint_as_float(u1)); u2 = cvt_pack(a2, __uint_as_float(u2)); u3 = cvt_pack(a3, __uint_as_float(u3)); }
        if (MODE == 7) { u0 = prmt_pack(a0, __uint_as_float(u0)); u1 = prmt_pack(a1, __uint_as_float(u1)); u2 = prmt_pack(a2, __uint_as_float(u2)); u3 = prmt_pack(a3, __uint_as_float(u3)); }
        if (MODE == 8) { a0 = fmaxf(a0, a1 + 1e-3f * i); a1 = fmaxf(a1, a2); a2 = fmaxf(a2, a3); a3 = fmaxf(a3, a0); }
        if (MODE == 9) {   // softmax inner step on 4 scores: FFMA + ex2 each, 2 cvt packs, 4 max
            float e0 = ex2f(fmaf(a0, 0.7f, -1.f)), e1 = ex2f(fmaf(a1, 0.7f, -1.f)), e2 = ex2f(fmaf(a2, 0.7f, -1.f)), e3 = ex2f(fmaf(a3, 0.7f, -1.f));
            u0 ^= cvt_pack(e0, e1); u1 ^= cvt_pack(e2, e3);
            a0 = fmaxf(a0, e1) - 1.f; a1 = fmaxf(a1, e2) - 1.f; a2 = fmaxf(a2, e3) - 1.f; a3 = fmaxf(a3, e0) - 1.f;
        }
        if (MODE == 10) {  // same with PRMT truncation packing
            float e0 = ex2f(fmaf(a0, 0.7f, -1.f)), e1 = ex2f(fmaf(a1, 0.7f, -1.f)), e2 = ex2f(fmaf(a2, 0.7f, -1.f)), e3 = ex2f(fmaf(a3, 0.7f, -1.f));
            u0 ^= prmt_pack(e0, e1); u1 ^= prmt_pack(e2, e3);
            a0 = fmaxf(a0, e1) - 1.f; a1 = fmaxf(a1, e2) - 1.f; a2 = fmaxf(a2, e3) - 1.f; a3 = fmaxf(a3, e0) - 1.f;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
}

template <int MODE>
void run(const char* name, int elems_per_op) {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    const int iters = 20000, blocks = 148 * 2, threads = 1024;
    k<MODE><<<blocks, threads>>>(out, 100);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double elems = (double)blocks * threads * iters * 4.0 * elems_per_op;
    printf("%-28s %8.3f ms  %7.2f Gelem/s  (~%5.1f elem/clk/SM at %d MHz nominal)\n", name, ms, elems / ms / 1e6,
           elems / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("ex2.approx.ftz.f32", 1);
    run<1>("ex2.approx.ftz.bf16x2", 2);
    run<2>("ex2.approx.f16x2", 2);
    run<3>("poly exp2 (FMA pipe)", 1);
    run<4>("3 MUFU : 1 poly", 1);
    run<5>("1 MUFU : 1 poly", 1);
    run<6>("cvt.rn.bf16x2.f32 (per instr)", 1);
    run<7>("prmt pack (per instr)", 1);
    run<8>("fmax (FMNMX)", 1);
    run<9>("softmax step, cvt pack", 1);
    run<10>("softmax step, prmt pack", 1);
    return 0;
}
