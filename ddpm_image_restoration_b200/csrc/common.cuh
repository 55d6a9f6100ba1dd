// Shared helpers for the ddpmir CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ddpmir.h"

typedef __nv_bfloat16 bf16;

// ---- error reporting across the C ABI (no exceptions; last message kept per thread) ----------------------
void ddpmir_set_error(const char* fmt, ...);

#define DDPMIR_CHECK_ARG(cond, ...)                 \
    do {                                            \
        if (!(cond)) {                              \
            ddpmir_set_error(__VA_ARGS__);          \
            return DDPMIR_ERR_INVALID;              \
        }                                           \
    } while (0)

#define DDPMIR_LAUNCH_CHECK()                                                        \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ddpmir_set_error("%s:%d launch failed: %s", __FILE__, __LINE__,          \
                             cudaGetErrorString(e__));                               \
            return DDPMIR_ERR_CUDA;                                                  \
        }                                                                            \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- scalar conversions ---------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 8-wide channel vectors (the natural NHWC access unit: 16 B of bf16 / 32 B of fp32) -----------------
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    float v[8];
    __device__ __forceinline__ void load(const float* p) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<bf16> {
    float v[8];
    __device__ __forceinline__ void load(const bf16* p) {
        uint4 r = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void store(bf16* p) const {
        uint4 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: remember the opt-in per device, not per process
// (a process that moves to a second GPU would otherwise launch there without it).  Returns the slot for the current device.
struct PerDevice {
    int v[64] = {0};
    int& cur() {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & 63];
    }
};

// ---- activations ----------------------------------------------------------------------------------------
// exact (erf) GELU, F.gelu default (webp_inference.py:312); SiLU (webp_inference.py:364)
__device__ __forceinline__ float act_apply(int act, float v) {
    switch (act) {
        case DDPMIR_ACT_RELU: return v > 0.f ? v : 0.f;
        case DDPMIR_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
        case DDPMIR_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case DDPMIR_ACT_SILU: return v / (1.f + expf(-v));
        case DDPMIR_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case DDPMIR_ACT_TANH: return tanhf(v);
        default: return v;
    }
}

// low-frequency membership of pixel (h,w): restates the block loop of WebPFreqAwareBlock.forward
// (webp_inference.py:241-252) incl. ragged edge blocks: low_size = max(1, min(low, rows_left, cols_left)).
__device__ __forceinline__ bool is_low_freq(int h, int w, int H, int W, int bs, int low) {
    int rl = H - (h / bs) * bs; rl = rl < bs ? rl : bs;
    int cl = W - (w / bs) * bs; cl = cl < bs ? cl : bs;
    int ls = rl < cl ? rl : cl;
    ls = ls < low ? ls : low;
    ls = ls > 1 ? ls : 1;
    return (h % bs) < ls && (w % bs) < ls;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- geometry shared by the AVIF gate kernels (forward: spatial.cu, backward: backward_avif.cu) ----------------------------
// cell index of the [85, B, C] pyramid (s = 1, 2, 4, 8 -> bases 0, 1, 5, 21) -> scale and cell coordinates
__device__ __forceinline__ void pyramid_cell(int cell, int& s, int& ci, int& cj) {
    int base;
    if (cell < 1) { s = 1; base = 0; }
    else if (cell < 5) { s = 2; base = 1; }
    else if (cell < 21) { s = 4; base = 5; }
    else { s = 8; base = 21; }
    const int k = cell - base;
    ci = k / s; cj = k % s;
}

__device__ __forceinline__ void bilin_src(int d, int n_in, int n_out, int& i0, int& i1, float& lam) {
    // F.interpolate(size=...), mode='bilinear', align_corners=False: scale = n_in / n_out
    float s = ((float)d + 0.5f) * ((float)n_in / (float)n_out) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = s - (float)i0;
}
