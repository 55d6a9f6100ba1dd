// Epilogue shared by every GEMM-shaped kernel (SIMT and tensor-core): see ddpmir_epilogue_t in ddpmir.h.
#pragma once
#include "common.cuh"

struct EpiDev {
    const float* bias;
    const float* bias2;
    const float* row_bias;
    const float* img_scale;
    const void* mul;
    const void* res;
    void* out2;
    int act;
    int freq_mode;
    int bs;
    int low;
    int out_dtype, out2_dtype, mul_dtype, res_dtype;
    int H, W, N;
};

__device__ __forceinline__ float ld_any(const void* p, int dtype, long long i) {
    return dtype == DDPMIR_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, long long i, float v) {
    if (dtype == DDPMIR_F32) reinterpret_cast<float*>(p)[i] = v;
    else if (dtype == DDPMIR_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

static inline EpiDev make_epi(const ddpmir_epilogue_t* e, int H, int W, int N, int operand_dtype) {
    EpiDev d;
    d.out2 = e ? e->out2 : nullptr;
    d.out_dtype = e ? e->out_dtype : operand_dtype;
    d.out2_dtype = e ? e->out2_dtype : operand_dtype;
    d.mul_dtype = e ? e->mul_dtype : operand_dtype;
    d.res_dtype = e ? e->res_dtype : operand_dtype;
    d.bias = e ? e->bias : nullptr;
    d.bias2 = e ? e->bias2 : nullptr;
    d.row_bias = e ? e->row_bias : nullptr;
    d.img_scale = e ? e->img_scale : nullptr;
    d.mul = e ? e->mul : nullptr;
    d.res = e ? e->res : nullptr;
    d.act = e ? e->act : 0;
    d.freq_mode = e ? e->freq_mode : 0;
    d.bs = e ? e->bs : 0;
    d.low = e ? e->low : 0;
    d.H = H; d.W = W; d.N = N;
    return d;
}

static inline int check_epi(const ddpmir_epilogue_t* e, int N) {
    if (!e) return DDPMIR_OK;
    if (e->freq_mode < 0 || e->freq_mode > 2) { ddpmir_set_error("epilogue: freq_mode %d", e->freq_mode); return DDPMIR_ERR_INVALID; }
    if (e->freq_mode && (e->bs <= 0 || e->low <= 0)) { ddpmir_set_error("epilogue: freq_mode needs bs/low"); return DDPMIR_ERR_INVALID; }
    if (e->freq_mode == 1 && (N & 1)) { ddpmir_set_error("epilogue: freq_mode 1 needs even N"); return DDPMIR_ERR_INVALID; }
    if (e->freq_mode == 2 && (!e->bias2 || !e->bias)) { ddpmir_set_error("epilogue: freq_mode 2 needs bias and bias2"); return DDPMIR_ERR_INVALID; }
    const int dts[4] = {e->out_dtype, e->out2_dtype, e->mul_dtype, e->res_dtype};
    for (int i = 0; i < 4; ++i)     // binary16 only as an OUTPUT format (the qkv of the half-precision attention tier)
        if (dts[i] != DDPMIR_F32 && dts[i] != DDPMIR_BF16 && !(i < 2 && dts[i] == DDPMIR_F16)) {
            ddpmir_set_error("epilogue: bad dtype field %d", dts[i]);
            return DDPMIR_ERR_INVALID;
        }
    return DDPMIR_OK;
}

// Per-row (pixel) context, computed once per output row.
struct EpiRow {
    int b;
    bool low;
    float scale;  // per-image scale to apply (already resolved for freq_mode 2)
};

__device__ __forceinline__ EpiRow epi_row(const EpiDev& p, long long m) {
    EpiRow r;
    const int hw = p.H * p.W;
    r.b = (int)(m / hw);
    r.low = false;
    if (p.freq_mode) {
        const int rem = (int)(m - (long long)r.b * hw);
        const int h = rem / p.W, w = rem - h * p.W;
        r.low = is_low_freq(h, w, p.H, p.W, p.bs, p.low);
    }
    r.scale = p.img_scale ? p.img_scale[r.b] : 1.f;
    if (p.freq_mode == 2 && r.low) r.scale = 1.f;
    return r;
}

__device__ __forceinline__ float epi_apply(const EpiDev& p, const EpiRow& r, float acc, long long m, int n) {
    float v = acc;
    if (p.freq_mode == 2 && !r.low) v += p.bias2[n];
    else if (p.bias) v += p.bias[n];
    if (p.row_bias) v += p.row_bias[(long long)r.b * p.N + n];
    v = act_apply(p.act, v);
    if (p.freq_mode == 1) {
        if ((n < (p.N >> 1)) != r.low) v = 0.f;
    }
    v *= r.scale;
    if (p.mul) v *= ld_any(p.mul, p.mul_dtype, m * p.N + n);
    if (p.res) v += ld_any(p.res, p.res_dtype, m * p.N + n);
    return v;
}

__device__ __forceinline__ void epi_store(const EpiDev& p, void* out, long long m, int n, float v) {
    st_any(out, p.out_dtype, m * p.N + n, v);
    if (p.out2) st_any(p.out2, p.out2_dtype, m * p.N + n, v);
}

// ---- vectorised epilogue over 16 consecutive output channels of one row (tcgen05 kernel: one TMEM lane = one row) ----
__device__ __forceinline__ void ld16_f32(const float* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 a = reinterpret_cast<const float4*>(p)[i];
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
}
__device__ __forceinline__ void ld16_any(const void* base, int dtype, long long idx, float (&v)[16]) {
    if (dtype == DDPMIR_F32) {
        ld16_f32(reinterpret_cast<const float*>(base) + idx, v);
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + idx);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint4 r = p[i];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(h[k]);
                v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y;
            }
        }
    }
}
__device__ __forceinline__ void st16_any(void* base, int dtype, long long idx, const float (&v)[16]) {
    if (dtype == DDPMIR_F32) {
        float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if (dtype == DDPMIR_F16) {
        uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(base) + idx);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint4 r;
            __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
            p[i] = r;
        }
    } else {
        uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + idx);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            uint4 r;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[8 * i + 2 * k], v[8 * i + 2 * k + 1]);
            p[i] = r;
        }
    }
}

// v[16] = accumulators of row m, channels n .. n+15 (n % 16 == 0, n + 16 <= N, N % 16 == 0)
__device__ __forceinline__ void epi_chunk16(const EpiDev& p, const EpiRow& r, float (&v)[16], long long m, int n, void* out) {
    float t[16];
    const float* bias = (p.freq_mode == 2 && !r.low) ? p.bias2 : p.bias;
    if (bias) {
        ld16_f32(bias + n, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += t[j];
    }
    if (p.row_bias) {
        ld16_f32(p.row_bias + (long long)r.b * p.N + n, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += t[j];
    }
    switch (p.act) {
        case DDPMIR_ACT_RELU:
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            break;
        case DDPMIR_ACT_LRELU02:
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
            break;
        case DDPMIR_ACT_SIGMOID:
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
            break;
        case DDPMIR_ACT_NONE:
            break;
        default:
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_apply(p.act, v[j]);
    }
    if (p.freq_mode == 1 && ((n < (p.N >> 1)) != r.low)) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    if (r.scale != 1.f) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= r.scale;
    }
    const long long idx = m * p.N + n;
    if (p.mul) {
        ld16_any(p.mul, p.mul_dtype, idx, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= t[j];
    }
    if (p.res) {
        ld16_any(p.res, p.res_dtype, idx, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += t[j];
    }
    st16_any(out, p.out_dtype, idx, v);
    if (p.out2) st16_any(p.out2, p.out2_dtype, idx, v);
}
