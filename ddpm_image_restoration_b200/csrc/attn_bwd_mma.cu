// Tensor-core (mma.sync bf16) backward of the self-attention core for the training step.
//
// Two launches that mirror the forward kernel's structure (rows owned by warps, 64-wide column tiles streamed through
// shared memory with cp.async) so that no transposes through shared memory and no atomics are needed:
//   MODE 0 (dQ):    rows = queries.  S = Q K^T, P = exp(S - lse_row), dP = dO V^T, dS = P (dP - delta_row) scale,
//                   dQ += dS K.
//   MODE 1 (dK,dV): rows = keys.     S^T = K Q^T, P^T = exp(S^T - lse_col), dP^T = V dO^T, dS^T = P^T (dP^T - delta_col) scale,
//                   dV += P^T dO,  dK += dS^T Q.
// In both modes the row operands (R1 for the score product, R2 for the dP product) sit in registers as A fragments, the
// column tiles T1 / T2 are the B operands (ldmatrix for the products over head_dim, ldmatrix.trans for the products
// over the tile's 64 columns).  P is recomputed from the saved log-sum-exp (flash-attention style), so the L x L
// matrices never exist.  head_dim 8 uses the k=8 MMA for the head_dim products like the forward.
#include "common.cuh"

namespace {

constexpr int KT = 64;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <int CPR> __device__ __forceinline__ int swz(int row, int chunk) {
    if (CPR == 1) return chunk;
    constexpr int DIV = (8 / CPR) > 1 ? (8 / CPR) : 1;
    constexpr int MOD = CPR < 8 ? CPR : 8;
    return chunk ^ ((row / DIV) % MOD);
}

struct BwdParams {
    const bf16* r1; long long r1_stride;     // row operand of the score product   (Q | K), per-(b) base handled in-kernel
    const bf16* r2; long long r2_stride;     // row operand of the dP product      (dO | V)
    const bf16* t1; long long t1_stride;     // column tile of the score product   (K | Q)
    const bf16* t2; long long t2_stride;     // column tile of the dP product      (V | dO)
    const float* lse; const float* delta;    // [B, heads, L]
    float* out1; float* out2;                // dQ | dK, (unused) | dV   -- fp32, row stride out_stride
    long long out_stride;
    int L, heads;
    float scale, scale_log2;
};

template <int HD, int MODE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
attn_bwd_mma_kernel(BwdParams p) {
    constexpr int CPR = HD / 8;
    constexpr int NT = NWARPS * 32;
    constexpr int ROWS = NWARPS * 16;
    constexpr int NDT = HD / 8;
    constexpr int KSTEPS = HD >= 16 ? HD / 16 : 1;
    constexpr int TILE = KT * HD;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    bf16* T1s = reinterpret_cast<bf16*>(smem_raw);        // [2][KT][HD]
    bf16* T2s = T1s + 2 * TILE;                           // [2][KT][HD]
    float* Ls = reinterpret_cast<float*>(T2s + 2 * TILE); // [2][KT]  (MODE 1: lse of the column tile)
    float* Ds = Ls + 2 * KT;                              // [2][KT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z, h = blockIdx.y;
    const int L = p.L;
    const long long bh = (long long)b * p.heads + h;
    const bf16* r1 = p.r1 + (long long)b * L * p.r1_stride + (long long)h * HD;
    const bf16* r2 = p.r2 + (long long)b * L * p.r2_stride + (long long)h * HD;
    const bf16* t1 = p.t1 + (long long)b * L * p.t1_stride + (long long)h * HD;
    const bf16* t2 = p.t2 + (long long)b * L * p.t2_stride + (long long)h * HD;
    const float* lse = p.lse + bh * L;
    const float* delta = p.delta + bh * L;

    auto load_tile = [&](int t, int stage) {
        const int c0 = t * KT;
        for (int i = tid; i < KT * CPR; i += NT) {
            const int row = i / CPR, ch = i - row * CPR;
            const int so = stage * TILE + row * HD + swz<CPR>(row, ch) * 8;
            cp_async16(T1s + so, t1 + (long long)(c0 + row) * p.t1_stride + ch * 8);
            cp_async16(T2s + so, t2 + (long long)(c0 + row) * p.t2_stride + ch * 8);
        }
        if (MODE == 1)
            for (int i = tid; i < KT; i += NT) {
                cp_async4(Ls + stage * KT + i, lse + c0 + i);
                cp_async4(Ds + stage * KT + i, delta + c0 + i);
            }
    };

    // row fragments (A operands) straight from global memory
    const int r_lo = lane >> 2, c_lo = (lane & 3) * 2;
    const int row0 = blockIdx.x * ROWS + warp * 16;
    const int ra = row0 + r_lo, rb = ra + 8;
    const int rac = ra < L ? ra : L - 1, rbc = rb < L ? rb : L - 1;
    uint32_t f1[KSTEPS][4], f2[KSTEPS][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
        const bf16* a1 = r1 + (long long)rac * p.r1_stride + ks * 16 + c_lo;
        const bf16* b1 = r1 + (long long)rbc * p.r1_stride + ks * 16 + c_lo;
        const bf16* a2 = r2 + (long long)rac * p.r2_stride + ks * 16 + c_lo;
        const bf16* b2 = r2 + (long long)rbc * p.r2_stride + ks * 16 + c_lo;
        f1[ks][0] = *reinterpret_cast<const uint32_t*>(a1); f1[ks][1] = *reinterpret_cast<const uint32_t*>(b1);
        f2[ks][0] = *reinterpret_cast<const uint32_t*>(a2); f2[ks][1] = *reinterpret_cast<const uint32_t*>(b2);
        if (HD >= 16) {
            f1[ks][2] = *reinterpret_cast<const uint32_t*>(a1 + 8); f1[ks][3] = *reinterpret_cast<const uint32_t*>(b1 + 8);
            f2[ks][2] = *reinterpret_cast<const uint32_t*>(a2 + 8); f2[ks][3] = *reinterpret_cast<const uint32_t*>(b2 + 8);
        } else {
            f1[ks][2] = f1[ks][3] = f2[ks][2] = f2[ks][3] = 0u;
        }
    }
    float lse_a = 0.f, lse_b = 0.f, del_a = 0.f, del_b = 0.f;
    if (MODE == 0) {
        lse_a = lse[rac] * 1.4426950408889634f; lse_b = lse[rbc] * 1.4426950408889634f;
        del_a = delta[rac]; del_b = delta[rbc];
    }

    float o1[NDT][4], o2[NDT][4];
#pragma unroll
    for (int dt = 0; dt < NDT; ++dt)
#pragma unroll
        for (int i = 0; i < 4; ++i) { o1[dt][i] = 0.f; o2[dt][i] = 0.f; }

    const int ntiles = L / KT;
    load_tile(0, 0);
    cp_async_commit();
    const uint32_t t1_s = (uint32_t)__cvta_generic_to_shared(T1s), t2_s = (uint32_t)__cvta_generic_to_shared(T2s);
    const int lrow = lane & 7, lmat = lane >> 3;

    for (int t = 0; t < ntiles; ++t) {
        if (t + 1 < ntiles) { load_tile(t + 1, (t + 1) & 1); cp_async_commit(); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const uint32_t s1 = t1_s + (t & 1) * TILE * 2, s2 = t2_s + (t & 1) * TILE * 2;

        // ---- S = R1 T1^T and dP = R2 T2^T over head_dim ----------------------------------------------------------
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) { s[nt][i] = 0.f; dp[nt][i] = 0.f; }
        if (HD >= 16) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int row = (2 * q + (lmat >> 1)) * 8 + lrow;
                    const int ch = 2 * ks + (lmat & 1);
                    const uint32_t off = (row * HD + swz<CPR>(row, ch) * 8) * 2;
                    uint32_t kf[4], vf[4];
                    ldsm_x4(kf, s1 + off);
                    ldsm_x4(vf, s2 + off);
                    mma_16816(s[2 * q], f1[ks], kf[0], kf[1]);
                    mma_16816(s[2 * q + 1], f1[ks], kf[2], kf[3]);
                    mma_16816(dp[2 * q], f2[ks], vf[0], vf[1]);
                    mma_16816(dp[2 * q + 1], f2[ks], vf[2], vf[3]);
                }
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int row = (4 * q + lmat) * 8 + lrow;
                uint32_t kf[4], vf[4];
                ldsm_x4(kf, s1 + row * HD * 2);
                ldsm_x4(vf, s2 + row * HD * 2);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_1688(s[4 * q + j], f1[0][0], f1[0][1], kf[j]);
                    mma_1688(dp[4 * q + j], f2[0][0], f2[0][1], vf[j]);
                }
            }
        }

        // ---- P = exp(S scale - lse), dS = P (dP - delta) scale; packed as A operands of the 64-column products -----
        uint32_t pf[4][4], dsf[4][4];
        const float* lcol = Ls + (t & 1) * KT;
        const float* dcol = Ds + (t & 1) * KT;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float l0, l1, l2, l3, d0, d1, d2, d3;
            if (MODE == 0) {
                l0 = l1 = lse_a; l2 = l3 = lse_b; d0 = d1 = del_a; d2 = d3 = del_b;
            } else {
                const int c = nt * 8 + c_lo;
                l0 = l2 = lcol[c] * 1.4426950408889634f; l1 = l3 = lcol[c + 1] * 1.4426950408889634f;
                d0 = d2 = dcol[c]; d1 = d3 = dcol[c + 1];
            }
            const float p0 = ex2f(fmaf(s[nt][0], p.scale_log2, -l0)), p1 = ex2f(fmaf(s[nt][1], p.scale_log2, -l1));
            const float p2 = ex2f(fmaf(s[nt][2], p.scale_log2, -l2)), p3 = ex2f(fmaf(s[nt][3], p.scale_log2, -l3));
            const float e0 = p0 * (dp[nt][0] - d0) * p.scale, e1 = p1 * (dp[nt][1] - d1) * p.scale;
            const float e2 = p2 * (dp[nt][2] - d2) * p.scale, e3 = p3 * (dp[nt][3] - d3) * p.scale;
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
            dsf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(e0, e1);
            dsf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
        }

        // ---- out1 += dS T1 ; (MODE 1) out2 += P T2 : products over the tile's 64 columns ----------------------------
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (HD >= 16) {
#pragma unroll
                for (int a = 0; a < NDT / 2; ++a) {
                    const int row = kk * 16 + (lmat & 1) * 8 + lrow;
                    const int ch = 2 * a + (lmat >> 1);
                    const uint32_t off = (row * HD + swz<CPR>(row, ch) * 8) * 2;
                    uint32_t kf[4];
                    ldsm_x4_t(kf, s1 + off);
                    mma_16816(o1[2 * a], dsf[kk], kf[0], kf[1]);
                    mma_16816(o1[2 * a + 1], dsf[kk], kf[2], kf[3]);
                    if (MODE == 1) {
                        uint32_t vf[4];
                        ldsm_x4_t(vf, s2 + off);
                        mma_16816(o2[2 * a], pf[kk], vf[0], vf[1]);
                        mma_16816(o2[2 * a + 1], pf[kk], vf[2], vf[3]);
                    }
                }
            } else {
                const int row = kk * 16 + (lmat & 1) * 8 + lrow;
                uint32_t kf[2];
                ldsm_x2_t(kf, s1 + row * HD * 2);
                mma_16816(o1[0], dsf[kk], kf[0], kf[1]);
                if (MODE == 1) {
                    uint32_t vf[2];
                    ldsm_x2_t(vf, s2 + row * HD * 2);
                    mma_16816(o2[0], pf[kk], vf[0], vf[1]);
                }
            }
        }
        __syncthreads();
    }

    // ---- store fp32 gradients -------------------------------------------------------------------------------------------
    float* oa1 = p.out1 + ((long long)b * L + ra) * p.out_stride + (long long)h * HD + c_lo;
    float* ob1 = p.out1 + ((long long)b * L + rb) * p.out_stride + (long long)h * HD + c_lo;
#pragma unroll
    for (int dt = 0; dt < NDT; ++dt) {
        if (ra < L) *reinterpret_cast<float2*>(oa1 + dt * 8) = make_float2(o1[dt][0], o1[dt][1]);
        if (rb < L) *reinterpret_cast<float2*>(ob1 + dt * 8) = make_float2(o1[dt][2], o1[dt][3]);
    }
    if (MODE == 1) {
        float* oa2 = p.out2 + ((long long)b * L + ra) * p.out_stride + (long long)h * HD + c_lo;
        float* ob2 = p.out2 + ((long long)b * L + rb) * p.out_stride + (long long)h * HD + c_lo;
#pragma unroll
        for (int dt = 0; dt < NDT; ++dt) {
            if (ra < L) *reinterpret_cast<float2*>(oa2 + dt * 8) = make_float2(o2[dt][0], o2[dt][1]);
            if (rb < L) *reinterpret_cast<float2*>(ob2 + dt * 8) = make_float2(o2[dt][2], o2[dt][3]);
        }
    }
}

template <int HD, int MODE>
int launch(const BwdParams& p, int B, cudaStream_t st) {
    constexpr int NW = 4;
    const size_t smem = (size_t)4 * KT * HD * sizeof(bf16) + 4 * KT * sizeof(float);
    auto kern = attn_bwd_mma_kernel<HD, MODE, NW>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ddpmir_set_error("attention_backward: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
    }
    dim3 grid(ceil_div(p.L, NW * 16), p.heads, B);
    kern<<<grid, NW * 32, smem, st>>>(p);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

template <int HD>
int run(const bf16* qkv, const bf16* dout, const float* lse, const float* delta, float* dqkv, int B, int L, int C, int heads,
        cudaStream_t st) {
    BwdParams p;
    p.lse = lse; p.delta = delta; p.L = L; p.heads = heads;
    p.scale = 1.f / sqrtf((float)HD); p.scale_log2 = p.scale * 1.4426950408889634f;
    p.out_stride = 3LL * C;
    // dQ: rows = queries (Q, dO); column tiles K, V
    p.r1 = qkv; p.r1_stride = 3LL * C; p.r2 = dout; p.r2_stride = C;
    p.t1 = qkv + C; p.t1_stride = 3LL * C; p.t2 = qkv + 2 * C; p.t2_stride = 3LL * C;
    p.out1 = dqkv; p.out2 = nullptr;
    int rc = launch<HD, 0>(p, B, st);
    if (rc) return rc;
    // dK, dV: rows = keys (K, V); column tiles Q, dO
    p.r1 = qkv + C; p.r1_stride = 3LL * C; p.r2 = qkv + 2 * C; p.r2_stride = 3LL * C;
    p.t1 = qkv; p.t1_stride = 3LL * C; p.t2 = dout; p.t2_stride = C;
    p.out1 = dqkv + C; p.out2 = dqkv + 2 * C;
    return launch<HD, 1>(p, B, st);
}

}  // namespace

// bf16 qkv [B,L,3C], bf16 dout [B,L,C]; lse / delta [B,heads,L] fp32; dqkv [B,L,3C] fp32.  L % 64 == 0.
int ddpmir_attention_backward_mma(const void* qkv, const void* dout_bf16, const float* lse, const float* delta, float* dqkv, int B,
                                  int L, int C, int heads, cudaStream_t st) {
    if (L % KT != 0) return DDPMIR_ERR_UNSUPPORTED;
    const bf16* q = (const bf16*)qkv;
    const bf16* d = (const bf16*)dout_bf16;
    switch (C / heads) {
        case 8: return run<8>(q, d, lse, delta, dqkv, B, L, C, heads, st);
        case 16: return run<16>(q, d, lse, delta, dqkv, B, L, C, heads, st);
        case 32: return run<32>(q, d, lse, delta, dqkv, B, L, C, heads, st);
        case 64: return run<64>(q, d, lse, delta, dqkv, B, L, C, heads, st);
        default: return DDPMIR_ERR_UNSUPPORTED;
    }
}
