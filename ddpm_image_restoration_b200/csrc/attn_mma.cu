// Fused (flash-style) self-attention for the UNet's full-resolution attention blocks: bf16 operands on the
// tensor cores (mma.sync m16n8k16 / m16n8k8, fp32 accumulate), online softmax in registers, K/V tiles streamed
// through shared memory with cp.async double buffering.  The L x L score matrix is never materialised
// (the reference's nn.MultiheadAttention does, webp_inference.py:317-319, 68.7 GB per image at 256x256).
//
// With head_dim 8/16 this kernel is bound by the exp (MUFU) and FMA pipes, not by the tensor pipe: one 16x64
// score tile costs 16 HMMA but 1024 exp2 + ~3 FP32 ops per score.  Hence the design choices below:
//   * softmax scale folded into one FFMA per score (exp2 domain), row sums taken by the tensor core
//     (P times a constant ones-column fragment) instead of 1 FADD per score,
//   * optional packed bf16x2 exp2 (EXPMODE 1): two scores per MUFU op,
//   * P stays in registers as the A operand of the PV product (no shared-memory round trip).
#include "common.cuh"

namespace {

constexpr int KT = 64;  // keys per shared-memory tile

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
// D += A(16x16) * B(16x8)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D += A(16x8) * B(8x8)
__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// truncating pack: keeps the high halves (1 PRMT on the ALU pipe instead of 1 F2FP).  Truncation biases every P
// entry downward by < 2^-8 relative, but the row sums are taken from the SAME packed P (ones-column MMA), so the mean
// bias cancels in O = (P V) / (P 1); what is left is rounding noise of the same order as round-to-nearest.
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;\n" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
    return r;
}
// exp2 for x <= 0 on the FMA/ALU pipes only (no MUFU, no F2I): Cody-Waite split with the 1.5*2^23 magic constant,
// t = x + M holds n = rint(x) in its low mantissa bits, f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial
// (max relative error 7.5e-5, far below the bf16 rounding of P), exponent add by integer arithmetic.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.f);
    const float M = 12582912.f;
    const float t = x + M;
    const float n = t - M;
    const float f = x - n;
    const float p = fmaf(fmaf(fmaf(0.05517166f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.ftz.bf16x2 %0, %1;\n" : "=r"(y) : "r"(x));
    return y;
}

// swizzle of 16-byte chunks inside a row of CPR chunks so that 8 consecutive rows x one logical chunk hit
// 8 distinct 16-byte bank groups (conflict-free ldmatrix)
template <int CPR> __device__ __forceinline__ int swz(int row, int chunk) {
    if (CPR == 1) return chunk;
    constexpr int DIV = (8 / CPR) > 1 ? (8 / CPR) : 1;
    constexpr int MOD = CPR < 8 ? CPR : 8;
    return chunk ^ ((row / DIV) % MOD);
}

// EXPMODE: 0 = ex2.approx.f32 + cvt.rn pack, 1 = packed bf16x2 ex2, 2 = ex2.approx.f32 + truncating PRMT pack,
//          3 / 4 = mode 2 with 2 / 3 of every 8 score tiles evaluated by ex2_poly (MUFU : FMA-pipe split 75:25 / 62:38)
// SK = keys per pipeline stage (a multiple of the 64-key compute tile); 3-stage cp.async ring, ONE block barrier per
// stage: at iteration t the barrier both publishes stage t and proves every warp is done with stage t-1, whose slot
// is then refilled with stage t+2.
template <int HD, int MT, int NWARPS, int EXPMODE, int SK, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
attn_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int L, int C, float scale_log2,
                const int* __restrict__ redo) {
    // second pass of the bounded-softmax path: only the CTAs that kernel declined (flag != 0) do any work
    if (redo && redo[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] == 0) return;
    constexpr int CPR = HD / 8;            // 16-byte chunks per K/V row
    constexpr int NT = NWARPS * 32;
    constexpr int ROWS = NWARPS * MT * 16;  // query rows per CTA
    constexpr int NDT = HD / 8;            // 8-wide output column tiles
    constexpr int KSTEPS = HD >= 16 ? HD / 16 : 1;
    constexpr int TILE_ELEMS = KT * HD;
    constexpr int STAGE_ELEMS = SK * HD;
    constexpr int NSLOT = 3;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    bf16* Ks = reinterpret_cast<bf16*>(smem_raw);              // [3][SK][HD]
    bf16* Vs = Ks + NSLOT * STAGE_ELEMS;                       // [3][SK][HD]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z, h = blockIdx.y;
    const long long rstride = 3LL * C;
    const bf16* qbase = qkv + (long long)b * L * rstride + (long long)h * HD;
    const bf16* kbase = qbase + C;
    const bf16* vbase = qbase + 2 * C;

    // per-thread copy plan of one stage: SK*CPR 16-byte chunks of K and as many of V, CH chunk pairs per thread
    constexpr int CH = (SK * CPR + NT - 1) / NT;
    auto load_stage = [&](int t, int slot) {
        const bf16* ksrc = kbase + (long long)t * SK * rstride;
        const bf16* vsrc = vbase + (long long)t * SK * rstride;
        bf16* kd = Ks + slot * STAGE_ELEMS;
        bf16* vd = Vs + slot * STAGE_ELEMS;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int i = tid + c * NT;
            if (SK * CPR % NT == 0 || i < SK * CPR) {
                const int row = i / CPR, ch = i % CPR;
                const long long g = (long long)row * rstride + ch * 8;
                const int so = row * HD + swz<CPR>(row, ch) * 8;
                cp_async16(kd + so, ksrc + g);
                cp_async16(vd + so, vsrc + g);
            }
        }
    };

    // ---- Q fragments straight from global memory (A operand layout) ------------------------------------
    const int r_lo = lane >> 2, c_lo = (lane & 3) * 2;
    uint32_t qf[MT][KSTEPS][4];
    const int row0 = blockIdx.x * ROWS + warp * (MT * 16);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int ra = row0 + mt * 16 + r_lo, rb = ra + 8;
        const bf16* pa = qbase + (long long)(ra < L ? ra : L - 1) * rstride;
        const bf16* pb = qbase + (long long)(rb < L ? rb : L - 1) * rstride;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            qf[mt][ks][0] = *reinterpret_cast<const uint32_t*>(pa + ks * 16 + c_lo);
            qf[mt][ks][1] = *reinterpret_cast<const uint32_t*>(pb + ks * 16 + c_lo);
            if (HD >= 16) {
                qf[mt][ks][2] = *reinterpret_cast<const uint32_t*>(pa + ks * 16 + 8 + c_lo);
                qf[mt][ks][3] = *reinterpret_cast<const uint32_t*>(pb + ks * 16 + 8 + c_lo);
            } else {
                qf[mt][ks][2] = 0u; qf[mt][ks][3] = 0u;
            }
        }
    }

    float o[MT][NDT][4];
    float ls[MT][4];
    float mrow[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int dt = 0; dt < NDT; ++dt)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[mt][dt][i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) ls[mt][i] = 0.f;
        mrow[mt][0] = -INFINITY; mrow[mt][1] = -INFINITY;
    }
    const uint32_t ones = (lane >> 2) == 0 ? 0x3F803F80u : 0u;  // B fragment of a ones column at n = 0

    const int nstages = L / SK;
    load_stage(0, 0);
    cp_async_commit();
    if (nstages > 1) { load_stage(1, 1); cp_async_commit(); }

    const uint32_t ks_s = (uint32_t)__cvta_generic_to_shared(Ks);
    const uint32_t vs_s = (uint32_t)__cvta_generic_to_shared(Vs);
    const int lrow = lane & 7, lmat = lane >> 3;

    int slot = 0;
    for (int t = 0; t < nstages; ++t) {
        if (t + 1 < nstages) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        if (t + 2 < nstages) {
            load_stage(t + 2, slot >= 1 ? slot - 1 : NSLOT - 1);   // slot of stage t-1 == slot of stage t+2
            cp_async_commit();
        }
#pragma unroll 1
      for (int sub = 0; sub < SK / KT; ++sub) {
        const uint32_t kst = ks_s + (slot * STAGE_ELEMS + sub * TILE_ELEMS) * 2;
        const uint32_t vst = vs_s + (slot * STAGE_ELEMS + sub * TILE_ELEMS) * 2;

        // ---- S = Q K^T ----------------------------------------------------------------------------------
        float s[MT][8][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.f;

        if (HD >= 16) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
                for (int p = 0; p < 4; ++p) {  // n-tile pairs
                    const int row = (2 * p + (lmat >> 1)) * 8 + lrow;
                    const int ch = 2 * ks + (lmat & 1);
                    uint32_t kf[4];
                    ldsm_x4(kf, kst + (row * HD + swz<CPR>(row, ch) * 8) * 2);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_16816(s[mt][2 * p], qf[mt][ks], kf[0], kf[1]);
                        mma_16816(s[mt][2 * p + 1], qf[mt][ks], kf[2], kf[3]);
                    }
                }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) {  // four n-tiles per ldmatrix.x4 (rows are 16 B)
                const int row = (4 * p + lmat) * 8 + lrow;
                uint32_t kf[4];
                ldsm_x4(kf, kst + row * HD * 2);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) mma_1688(s[mt][4 * p + q], qf[mt][0][0], qf[mt][0][1], kf[q]);
            }
        }

        // ---- online softmax (exp2 domain), P packed to bf16 as the A operand of PV ------------------------
        uint32_t pf[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            float mx0 = s[mt][0][0], mx1 = s[mt][0][2];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                mx0 = fmaxf(mx0, fmaxf(s[mt][nt][0], s[mt][nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[mt][nt][2], s[mt][nt][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float mn0 = fmaxf(mrow[mt][0], mx0 * scale_log2);
            const float mn1 = fmaxf(mrow[mt][1], mx1 * scale_log2);
            // rescale the running output / row sums only when some row of this warp saw a new maximum (after the first
            // few key tiles that is rare: the expected number of record maxima over n tiles is ~ln n)
            if (__any_sync(0xffffffffu, mn0 != mrow[mt][0] || mn1 != mrow[mt][1])) {
                const float c0 = ex2f(mrow[mt][0] - mn0), c1 = ex2f(mrow[mt][1] - mn1);
                mrow[mt][0] = mn0; mrow[mt][1] = mn1;
#pragma unroll
                for (int dt = 0; dt < NDT; ++dt) {
                    o[mt][dt][0] *= c0; o[mt][dt][1] *= c0; o[mt][dt][2] *= c1; o[mt][dt][3] *= c1;
                }
                ls[mt][0] *= c0; ls[mt][1] *= c0; ls[mt][2] *= c1; ls[mt][3] *= c1;
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float x0 = fmaf(s[mt][nt][0], scale_log2, -mn0), x1 = fmaf(s[mt][nt][1], scale_log2, -mn0);
                const float x2 = fmaf(s[mt][nt][2], scale_log2, -mn1), x3 = fmaf(s[mt][nt][3], scale_log2, -mn1);
                uint32_t p01, p23;
                if (EXPMODE == 1) {
                    p01 = ex2_bf16x2(pack_bf16(x0, x1));
                    p23 = ex2_bf16x2(pack_bf16(x2, x3));
                } else if (EXPMODE == 2) {
                    p01 = pack_bf16_trunc(ex2f(x0), ex2f(x1));
                    p23 = pack_bf16_trunc(ex2f(x2), ex2f(x3));
                } else if (EXPMODE == 3 || EXPMODE == 4) {
                    const bool poly = (EXPMODE == 3) ? ((nt & 3) == 3) : (nt == 2 || nt == 5 || nt == 7);
                    if (poly) {
                        p01 = pack_bf16_trunc(ex2_poly(x0), ex2_poly(x1));
                        p23 = pack_bf16_trunc(ex2_poly(x2), ex2_poly(x3));
                    } else {
                        p01 = pack_bf16_trunc(ex2f(x0), ex2f(x1));
                        p23 = pack_bf16_trunc(ex2f(x2), ex2f(x3));
                    }
                } else {
                    p01 = pack_bf16(ex2f(x0), ex2f(x1));
                    p23 = pack_bf16(ex2f(x2), ex2f(x3));
                }
                pf[mt][nt >> 1][(nt & 1) * 2 + 0] = p01;
                pf[mt][nt >> 1][(nt & 1) * 2 + 1] = p23;
            }
        }

        // ---- O += P V ; row sums += P 1 ---------------------------------------------------------------------
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) mma_16816(ls[mt], pf[mt][kk], ones, ones);
            if (HD >= 16) {
#pragma unroll
                for (int a = 0; a < NDT / 2; ++a) {
                    const int row = kk * 16 + (lmat & 1) * 8 + lrow;
                    const int ch = 2 * a + (lmat >> 1);
                    uint32_t vf[4];
                    ldsm_x4_t(vf, vst + (row * HD + swz<CPR>(row, ch) * 8) * 2);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_16816(o[mt][2 * a], pf[mt][kk], vf[0], vf[1]);
                        mma_16816(o[mt][2 * a + 1], pf[mt][kk], vf[2], vf[3]);
                    }
                }
            } else {
                const int row = kk * 16 + (lmat & 1) * 8 + lrow;
                uint32_t vf[2];
                ldsm_x2_t(vf, vst + row * HD * 2);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) mma_16816(o[mt][0], pf[mt][kk], vf[0], vf[1]);
            }
        }
      }
        slot = slot + 1 == NSLOT ? 0 : slot + 1;
    }

    // ---- normalise and store --------------------------------------------------------------------------------
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const float l0 = __shfl_sync(0xffffffffu, ls[mt][0], lane & ~3);
        const float l1 = __shfl_sync(0xffffffffu, ls[mt][2], lane & ~3);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        const int ra = row0 + mt * 16 + r_lo, rb = ra + 8;
        if (lse && (lane & 3) == 0) {   // natural-log log-sum-exp of the scaled scores (training backward)
            float* lp = lse + ((long long)b * gridDim.y + h) * L;
            if (ra < L) lp[ra] = (mrow[mt][0] + log2f(l0)) * 0.69314718055994531f;
            if (rb < L) lp[rb] = (mrow[mt][1] + log2f(l1)) * 0.69314718055994531f;
        }
        bf16* oa = out + ((long long)b * L + ra) * C + (long long)h * HD + c_lo;
        bf16* ob = out + ((long long)b * L + rb) * C + (long long)h * HD + c_lo;
#pragma unroll
        for (int dt = 0; dt < NDT; ++dt) {
            if (ra < L) *reinterpret_cast<uint32_t*>(oa + dt * 8) = pack_bf16(o[mt][dt][0] * i0, o[mt][dt][1] * i0);
            if (rb < L) *reinterpret_cast<uint32_t*>(ob + dt * 8) = pack_bf16(o[mt][dt][2] * i1, o[mt][dt][3] * i1);
        }
    }
}

template <int HD, int MT, int NWARPS, int EXPMODE, int SK, int MINB>
int launch_sk(const void* qkv, void* out, float* lse, int B, int L, int C, int heads, float scale_log2, const int* redo, cudaStream_t st) {
    constexpr int ROWS = NWARPS * MT * 16;
    const size_t smem = (size_t)3 * 2 * SK * HD * sizeof(bf16);
    auto kern = attn_mma_kernel<HD, MT, NWARPS, EXPMODE, SK, MINB>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ddpmir_set_error("attention: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
    }
    dim3 grid(ceil_div(L, ROWS), heads, B);
    kern<<<grid, NWARPS * 32, smem, st>>>((const bf16*)qkv, (bf16*)out, lse, L, C, scale_log2, redo);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// stage size: as many keys as fit the budget and divide L (fewer block barriers per key), 64 otherwise
template <int HD, int MT, int NWARPS, int EXPMODE>
int launch(const void* qkv, void* out, float* lse, int B, int L, int C, int heads, float scale_log2, const int* redo, cudaStream_t st) {
    constexpr int SKMAX = HD <= 16 ? 256 : (HD == 32 ? 128 : 64);
    constexpr int MINB = (HD <= 16 && MT == 1) ? 4 : 1;
    if (SKMAX > 64 && L % SKMAX == 0) return launch_sk<HD, MT, NWARPS, EXPMODE, SKMAX, MINB>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st);
    return launch_sk<HD, MT, NWARPS, EXPMODE, 64, MINB>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// Bounded-softmax attention (inference, head_dim 8 / 16).
//
// The kernel above spends, per score, one FFMA (scale, subtract the running maximum), one FMNMX (maximum) and its share
// of shuffles / votes / rescales next to the one exp2 that is the real work -- and the exp2 pipe (16 /clk/SM) is the
// bound.  softmax only needs SOME per-row offset m_i that keeps exp2(s - m_i) inside the fp32 range, not the maximum:
//   * q arrives already multiplied by log2(e)/sqrt(head_dim) (folded into in_proj when the weights are packed), so the
//     tensor core produces s' = q'.k in the exp2 domain -- no scaling instruction;
//   * |s'_ij| <= B_i = |q'_i| * max_j |k_j| (Cauchy-Schwarz; max_j |k_j| per (image, head) from a small pre-pass).  While
//     B_i <= 60 the offset can simply be m_i = 0: every P = exp2(s') lies in [2^-60, 2^60], row sums stay below 2^76,
//     nothing overflows and no row underflows to zero.  P, O and the row sums are floating point, so the common factor
//     2^(max_j s'_ij) cancels in O / l exactly as it does with the true maximum.  No running maximum, no shuffle, no
//     rescale of O, and the FMA-pipe exp2 needs no range clamp;
//   * CTAs with a row whose B_i > 60 (logit bound beyond 41 nats) set a flag and leave; a second launch of the exact
//     online-maximum kernel redoes only those CTAs.
// What remains per score: the exp2 (MUFU, or the FMA-pipe polynomial for POLY of every 8 score tiles), half a pack, and
// the MMA / ldmatrix share.
__device__ __forceinline__ void mma_1688_c(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, float c0, float c1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%8,%8};\n"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mma_16816_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%11,%11};\n"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ float sumsq_bf16x2(uint32_t v) {
    const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
    return lo * lo + hi * hi;
}

constexpr float BOUND_LIMIT = 60.f;    // rows with a larger logit bound (exp2 domain) go to the exact kernel

// ex2_poly without the range clamp (|x| <= BOUND_LIMIT here) and with a degree-2 polynomial option
template <int DEG> __device__ __forceinline__ float ex2_poly_bounded(float x) {
    const float M = 12582912.f;
    const float t = x + M;
    const float n = t - M;
    const float f = x - n;
    const float p = DEG == 3 ? fmaf(fmaf(fmaf(0.05517166f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f)
                             : fmaf(fmaf(0.23842894f, f, 0.70344800f), f, 1.00044310f);   // minimax, |rel err| < 1.73e-3 on [-0.5, 0.5]
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// kmax[b, h] = max_j |k_j|  (bf16 values as the MMA sees them)
__device__ __forceinline__ float sumsq_f16x2(uint32_t v) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    return f.x * f.x + f.y * f.y;
}
template <int HD, bool F16 = false>
__global__ void __launch_bounds__(256)
attn_kbound_kernel(const bf16* __restrict__ qkv, float* __restrict__ kmax, int L, int C, int* __restrict__ zero_me = nullptr) {
    if (zero_me && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *zero_me = 0;   // the f16 tier's decline counter
    const int b = blockIdx.y, h = blockIdx.x;
    const long long rstride = 3LL * C;
    const bf16* kbase = qkv + (long long)b * L * rstride + C + (long long)h * HD;
    float best = 0.f;
#pragma unroll 4
    for (int j = threadIdx.x; j < L; j += 256) {
        const uint4* kp = reinterpret_cast<const uint4*>(kbase + (long long)j * rstride);
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
            const uint4 v = __ldg(kp + c);
            ss += F16 ? sumsq_f16x2(v.x) + sumsq_f16x2(v.y) + sumsq_f16x2(v.z) + sumsq_f16x2(v.w)
                      : sumsq_bf16x2(v.x) + sumsq_bf16x2(v.y) + sumsq_bf16x2(v.z) + sumsq_bf16x2(v.w);
        }
        best = fmaxf(best, ss);
    }
    __shared__ float red[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) best = fmaxf(best, red[w]);
        kmax[b * gridDim.x + h] = sqrtf(best);
    }
}

// POLY = how many of every 8 score tiles take the FMA-pipe exp2 (polynomial degree DEG); P is packed by truncation (PRMT)
template <int HD, int POLY, int DEG, int SK>
__global__ void __launch_bounds__(256, 4)
attn_bounded_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, const float* __restrict__ kmax, int* __restrict__ flags,
                    int L, int C) {
    constexpr int CPR = HD / 8;
    constexpr int NT = 256;
    constexpr int ROWS = 128;
    constexpr int NDT = HD / 8;
    constexpr int STAGE_ELEMS = SK * HD;
    constexpr int NSLOT = 3;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
    bf16* Vs = Ks + NSLOT * STAGE_ELEMS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z, h = blockIdx.y;
    const long long rstride = 3LL * C;
    const bf16* qbase = qkv + (long long)b * L * rstride + (long long)h * HD;
    const bf16* kbase = qbase + C;
    const bf16* vbase = qbase + 2 * C;

    // ---- Q fragments, per-row logit bound, fixed offsets -------------------------------------------------------------
    const int r_lo = lane >> 2, c_lo = (lane & 3) * 2;
    const int row0 = blockIdx.x * ROWS + warp * 16;
    const int ra = row0 + r_lo, rb = ra + 8;
    uint32_t qf[4];
    {
        const bf16* pa = qbase + (long long)(ra < L ? ra : L - 1) * rstride;
        const bf16* pb = qbase + (long long)(rb < L ? rb : L - 1) * rstride;
        qf[0] = *reinterpret_cast<const uint32_t*>(pa + c_lo);
        qf[1] = *reinterpret_cast<const uint32_t*>(pb + c_lo);
        qf[2] = HD >= 16 ? *reinterpret_cast<const uint32_t*>(pa + 8 + c_lo) : 0u;
        qf[3] = HD >= 16 ? *reinterpret_cast<const uint32_t*>(pb + 8 + c_lo) : 0u;
    }
    float qa = sumsq_bf16x2(qf[0]) + sumsq_bf16x2(qf[2]), qb = sumsq_bf16x2(qf[1]) + sumsq_bf16x2(qf[3]);
    qa += __shfl_xor_sync(0xffffffffu, qa, 1); qa += __shfl_xor_sync(0xffffffffu, qa, 2);
    qb += __shfl_xor_sync(0xffffffffu, qb, 1); qb += __shfl_xor_sync(0xffffffffu, qb, 2);
    const float kb = kmax[b * gridDim.y + h] * 1.0001f;
    const float ba = sqrtf(qa) * kb, bb = sqrtf(qb) * kb;
    const int bad = __syncthreads_or(!(ba <= BOUND_LIMIT) || !(bb <= BOUND_LIMIT));   // NaN counts as bad
    if (tid == 0) flags[((long long)b * gridDim.y + h) * gridDim.x + blockIdx.x] = bad;
    if (bad) return;

    constexpr int CH = (SK * CPR + NT - 1) / NT;
    auto load_stage = [&](int t, int slot) {
        const bf16* ksrc = kbase + (long long)t * SK * rstride;
        const bf16* vsrc = vbase + (long long)t * SK * rstride;
        bf16* kd = Ks + slot * STAGE_ELEMS;
        bf16* vd = Vs + slot * STAGE_ELEMS;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int i = tid + c * NT;
            if (SK * CPR % NT == 0 || i < SK * CPR) {
                const int row = i / CPR, ch = i % CPR;
                const long long g = (long long)row * rstride + ch * 8;
                const int so = row * HD + swz<CPR>(row, ch) * 8;
                cp_async16(kd + so, ksrc + g);
                cp_async16(vd + so, vsrc + g);
            }
        }
    };

    float o[NDT][4];
    float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dt = 0; dt < NDT; ++dt)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[dt][i] = 0.f;
    const uint32_t ones = (lane >> 2) == 0 ? 0x3F803F80u : 0u;

    const int nstages = L / SK;
    load_stage(0, 0);
    cp_async_commit();
    if (nstages > 1) { load_stage(1, 1); cp_async_commit(); }

    const uint32_t ks_s = (uint32_t)__cvta_generic_to_shared(Ks);
    const uint32_t vs_s = (uint32_t)__cvta_generic_to_shared(Vs);
    const int lrow = lane & 7, lmat = lane >> 3;

    int slot = 0;
    for (int t = 0; t < nstages; ++t) {
        if (t + 1 < nstages) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        if (t + 2 < nstages) {
            load_stage(t + 2, slot >= 1 ? slot - 1 : NSLOT - 1);
            cp_async_commit();
        }
        // The stage is consumed in 32-key half tiles, software-pipelined inside the warp: the QK^T MMAs of half tile
        // i+1 are issued BEFORE the exps of half tile i, so the tensor pipe works under the MUFU/FMA phase of the same
        // warp instead of only under other warps' (which the per-stage barrier keeps in the same phase).
        const uint32_t kst = ks_s + slot * STAGE_ELEMS * 2;
        const uint32_t vst = vs_s + slot * STAGE_ELEMS * 2;
        constexpr int NH = SK / 32;
        float s[2][4][4];
        auto qk = [&](int hh, float (&d)[4][4]) {
            if (HD >= 16) {
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int row = hh * 32 + (2 * p + (lmat >> 1)) * 8 + lrow;
                    const int ch = lmat & 1;
                    uint32_t kf[4];
                    ldsm_x4(kf, kst + (row * HD + swz<CPR>(row, ch) * 8) * 2);
                    mma_16816_c(d[2 * p], qf, kf[0], kf[1], 0.f, 0.f);
                    mma_16816_c(d[2 * p + 1], qf, kf[2], kf[3], 0.f, 0.f);
                }
            } else {
                const int row = hh * 32 + lmat * 8 + lrow;
                uint32_t kf[4];
                ldsm_x4(kf, kst + row * HD * 2);
#pragma unroll
                for (int q = 0; q < 4; ++q) mma_1688_c(d[q], qf[0], qf[1], kf[q], 0.f, 0.f);
            }
        };
        qk(0, s[0]);
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            float (&sc)[4][4] = s[hh & 1];
            if (hh + 1 < NH) qk(hh + 1, s[(hh + 1) & 1]);

            // ---- P = exp2(.) packed to bf16 (A operand of PV) -------------------------------------------------------
            uint32_t pf[2][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n8 = (hh & 1) * 4 + nt;   // position inside the 8-tile pattern
                const bool poly = POLY == 1 ? n8 == 5 : POLY == 2 ? (n8 & 3) == 3 : POLY == 3 ? (n8 == 2 || n8 == 5 || n8 == 7)
                                : POLY == 4 ? (n8 & 1) == 1 : POLY == 5 ? (n8 != 0 && n8 != 3 && n8 != 6) : false;
                float e0, e1, e2, e3;
                if (poly) { e0 = ex2_poly_bounded<DEG>(sc[nt][0]); e1 = ex2_poly_bounded<DEG>(sc[nt][1]); e2 = ex2_poly_bounded<DEG>(sc[nt][2]); e3 = ex2_poly_bounded<DEG>(sc[nt][3]); }
                else { e0 = ex2f(sc[nt][0]); e1 = ex2f(sc[nt][1]); e2 = ex2f(sc[nt][2]); e3 = ex2f(sc[nt][3]); }
                pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16_trunc(e0, e1);
                pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16_trunc(e2, e3);
            }

            // ---- O += P V ; row sums += P 1 ------------------------------------------------------------------------
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                mma_16816(ls, pf[kk], ones, ones);
                const int row = hh * 32 + kk * 16 + (lmat & 1) * 8 + lrow;
                if (HD >= 16) {
                    const int ch = lmat >> 1;
                    uint32_t vf[4];
                    ldsm_x4_t(vf, vst + (row * HD + swz<CPR>(row, ch) * 8) * 2);
                    mma_16816(o[0], pf[kk], vf[0], vf[1]);
                    mma_16816(o[NDT - 1], pf[kk], vf[2], vf[3]);
                } else {
                    uint32_t vf[2];
                    ldsm_x2_t(vf, vst + row * HD * 2);
                    mma_16816(o[0], pf[kk], vf[0], vf[1]);
                }
            }
        }
        slot = slot + 1 == NSLOT ? 0 : slot + 1;
    }

    const float l0 = __shfl_sync(0xffffffffu, ls[0], lane & ~3);
    const float l1 = __shfl_sync(0xffffffffu, ls[2], lane & ~3);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    bf16* oa = out + ((long long)b * L + ra) * C + (long long)h * HD + c_lo;
    bf16* ob = out + ((long long)b * L + rb) * C + (long long)h * HD + c_lo;
#pragma unroll
    for (int dt = 0; dt < NDT; ++dt) {
        if (ra < L) *reinterpret_cast<uint32_t*>(oa + dt * 8) = pack_bf16(o[dt][0] * i0, o[dt][1] * i0);
        if (rb < L) *reinterpret_cast<uint32_t*>(ob + dt * 8) = pack_bf16(o[dt][2] * i1, o[dt][3] * i1);
    }
}

template <int HD, int POLY, int DEG, int SK>
int launch_bounded_sk(const void* qkv, void* out, const float* kmax, int* flags, int B, int L, int C, int heads, cudaStream_t st) {
    const size_t smem = (size_t)3 * 2 * SK * HD * sizeof(bf16);
    auto kern = attn_bounded_kernel<HD, POLY, DEG, SK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { ddpmir_set_error("attention: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
    }
    dim3 grid(ceil_div(L, 128), heads, B);
    kern<<<grid, 256, smem, st>>>((const bf16*)qkv, (bf16*)out, kmax, flags, L, C);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
template <int HD, int POLY, int DEG>
int launch_bounded(const void* qkv, void* out, const float* kmax, int* flags, int B, int L, int C, int heads, cudaStream_t st) {
    if (L % 256 == 0) return launch_bounded_sk<HD, POLY, DEG, 256>(qkv, out, kmax, flags, B, L, C, heads, st);
    return launch_bounded_sk<HD, POLY, DEG, 64>(qkv, out, kmax, flags, B, L, C, heads, st);
}

int g_expmode = -1;   // -1 = auto: mode 3 (25 % of the exps on the FMA pipe) at head_dim 8, mode 0 otherwise (measured on B200)
int g_mt = 1;
int g_poly = -1;      // bounded kernel: score tiles of 8 on the FMA pipe (-1 = auto), +8 = degree-2 instead of degree-3 polynomial
int g_split16 = 0;    // half-precision tier (attn_tc16.cu): MUFU share of its two FMA-pipe variants, 0 = default
int g_lin_max_set = 6; // polynomial-kernel tier (attn_lin.cu): largest polynomial set it may use, -1 = tier off
int g_lin_simt = 0;    // ... 1 = its fp32 SIMT kernels instead of the tcgen05 ones

}  // namespace

// test / tuning hook: expmode 0 = fp32 ex2 + cvt pack, 1 = packed bf16x2 ex2, 2 = fp32 ex2 + truncating pack;
// +16 selects two 16-row tiles per warp (head_dim 8/16 only); bits 8..: (value + 1) of the bounded kernel's POLY/pack choice
extern "C" int ddpmir_attention_set_expmode(int mode) {
    if (mode < 0) { g_expmode = -1; g_mt = 1; g_poly = -1; g_split16 = 0; return DDPMIR_OK; }
    g_split16 = (mode >> 16) & 0xfff;             // bits 16..23: attn_tc16's split (see ddpmir_attention_tc16)
    g_poly = ((mode >> 8) & 255) - 1;           // bits 8..15: (value + 1) of the bounded kernels' choice; +32 skips the half-precision tier
    mode &= 255;
    g_expmode = mode & 15;
    if (g_expmode > 4) g_expmode = 0;
    g_mt = (mode & 16) ? 2 : 1;
    return DDPMIR_OK;
}

int ddpmir_attention_simt(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, float qscale, cudaStream_t st);
int ddpmir_attention_tc(const void* qkv, void* out, const float* kmax, int* flags, int B, int L, int C, int heads, int sel, int redo, cudaStream_t st);
int ddpmir_attention_tc16(const void* qkv, void* out, const float* kmax, int* flags, int* declined, const int* skip, int B, int L, int C,
                          int heads, int split, cudaStream_t st);
size_t ddpmir_attention_lin_workspace(int B, int L, int C, int heads);
int ddpmir_attention_lin(const void* qkv, void* out, float* kmax, int* flags, int* declined, void* lin_ws, const int** tier_out,
                         int B, int L, int C, int heads, int max_set, int simt, cudaStream_t st);

// test / tuning hook: largest polynomial set of the polynomial-kernel tier (0..6, see attn_lin.cuh), -1 switches the tier off;
// +16 takes the tier's fp32 SIMT kernels instead of the tcgen05 ones
extern "C" int ddpmir_attention_set_lin(int max_set) {
    if (max_set < 0) { g_lin_max_set = -1; g_lin_simt = 0; return DDPMIR_OK; }
    g_lin_simt = (max_set >> 4) & 1;
    g_lin_max_set = (max_set & 15) > 6 ? 6 : (max_set & 15);
    return DDPMIR_OK;
}

static int attention_mma_scaled(const void* qkv, int B, int L, int C, int heads, void* out, float* lse, float scale_log2, const int* redo,
                                bool force_mt1, cudaStream_t st) {
    const int hd = C / heads;
    if (L % KT != 0) return DDPMIR_ERR_UNSUPPORTED;
    const int em = g_expmode >= 0 ? g_expmode : (hd == 8 ? 3 : 0);
    const int mt = force_mt1 ? 1 : g_mt;
#define GO(HD, MT, NW) (em == 1 ? launch<HD, MT, NW, 1>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st) : \
                        em == 2 ? launch<HD, MT, NW, 2>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st) : \
                        em == 3 ? launch<HD, MT, NW, 3>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st) : \
                        em == 4 ? launch<HD, MT, NW, 4>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st) : \
                                  launch<HD, MT, NW, 0>(qkv, out, lse, B, L, C, heads, scale_log2, redo, st))
    switch (hd) {
        case 8: return mt == 2 ? GO(8, 2, 8) : GO(8, 1, 8);
        case 16: return mt == 2 ? GO(16, 2, 8) : GO(16, 1, 8);
        case 32: return GO(32, 1, 8);
        case 64: return GO(64, 1, 4);
        case 128: return GO(128, 1, 4);
        default: return DDPMIR_ERR_UNSUPPORTED;
    }
#undef GO
}

int ddpmir_attention_mma(const void* qkv, int B, int L, int C, int heads, void* out, float* lse, cudaStream_t st) {
    return attention_mma_scaled(qkv, B, L, C, heads, out, lse, 1.4426950408889634f / sqrtf((float)(C / heads)), nullptr, false, st);
}

extern "C" int ddpmir_attention(const void* qkv, int dtype, int B, int L, int C, int heads, void* out, int impl,
                                ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(qkv && out, "attention: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && L > 0 && heads > 0 && C % heads == 0 && (C / heads) % 8 == 0, "attention: bad shape");
    DDPMIR_CHECK_ARG(B <= 65535 && heads <= 65535, "attention: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DDPMIR_BF16 && impl != DDPMIR_IMPL_SIMT) {
        int rc = ddpmir_attention_mma(qkv, B, L, C, heads, out, nullptr, st);
        if (rc != DDPMIR_ERR_UNSUPPORTED) return rc;
        if (impl == DDPMIR_IMPL_TENSOR) {
            ddpmir_set_error("attention: tensor-core kernel needs L %% 64 == 0 and head_dim in {8,16,32,64,128}");
            return rc;
        }
    } else if (impl == DDPMIR_IMPL_TENSOR) {
        ddpmir_set_error("attention: tensor-core kernel is bf16 only");
        return DDPMIR_ERR_UNSUPPORTED;
    }
    return ddpmir_attention_simt(qkv, dtype, B, L, C, heads, out, 1.f / sqrtf((float)(C / heads)), st);
}

extern "C" size_t ddpmir_attention_prescaled_workspace(int B, int L, int heads) {
    return ((size_t)B * heads + (size_t)B * heads * ceil_div(L, 128)) * 4;
}

extern "C" int ddpmir_attention_prescaled(const void* qkv, int B, int L, int C, int heads, void* workspace, void* out,
                                          ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(qkv && out && workspace, "attention_prescaled: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && L > 0 && heads > 0 && C % heads == 0 && (C / heads) % 8 == 0, "attention_prescaled: bad shape");
    DDPMIR_CHECK_ARG(B <= 65535 && heads <= 65535, "attention_prescaled: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int hd = C / heads;
    if ((hd == 8 || hd == 16) && L % KT == 0) {
        float* kmax = (float*)workspace;
        int* flags = (int*)(kmax + (size_t)B * heads);
        if (hd == 8) attn_kbound_kernel<8><<<dim3(heads, B), 256, 0, st>>>((const bf16*)qkv, kmax, L, C);
        else attn_kbound_kernel<16><<<dim3(heads, B), 256, 0, st>>>((const bf16*)qkv, kmax, L, C);
        DDPMIR_LAUNCH_CHECK();
        // exp split (measured on B200, see profiles/): tcgen05 kernel -> every other score pair on the FMA/ALU pipes as packed
        // bf16 (20), mma.sync kernel -> 2 of 8 with the degree-3 fp32 polynomial (2); the UNet's bf16 error moves by < 4 %
        const bool tc_ok = !(g_poly >= 0 && (g_poly & 64)) && L >= 1024 && L % 128 == 0;
        const int sel = g_poly >= 0 ? g_poly : (tc_ok ? 20 : 2);
        const int poly = sel & 7;
        const bool rn = (sel & 8) != 0;   // +8: degree-2 polynomial
        int rc = DDPMIR_ERR_UNSUPPORTED;
        // long sequences: tcgen05 / TMEM kernel (attn_tc.cu); +64 in the tuning hook keeps the mma.sync kernel
        if (tc_ok) rc = ddpmir_attention_tc(qkv, out, kmax, flags, B, L, C, heads, sel & 31, 0, st);
        if (rc == DDPMIR_ERR_UNSUPPORTED) {
#define GB(HD) (rn ? (poly == 0 ? launch_bounded<HD, 0, 2>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 1 ? launch_bounded<HD, 1, 2>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 2 ? launch_bounded<HD, 2, 2>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 3 ? launch_bounded<HD, 3, 2>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 4 ? launch_bounded<HD, 4, 2>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                                  launch_bounded<HD, 5, 2>(qkv, out, kmax, flags, B, L, C, heads, st)) \
                   : (poly == 0 ? launch_bounded<HD, 0, 3>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 1 ? launch_bounded<HD, 1, 3>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 2 ? launch_bounded<HD, 2, 3>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 3 ? launch_bounded<HD, 3, 3>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                      poly == 4 ? launch_bounded<HD, 4, 3>(qkv, out, kmax, flags, B, L, C, heads, st) : \
                                  launch_bounded<HD, 5, 3>(qkv, out, kmax, flags, B, L, C, heads, st)))
        rc = hd == 8 ? GB(8) : GB(16);
#undef GB
        }
        if (rc != DDPMIR_OK) return rc;
        // exact kernel for the CTAs the bounded kernel declined (same 128-row partition)
        return attention_mma_scaled(qkv, B, L, C, heads, out, nullptr, 1.f, flags, true, st);
    }
    int rc = attention_mma_scaled(qkv, B, L, C, heads, out, nullptr, 1.f, nullptr, false, st);
    if (rc != DDPMIR_ERR_UNSUPPORTED) return rc;
    return ddpmir_attention_simt(qkv, DDPMIR_BF16, B, L, C, heads, out, 0.69314718055994531f, st);
}

// ---- binary16 qkv: the three-tier path of the full-resolution blocks ----------------------------------------------------
namespace {
// dst (bf16) = src (f16), only when the half-precision tier declined at least one CTA
__global__ void __launch_bounds__(256)
f16_to_bf16_if_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n8, const int* __restrict__ declined) {
    if (*declined == 0) return;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
        const uint4 v = __ldg(src + i);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
        uint4 o;
        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) ob[k] = __float22bfloat162_rn(__half22float2(h[k]));
        dst[i] = o;
    }
}
inline size_t f16_header_bytes(int B, int L, int heads) {
    const size_t n = (size_t)B * heads + (size_t)B * heads * ceil_div(L, 128) + 4;
    return (n * 4 + 255) / 256 * 256;
}
}  // namespace

extern "C" size_t ddpmir_attention_prescaled_f16_workspace(int B, int L, int C, int heads) {
    return f16_header_bytes(B, L, heads) + ddpmir_attention_lin_workspace(B, L, C, heads) + (size_t)B * L * 3 * C * 2;
}

extern "C" int ddpmir_attention_prescaled_f16(const void* qkv, int B, int L, int C, int heads, void* workspace, void* out,
                                              ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(qkv && out && workspace, "attention_prescaled_f16: null pointer");
    DDPMIR_CHECK_ARG(B > 0 && L > 0 && heads > 0 && C % heads == 0, "attention_prescaled_f16: bad shape");
    DDPMIR_CHECK_ARG(B <= 65535 && heads <= 65535, "attention_prescaled_f16: grid too large");
    const int hd = C / heads;
    if ((hd != 8 && hd != 16) || L < 1024 || L % 128 != 0) {
        ddpmir_set_error("attention_prescaled_f16: head_dim 8/16 and L >= 1024, L %% 128 == 0 only (got hd %d, L %d)", hd, L);
        return DDPMIR_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* kmax = (float*)workspace;
    int* flags = (int*)(kmax + (size_t)B * heads);
    int* declined = flags + (size_t)B * heads * ceil_div(L, 128);
    void* lin_ws = (char*)workspace + f16_header_bytes(B, L, heads);
    void* copy = (char*)lin_ws + ddpmir_attention_lin_workspace(B, L, C, heads);
    // pre-pass (max |q'|, max |k| per image and head) + tier 0: (image, head) pairs whose logit bound is <= 2 are computed through
    // the polynomial feature map in O(L) (attn_lin.cu); tier[] >= 0 marks them and the kernels below skip their CTAs
    const int* tier = nullptr;
    int rc = ddpmir_attention_lin(qkv, out, kmax, flags, declined, lin_ws, &tier, B, L, C, heads, g_lin_max_set, g_lin_simt, st);
    if (rc != DDPMIR_OK) return rc;
    // tier 1: S and P in binary16, logit bound <= 11
    rc = ddpmir_attention_tc16(qkv, out, kmax, flags, declined, tier, B, L, C, heads, g_split16, st);
    if (rc != DDPMIR_OK) return rc;
    // tiers 2 and 3 read bf16: a copy made only if some CTA was declined (the kernels below return at once otherwise)
    const long long n8 = (long long)B * L * 3 * C / 8;
    f16_to_bf16_if_kernel<<<148 * 8, 256, 0, st>>>((const uint4*)qkv, (uint4*)copy, n8, declined);
    DDPMIR_LAUNCH_CHECK();
    const int sel = g_poly >= 0 ? (g_poly & 31) : 20;
    rc = ddpmir_attention_tc(copy, out, kmax, flags, B, L, C, heads, sel, 1, st);           // tier 2: P in bf16, bound <= 60
    if (rc != DDPMIR_OK) return rc;
    return attention_mma_scaled(copy, B, L, C, heads, out, nullptr, 1.f, flags, true, st);  // tier 3: exact online-max kernel
}
