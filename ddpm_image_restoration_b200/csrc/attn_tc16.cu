// Half-precision tier of the bounded-softmax self-attention on tcgen05 / TMEM / TMA (inference, head_dim 8 / 16, L % 128 == 0).
//
// attn_tc.cu keeps P in bf16 so that logits up to +-60 (exp2 domain) never overflow; every score then costs 1.5 (MUFU path) to
// 3.5 (packed bf16 path) issue slots plus its share of a 64-register TMEM load.  The UNets this kernel serves produce
// logits of a few units (a Cauchy-Schwarz bound of 1.5 for a fresh network), and for |s'| <= 11 the whole softmax fits
// binary16: S = Q' K^T is accumulated by the tensor core and stored as f16 (kind::f16, D = f16), tcgen05.ld.pack::16b hands a
// thread TWO scores per register, and P = exp2(s') is evaluated on register pairs:
//   * MUFU:  ex2.approx.f16x2 -- one instruction per PAIR, result already packed for the P store;
//   * FMA pipe, CTA logit bound <= 2:  degree-4 minimax polynomial of 2^x on [-2, 2], four packed HFMA2 per pair
//     (0.25 % max relative error in exact arithmetic, 0.17 % rms with the f16 roundings);
//   * FMA pipe, bound <= 11:  t = s + 1036 rounds s to an integer n in f16 (ulp 1 in [1024, 2048)), f = s - n, a degree-2
//     minimax 2^f with coefficients pre-scaled by 2^-12, and (t & 31) << 10 added to the exponent field: seven instructions.
// P (f16, 11-bit significand -- three bits more than the bf16 tier keeps) is the A operand of O += P V straight from TMEM.
// The tensor core takes ONE operand format per instruction and accumulates into f16 only from f16 operands (measured on
// B200: D = f16 with bf16 operands, or A = f16 with B = bf16, raise "illegal instruction"), so this tier reads an f16 qkv:
// the in_proj GEMM writes binary16 for the layers that come here (epilogue out_dtype DDPMIR_F16).  V carries the ones
// column that accumulates the row sums, O is fp32.  CTAs whose bound exceeds 11 write flag 1, bump the decline counter and
// leave: the entry point converts qkv to bf16 (a kernel that returns at once while the counter is 0) and attn_tc.cu's bf16
// kernel redoes those CTAs (and hands rows beyond its own window to the exact kernel).  Structure (TMA ring, one MMA issuer
// warp per softmax warpgroup, two CTAs per SM, thread = query row) is attn_tc.cu's; see there for the reasoning.
#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TQ = 128;                         // query rows per CTA
constexpr int TK = 64;                          // keys per tile
// CTA geometry (template parameters NWG, CTAS): NWG softmax warpgroups share the CTA's 128 query rows and split the keys;
// every warpgroup has its own S and P buffer (96 TMEM columns) and its own MMA-issuer warp.  Two CTAs of two warpgroups per SM
// (256 TMEM columns each) overlap one CTA's prologue / epilogue with the other's main loop; one CTA of five warpgroups
// (5 x 96 + 32 = 512 columns) puts 25 % more softmax warps on the SM but measured no faster (see launch16).
constexpr int nstage(int nwg) { return nwg > 2 ? 16 : 8; }   // K/V ring: a power of two (the issuers' stage arithmetic is masks), >= 3 tiles ahead of the warpgroups
constexpr int BLK = TK * 16;                    // TK rows x 16 bytes: TK/8 core matrices of 8 rows
constexpr int QBLK = TQ * 16;
constexpr float BOUND_DIRECT = 2.0f;            // CTA logit bound up to which the unreduced polynomial is used
constexpr float BOUND_F16 = 11.0f;              // ... and up to which P fits binary16 with the offset 0

__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;\n" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t fma_h2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;\n" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// 2^x for a pair, |x| <= 2: 0.99772 + x (0.68905 + x (0.24515 + x (0.061450 + 0.0088752 x)))
__device__ __forceinline__ uint32_t ex2_direct_h2(uint32_t x) {
    uint32_t u = fma_h2(0x208B208Bu, x, 0x2BDE2BDEu);
    u = fma_h2(u, x, 0x33D833D8u);
    u = fma_h2(u, x, 0x39833983u);
    return fma_h2(u, x, 0x3BFB3BFBu);
}
// 2^x for a pair, |x| <= 11 (see the header)
__device__ __forceinline__ uint32_t ex2_reduced_h2(uint32_t x) {
    uint32_t t, n, f, r;
    asm("add.rn.f16x2 %0, %1, %2;\n" : "=r"(t) : "r"(x), "r"(0x640C640Cu));          // + 1036
    asm("add.rn.f16x2 %0, %1, %2;\n" : "=r"(n) : "r"(t), "r"(0xE40CE40Cu));          // - 1036: the integer part
    asm("sub.rn.f16x2 %0, %1, %2;\n" : "=r"(f) : "r"(x), "r"(n));
    uint32_t u = fma_h2(0x03D103D1u, f, 0x09A109A1u);                                  // 2^-12 (0.23843 f + 0.70345)
    u = fma_h2(u, f, 0x0C000C00u);                                                     // ... f + 2^-12 * 1.00044
    asm("{\n\t.reg .b32 e;\n\tand.b32 e, %1, 0x001F001F;\n\tmad.lo.u32 %0, e, 1024, %2;\n\t}" : "=r"(r) : "r"(t), "r"(u));
    return r;
}
__device__ __forceinline__ float sumsq_h2(uint32_t v) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v));
    return f.x * f.x + f.y * f.y;
}
__device__ __forceinline__ void tmem_ld32_pack16_nowait(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// NM of every 8 register pairs take MUFU, the others the FMA pipe
template <int NM> __device__ __forceinline__ constexpr bool on_mufu(int i) {
    const int i8 = i & 7;
    return NM >= 8 ? true : NM == 5 ? (i8 != 1 && i8 != 4 && i8 != 6) : NM == 4 ? (i8 & 1) == 0
         : NM == 3 ? (i8 == 0 || i8 == 3 || i8 == 6) : NM == 2 ? (i8 & 3) == 0 : false;
}

template <int HD, int NM0, int NM1, int NWG, int CTAS>
__global__ void __launch_bounds__((NWG * 5 + 1) * 32, CTAS)
attn_tc16_kernel(const __grid_constant__ CUtensorMap tmap, bf16* __restrict__ out, const float* __restrict__ kmax, int* __restrict__ flags,
                 int* __restrict__ declined, const int* __restrict__ skip, int L, int C) {
    // (image, head) pairs the polynomial-kernel tier (attn_lin.cu) has computed: nothing to do (it cleared their flags)
    if (skip && skip[blockIdx.z * gridDim.y + blockIdx.y] >= 0) return;
    constexpr int KB = HD / 8;
    constexpr int NO = HD == 8 ? 16 : 32;       // PV accumulator columns: head_dim | ones | zero padding
    constexpr int TMEM_COLS = 512 / CTAS;
    constexpr int NSTAGE = nstage(NWG);
    constexpr int W_TMA = NWG * 4, W_MMA = W_TMA + 1;
    constexpr int NTHREADS = (W_MMA + NWG) * 32;
    // one O accumulator per CTA: every PV MMA of every warpgroup adds into it.  (Measured: giving each issuer its own pair of
    // accumulators changes nothing -- the tensor pipe does not serialise these small dependent MMAs.)
    constexpr int STAGE_BLOCKS = 2 + NO / 8;
    constexpr int STAGE_BYTES = STAGE_BLOCKS * BLK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char* stages = smem;
    unsigned char* qs = stages + NSTAGE * STAGE_BYTES;
    uint64_t* kv_full = reinterpret_cast<uint64_t*>(qs + 2 * QBLK);
    uint64_t* kv_free = kv_full + NSTAGE;
    uint64_t* s_full = kv_free + NSTAGE;                    // [NWG]
    uint64_t* s_free = s_full + NWG;                        // [NWG]
    uint64_t* p_full = s_free + NWG;                        // [NWG]
    uint64_t* p_free = p_full + NWG;                        // [NWG]
    uint64_t* q_full = p_free + NWG;
    uint64_t* o_full = q_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
    uint32_t* bound_slot = tmem_slot + 1;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the issuers' operands stay in uniform registers
    const int b = blockIdx.z, h = blockIdx.y, row0 = blockIdx.x * TQ;
    const int T = L / TK;

    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 1); }
        for (int g = 0; g < NWG; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 128);
            mbar_init(&p_full[g], 128); mbar_init(&p_free[g], 1);
        }
        mbar_init(q_full, 1);
        mbar_init(o_full, T < NWG ? T : NWG);
        *bound_slot = 0u;
        fence_barrier_init();
    }
    {
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u), one4 = make_uint4(0x00003C00u, 0u, 0u, 0u);   // f16 1.0 at n = head_dim
        if (HD == 8) for (int i = tid; i < QBLK / 16; i += NTHREADS) reinterpret_cast<uint4*>(qs + QBLK)[i] = z4;
        for (int i = tid; i < NSTAGE * (BLK / 16); i += NTHREADS) {
            const int s = i / (BLK / 16), r = i % (BLK / 16);
            unsigned char* st = stages + s * STAGE_BYTES;
            if (HD == 8) reinterpret_cast<uint4*>(st + BLK)[r] = z4;                 // K hi
            reinterpret_cast<uint4*>(st + (2 + KB) * BLK)[r] = one4;                 // ones block after the V blocks
            if (HD == 16) reinterpret_cast<uint4*>(st + (3 + KB) * BLK)[r] = z4;     // pad PV's N to 32
        }
    }
    fence_proxy_async();
    if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: NWG S buffers of TK (one f16 score per 32-bit cell) | NWG P buffers of TK/2 (f16 pairs) | O
    const uint32_t tmem_p = tmem_base + NWG * TK;
    const uint32_t tmem_o = tmem_p + NWG * (TK / 2);
    static_assert(NWG * TK + NWG * (TK / 2) + 32 <= TMEM_COLS, "TMEM budget");
    if (warp < 4) {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int c = 0; c < NO / 16; ++c) tmem_st16(tmem_o + ((uint32_t)(warp * 32) << 16) + c * 16, z);
        tmem_wait_st();
        tc_fence_before();
    }

    if (tid == 0) {
        mbar_expect_tx(q_full, KB * QBLK);
        for (int kb = 0; kb < KB; ++kb) {
            tma_load_3d(qs + kb * QBLK, &tmap, q_full, h * HD + 8 * kb, row0, b);
            tma_load_3d(qs + kb * QBLK + BLK, &tmap, q_full, h * HD + 8 * kb, row0 + TK, b);
        }
    }
    mbar_wait(q_full, 0);
    // logit bound of the CTA: max over its rows of |q'_i| * max_j |k_j|
    int direct;
    {
        const int r = tid & 127;
        float qq = 0.f;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const uint4 v = *reinterpret_cast<const uint4*>(qs + kb * QBLK + r * 16);
            qq += sumsq_h2(v.x) + sumsq_h2(v.y) + sumsq_h2(v.z) + sumsq_h2(v.w);
        }
        float bound = sqrtf(qq) * kmax[b * gridDim.y + h] * 1.0001f;
        if (!(bound >= 0.f)) bound = __int_as_float(0x7f800000);                      // NaN -> declined
        const uint32_t wmax = __reduce_max_sync(0xffffffffu, __float_as_uint(bound));  // non-negative floats order like their bits
        if (lane == 0) atomicMax(bound_slot, wmax);
        __syncthreads();
        const float cta_bound = __uint_as_float(*bound_slot);
        const int bad = !(cta_bound <= BOUND_F16);
        if (tid == 0) {
            flags[((long long)b * gridDim.y + h) * gridDim.x + blockIdx.x] = bad;
            if (bad) atomicAdd(declined, 1);
        }
        if (bad) {
            if (warp == W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
            return;
        }
        direct = cta_bound <= BOUND_DIRECT;
    }

    constexpr uint32_t DESC_HI_K = (uint32_t)(128 >> 4) | (1u << 14);          // SBO = 128 B (next 8 rows)
    constexpr uint32_t DESC_HI_V = (uint32_t)(BLK >> 4) | (1u << 14);          // SBO = one block (next 8 output columns)
    if (warp == W_TMA) {
        // ---- TMA producer -------------------------------------------------------------------------------------------
        const bool leader = elect_one();
        int s = 0, ph = 1;
        for (int j = 0; j < T; ++j) {
            mbar_wait_warp<200>(&kv_free[s], ph);
            if (leader) {
                unsigned char* st = stages + s * STAGE_BYTES;
                mbar_expect_tx(&kv_full[s], 2 * KB * BLK);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_load_3d(st + kb * BLK, &tmap, &kv_full[s], C + h * HD + 8 * kb, j * TK, b);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_load_3d(st + (2 + kb) * BLK, &tmap, &kv_full[s], 2 * C + h * HD + 8 * kb, j * TK, b);
            }
            __syncwarp();
            s = (s + 1) & (NSTAGE - 1);
            ph ^= (s == 0);
        }
    } else if (warp >= W_MMA) {
        // ---- MMA issuer of warpgroup g: tiles g, g + NWG, ...; S buffer g, P buffer g -----------------------------------
        // instruction descriptors: D format [4,6) (0 f16, 1 f32) | A format [7,10) (0 f16, 1 bf16) | B format [10,13) | B MN-major bit 16
        constexpr uint32_t IDESC_QK = (0u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
        constexpr uint32_t IDESC_PV = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 16) | ((uint32_t)(NO >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
        const uint64_t desc_q = ((uint64_t)DESC_HI_K << 32) | (((smem_u32(qs) & 0x3FFFF) >> 4) | ((uint32_t)(QBLK >> 4) << 16));
        const uint32_t k_lo0 = ((smem_u32(stages) & 0x3FFFF) >> 4) | ((uint32_t)(BLK >> 4) << 16);
        const uint32_t v_lo0 = ((smem_u32(stages + 2 * BLK) & 0x3FFFF) >> 4) | ((uint32_t)(128 >> 4) << 16);   // LBO = next 8 keys
        const int g = warp - W_MMA;
        const uint32_t tmem_s = tmem_base + g * TK, tmem_pg = tmem_p + g * (TK / 2);
        const bool leader = elect_one();
        auto qk = [&](int j) {      // S(j) -> this warpgroup's buffer
            const int s = j & (NSTAGE - 1);
            mbar_wait_warp<0>(&kv_full[s], (j / NSTAGE) & 1);
            tc_fence_after();
            if (leader) {
                umma_bf16(tmem_s, desc_q, ((uint64_t)DESC_HI_K << 32) | (k_lo0 + s * (STAGE_BYTES >> 4)), IDESC_QK, 0u);
                umma_commit(&s_full[g]);
            }
            __syncwarp();
        };
        if (g < T) qk(g);
        int it = 0;
        for (int j = g; j < T; j += NWG, ++it) {
            const int s = j & (NSTAGE - 1);
            // the warpgroup holds S(j) in registers: refill its buffer with its next tile while it exponentiates
            if (j + NWG < T) {
                mbar_wait_warp<0>(&s_free[g], it & 1);
                qk(j + NWG);
            }
            mbar_wait_warp<100>(&p_full[g], it & 1);
            tc_fence_after();
            if (leader) {
                const uint32_t v_lo = v_lo0 + s * (STAGE_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < TK / 16; ++ks)
                    umma_bf16_ts(tmem_o, tmem_pg + ks * 8, ((uint64_t)DESC_HI_V << 32) | (v_lo + ks * 16), IDESC_PV, 1u);
                umma_commit(&p_free[g]);
                umma_commit(&kv_free[s]);
                if (j + NWG >= T) umma_commit(&o_full[0]);
            }
            __syncwarp();
        }
    } else {
        // ---- softmax warpgroups: thread = query row -------------------------------------------------------------------
        const int g = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t s_addr = lane_base + g * TK;
        const uint32_t p_addr = lane_base + (tmem_p - tmem_base) + g * (TK / 2);
        auto run = [&](auto mode) {
            constexpr bool DIRECT = decltype(mode)::value;
            constexpr int NM = DIRECT ? NM0 : NM1;
            int it = 0;
            for (int j = g; j < T; j += NWG, ++it) {
                mbar_wait_warp<0>(&s_full[g], it & 1);
                tc_fence_after();
                uint32_t sv[32];
                tmem_ld32_pack16_nowait(s_addr, sv);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&s_free[g]);                  // S is in registers: the MMA warp may refill the buffer
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    sv[i] = on_mufu<NM>(i) ? ex2_h2(sv[i]) : DIRECT ? ex2_direct_h2(sv[i]) : ex2_reduced_h2(sv[i]);
                if (it > 0) {                             // PV of the previous tile has read P
                    mbar_wait_warp<0>(&p_free[g], (it - 1) & 1);
                    tc_fence_after();
                }
                tmem_st32(p_addr, sv);
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(&p_full[g]);
            }
        };
        if (direct) run(std::true_type{}); else run(std::false_type{});
        if (g == 0) {
            // ---- epilogue: O / row sum -> bf16 ---------------------------------------------------------------------------
            mbar_wait(o_full, 0);
            tc_fence_after();
            const uint32_t o_addr = tmem_o + ((uint32_t)((warp & 3) * 32) << 16);
            float o[NO];
#pragma unroll
            for (int c = 0; c < NO / 16; ++c) {
                float v[16];
                tmem_ld16(o_addr + c * 16, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) o[c * 16 + i] = v[i];
            }
            const float inv = 1.f / o[HD];
            const int row = row0 + (warp & 3) * 32 + lane;
            bf16* op = out + ((long long)b * L + row) * C + (long long)h * HD;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                uint4 w;
                uint32_t* wp = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __nv_bfloat162 pr = __floats2bfloat162_rn(o[kb * 8 + 2 * i] * inv, o[kb * 8 + 2 * i + 1] * inv);
                    wp[i] = *reinterpret_cast<const uint32_t*>(&pr);
                }
                *reinterpret_cast<uint4*>(op + kb * 8) = w;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int HD, int NM0, int NM1, int NWG, int CTAS>
int launch16g(const CUtensorMap& tm, void* out, const float* kmax, int* flags, int* declined, const int* skip, int B, int L, int C, int heads,
              cudaStream_t st) {
    constexpr int NO = HD == 8 ? 16 : 32;
    constexpr int NSTAGE = nstage(NWG);
    constexpr int need = NSTAGE * (2 + NO / 8) * BLK + 2 * QBLK + (2 * NSTAGE + 4 * NWG + 2) * 8 + 16 + 128;
    constexpr int floor_bytes = (227 * 1024) / (CTAS + 1) + 1024;     // no more than CTAS CTAs fit an SM (they own 512 / CTAS TMEM columns each)
    constexpr int smem = need > floor_bytes ? need : floor_bytes;
    constexpr int NTHREADS = (NWG * 5 + 1) * 32;
    auto kern = attn_tc16_kernel<HD, NM0, NM1, NWG, CTAS>;
    static PerDevice attr_set;
    if (int& done = attr_set.cur(); !done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { ddpmir_set_error("attention_tc16: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
        done = 1;
    }
    dim3 grid(L / TQ, heads, B);
    kern<<<grid, NTHREADS, smem, st>>>(tm, (bf16*)out, kmax, flags, declined, skip, L, C);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

// geometry: two CTAs of two warpgroups per SM everywhere.  The five-warpgroup CTA (tuning hook: split bit 7) measures the same
// at L = 65 536 (72.9 vs 72.1 ms) and loses below (4.89 vs 4.67 ms at L = 16 384): the kernel is bound by issue slots per
// score, not by the number of warps the schedulers can pick from -- see profiles/r2_ncu_summary.md
template <int HD, int NM0, int NM1>
int launch16(const CUtensorMap& tm, void* out, const float* kmax, int* flags, int* declined, const int* skip, int B, int L, int C, int heads,
             int geom, cudaStream_t st) {
    const bool wide = geom == 2;
    return wide ? launch16g<HD, NM0, NM1, 5, 1>(tm, out, kmax, flags, declined, skip, B, L, C, heads, st)
                : launch16g<HD, NM0, NM1, 2, 2>(tm, out, kmax, flags, declined, skip, B, L, C, heads, st);
}

}  // namespace

// qkv [B, L, 3C] f16 with pre-scaled q; out bf16; kmax [B*heads] from the key-norm pre-pass; flags [B*heads*L/128]: 1 = declined
// (logit bound > 11), to be redone by ddpmir_attention_tc(..., redo = 1) on a bf16 copy; *declined counts them.  split
// (tuning hook): MUFU share of the two FMA-pipe variants in eighths, bits 0-2 = bound <= 2 (degree-4 polynomial; 7 =
// everything on MUFU), bits 3-5 = bound <= 11 (range-reduced); 0 = default.  Bits 6-7: CTA geometry, 2 = one CTA of five
// warpgroups, otherwise two CTAs of two warpgroups per SM.
int ddpmir_attention_tc16(const void* qkv, void* out, const float* kmax, int* flags, int* declined, const int* skip, int B, int L, int C,
                          int heads, int split, cudaStream_t st) {
    const int hd = C / heads;
    if ((hd != 8 && hd != 16) || L % TQ != 0 || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15)) return DDPMIR_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) return DDPMIR_ERR_UNSUPPORTED;
    CUtensorMap tm;
    {
        cuuint64_t dims[3] = {(cuuint64_t)3 * C, (cuuint64_t)L, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)3 * C * 2, (cuuint64_t)L * 3 * C * 2};
        cuuint32_t box[3] = {8, (cuuint32_t)TK, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ddpmir_set_error("attention_tc16: tensor map failed (%d)", (int)r); return DDPMIR_ERR_CUDA; }
    }
    const int nm0 = split & 7, nm1 = (split >> 3) & 7, geom = (split >> 6) & 3;
#define L16(HD) (nm0 == 2 ? launch16<HD, 2, 4>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st) : \
                 nm0 == 4 ? launch16<HD, 4, 4>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st) : \
                 nm0 == 5 ? launch16<HD, 5, 4>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st) : \
                 nm0 == 7 ? launch16<HD, 8, 8>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st) : \
                 nm1 == 3 ? launch16<HD, 3, 3>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st) : \
                 nm1 == 5 ? launch16<HD, 3, 5>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st) : \
                            launch16<HD, 3, 4>(tm, out, kmax, flags, declined, skip, B, L, C, heads, geom, st))
    return hd == 8 ? L16(8) : L16(16);
#undef L16
}
