// Bounded-softmax self-attention on tcgen05 / TMEM / TMA (inference, head_dim 8 / 16, L % 128 == 0): the full-resolution
// attention blocks of the UNet (L = H*W up to 65 536 tokens), which are ~80 % of a sampler timestep.
//
// Why a second kernel next to attn_mma.cu's mma.sync one: at head_dim 8/16 the exp2 pipe (16 /clk/SM) is the bound, and
// on the mma.sync path the HMMA / LDSM / fragment shuffling compete with it for issue slots.  Here the tensor work
// leaves the instruction stream entirely:
//   * one CTA = 128 query rows of one (image, head) = the 128 TMEM lanes; K/V tiles of 64 keys stream through a TMA ring;
//   * S = Q' K^T by ONE tcgen05.mma (M128 N64 K16) per tile into TMEM (q already carries log2(e)/sqrt(hd); head_dim 8
//     is padded to K = 16 with a zero block that the shared-memory descriptor's leading-dimension offset points at);
//   * two softmax warpgroups per CTA (two CTAs per SM), each owning one S buffer and one P buffer, read their row with
//     tcgen05.ld (thread = row, so nothing is ever reduced across threads) and release the S buffer at once -- its refill
//     (the warpgroup's next tile) runs under the exponentiation, so the MMA / barrier round trip is off the critical
//     path --, evaluate P = exp2(s') -- MUFU for every other pair of columns, packed-bf16 FMA/ALU arithmetic for the
//     others (ex2_pair_bf16) -- and write P as packed bf16 (tcgen05.st);
//   * O += P V by tcgen05.mma with A = P read straight from TMEM and B = the V tile in its natural [key][dim] layout
//     (MN-major descriptor); V is widened by a constant ones column, so the row sums of P accumulate in TMEM column
//     head_dim for free;
//   * no maximum, no rescale (see attn_mma.cu, "bounded softmax"): the offset is 0, rows whose Cauchy-Schwarz logit bound
//     exceeds the fp32-safe window make the CTA set its flag and leave; the exact kernel redoes those CTAs.
// Per pair of scores the softmax threads issue two exp2 and a pack, or six FMA/ALU instructions, plus 1/16 of a TMEM load/store.
#include <cuda.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TQ = 128;                         // query rows per CTA
constexpr int TK = 64;                          // keys per tile
constexpr int NWG = 2;                          // softmax warpgroups per CTA
constexpr int CTAS_PER_SM = 2;                  // two CTAs share an SM (and its 512 TMEM columns): one CTA's prologue, pipeline
                                                // fill and epilogue run under the other's main loop
constexpr int TMEM_COLS = 512 / CTAS_PER_SM;
constexpr int NBUF = 1;                         // S buffers per warpgroup (released as soon as S sits in registers, so one suffices)
constexpr int LAG = NWG * NBUF;                 // S/P buffers in flight
constexpr int NSTAGE = 8;                       // K/V ring (a stage is released by the PV of its tile); a power of two: the stage arithmetic is masks
// single-role warps after the softmax warps.  One tcgen05.mma costs its issuing warp ~75 cycles of dependent uniform-datapath
// instructions (measured), so each softmax warpgroup gets its own MMA-issuer warp: it issues PV for a buffer as soon as the
// warpgroup has written P and, right behind it in the same in-order tensor pipe (no barrier needed), the QK^T that refills
// that buffer with the warpgroup's tile after next.
constexpr int W_TMA = NWG * 4, W_MMA = W_TMA + 1;
constexpr int NTHREADS = (W_MMA + NWG) * 32;
constexpr int BLK = TK * 16;                    // TK rows x 16 bytes: TK/8 core matrices of 8 rows
constexpr int QBLK = TQ * 16;
constexpr float BOUND_LIMIT = 60.f;             // same window as attn_mma.cu

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
template <int DEG> __device__ __forceinline__ float ex2_poly_bounded(float x) {
    const float M = 12582912.f;
    const float t = x + M;
    const float n = t - M;
    const float f = x - n;
    const float p = DEG == 3 ? fmaf(fmaf(fmaf(0.05517166f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f)
                             : fmaf(fmaf(0.23842894f, f, 0.70344800f), f, 1.00044310f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;\n" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
    return r;
}
// exp2 of a PAIR of scores on the FMA / ALU pipes, produced directly as packed bf16 (DEG == 1): no MUFU, no fp32 polynomial,
// no separate pack -- 7 instructions per pair, 2 of them on the (half-rate) ALU pipe.  |x| <= 60 (the kernel's logit bound).
//   t = x + (2^16 + 64) lies in [2^16, 2^17), where an fp32 ulp is 2^-7: the low 16 bits of its encoding are N * 128 + F with
//       N = floor(x) + 64 (7 bits) and F the 7-bit fraction of x (round to nearest) -- the exponent and mantissa fields of a
//       bf16 -- and the high 16 bits are the constant 0x4780 = 143 * 128;
//   w = t1 * 2^16 + t0 (one IMAD) therefore holds N0 * 128 + F0 in its low half and (N1 + 143) * 128 + F1 in its high half;
//   y = 1.F in [1, 2) for both halves by one LOP3; q = 2^(y-1) by a degree-2 minimax polynomial in two packed bf16 FMAs
//       (max 0.57 %, rms 0.25 % relative error including the bf16 arithmetic, zero mean: the order of truncating an exact
//       value to bf16, which the MUFU pairs do), with the coefficients of the low half pre-scaled by 2^63 and those of the
//       high half by 2^-80, so that
//   r = q + w - y (one IADD3 on the whole word) lands on q * 2^(N - 64) in both halves: the fraction bits of w and y cancel,
//       the exponent offsets (63 - 127 + 0 and -80 - 127 + 143) both come to -64.
__device__ __forceinline__ uint32_t ex2_pair_bf16(float x0, float x1) {
    uint32_t t0, t1, w, q;
    // both magic adds in one packed fp32 instruction (sm_100 add.f32x2 on a 64-bit register pair)
    asm("{\n\t.reg .b64 xx, kk, tt;\n\t"
        "mov.b64 xx, {%2, %3};\n\t"
        "mov.b64 kk, {%4, %4};\n\t"
        "add.rn.f32x2 tt, xx, kk;\n\t"
        "mov.b64 {%0, %1}, tt;\n\t}"
        : "=r"(t0), "=r"(t1) : "f"(x0), "f"(x1), "f"(65600.f));
    asm("mad.lo.u32 %0, %1, 65536, %2;\n" : "=r"(w) : "r"(t1), "r"(t0));
    uint32_t y;            // (w & 0x007F007F) | 0x3F803F80 as ONE lop3: both masks must be register operands for that
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;\n" : "=r"(y) : "r"(w), "r"(0x007F007Fu), "r"(0x3F803F80u));
    // 0.33789 y^2 - 0.016602 y + 0.67969 (bf16 values of the minimax coefficients), low | high halves scaled as above
    asm("{\n\t.reg .b32 u;\n\t"
        "fma.rn.bf16x2 u, %1, %2, %3;\n\t"
        "fma.rn.bf16x2 %0, u, %2, %4;\n\t}"
        : "=r"(q) : "r"(0x16AD5E2Du), "r"(y), "r"(0x9488DC08u), "r"(0x172E5EAEu));
    return q + w - y;
}
__device__ __forceinline__ float sumsq_bf16x2(uint32_t v) {
    const float lo = __uint_as_float(v << 16), hi = __uint_as_float(v & 0xffff0000u);
    return lo * lo + hi * hi;
}

// POLY of every 8 score pairs take the FMA-pipe exp2
template <int HD, int POLY, int DEG>
__global__ void __launch_bounds__(NTHREADS, CTAS_PER_SM)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap, bf16* __restrict__ out, const float* __restrict__ kmax, int* __restrict__ flags,
               int L, int C, int redo) {
    // redo: only the CTAs the half-precision tier (attn_tc16.cu) declined -- flag 1 -- are computed here
    if (redo && flags[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] == 0) return;
    constexpr int KB = HD / 8;                  // 16-byte blocks per Q / K row that hold data (the MMA always reads 2: K = 16)
    constexpr int NO = HD == 8 ? 16 : 32;       // PV accumulator columns: head_dim | ones | zero padding
    constexpr int STAGE_BLOCKS = 2 + NO / 8;    // K lo, K hi (zeros at head_dim 8) | V blocks, ones block (, zero block)
    constexpr int STAGE_BYTES = STAGE_BLOCKS * BLK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char* stages = smem;
    unsigned char* qs = stages + NSTAGE * STAGE_BYTES;      // Q lo, Q hi (zeros at head_dim 8), 128 rows each
    uint64_t* kv_full = reinterpret_cast<uint64_t*>(qs + 2 * QBLK);
    uint64_t* kv_free = kv_full + NSTAGE;
    uint64_t* s_full = kv_free + NSTAGE;                    // [LAG]  QK^T of the tile in this buffer has retired
    uint64_t* s_free = s_full + LAG;                        // [LAG]  the warpgroup has read S into registers: the buffer may be refilled
    uint64_t* p_full = s_free + LAG;                        // [NWG]  the warpgroup has written P
    uint64_t* p_free = p_full + NWG;                        // [NWG]  PV has consumed P
    uint64_t* q_full = p_free + NWG;
    uint64_t* o_full = q_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: the issuers' operands stay in uniform registers
    const int b = blockIdx.z, h = blockIdx.y, row0 = blockIdx.x * TQ;
    const int T = L / TK;

    // ---- one-time setup: barriers, constant blocks, TMEM, Q tile ------------------------------------------------
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_free[s], 1); }
        for (int g = 0; g < LAG; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_free[g], 128); }
        for (int g = 0; g < NWG; ++g) { mbar_init(&p_full[g], 128); mbar_init(&p_free[g], 1); }
        mbar_init(q_full, 1);
        mbar_init(o_full, T < NWG ? T : NWG);
        fence_barrier_init();
    }
    {
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u), one4 = make_uint4(0x00003F80u, 0u, 0u, 0u);   // bf16 1.0 at n = head_dim
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
    // TMEM columns: LAG S buffers of TK | NWG P buffers of TK/2 (bf16 pairs) | O
    const uint32_t tmem_p = tmem_base + LAG * TK;
    const uint32_t tmem_o = tmem_p + NWG * (TK / 2);
    // the PV MMAs of different tiles come from different warps in no fixed order, so all of them accumulate and the
    // accumulator (head_dim | row sum | padding) is cleared here, before the barrier below
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
        for (int kb = 0; kb < KB; ++kb) {       // the tensor map's box is TK rows
            tma_load_3d(qs + kb * QBLK, &tmap, q_full, h * HD + 8 * kb, row0, b);
            tma_load_3d(qs + kb * QBLK + BLK, &tmap, q_full, h * HD + 8 * kb, row0 + TK, b);
        }
    }
    mbar_wait(q_full, 0);
    // logit bound of this thread's row (thread t looks at row t % 128)
    {
        const int r = tid & 127;
        float qq = 0.f;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const uint4 v = *reinterpret_cast<const uint4*>(qs + kb * QBLK + r * 16);
            qq += sumsq_bf16x2(v.x) + sumsq_bf16x2(v.y) + sumsq_bf16x2(v.z) + sumsq_bf16x2(v.w);
        }
        const float bound = sqrtf(qq) * kmax[b * gridDim.y + h] * 1.0001f;
        const int bad = __syncthreads_or(!(bound <= BOUND_LIMIT));
        if (tid == 0) flags[((long long)b * gridDim.y + h) * gridDim.x + blockIdx.x] = bad;
        if (bad) {
            if (warp == W_MMA) tmem_dealloc(tmem_base, TMEM_COLS);
            return;
        }
    }

    // descriptor halves that never change: version 1, no swizzle; the low word is (address >> 4) | (LBO >> 4) << 16
    constexpr uint32_t DESC_HI_K = (uint32_t)(128 >> 4) | (1u << 14);          // SBO = 128 B (next 8 rows)
    constexpr uint32_t DESC_HI_V = (uint32_t)(BLK >> 4) | (1u << 14);          // SBO = one block (next 8 output columns)
    if (warp == W_TMA) {
        // ---- TMA producer -------------------------------------------------------------------------------------------
        const bool leader = elect_one();
        int s = 0, ph = 1;
        for (int j = 0; j < T; ++j) {
            mbar_wait_warp<200>(&kv_free[s], ph);      // plenty of slack (a ring of tiles ahead): sleep between probes
            if (leader) {
                unsigned char* st = stages + s * STAGE_BYTES;
                mbar_expect_tx(&kv_full[s], 2 * KB * BLK);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_load_3d(st + kb * BLK, &tmap, &kv_full[s], C + h * HD + 8 * kb, j * TK, b);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_load_3d(st + (2 + kb) * BLK, &tmap, &kv_full[s], 2 * C + h * HD + 8 * kb, j * TK, b);
            }
            __syncwarp();
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
    } else if (warp >= W_MMA) {
        // ---- MMA issuer of warpgroup g: tiles g, g + NWG, ...; buffer g + NWG * (it & 1) ------------------------------------
        constexpr uint32_t IDESC_QK = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
        constexpr uint32_t IDESC_PV = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NO >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
        const uint64_t desc_q = ((uint64_t)DESC_HI_K << 32) | (((smem_u32(qs) & 0x3FFFF) >> 4) | ((uint32_t)(QBLK >> 4) << 16));
        const uint32_t k_lo0 = ((smem_u32(stages) & 0x3FFFF) >> 4) | ((uint32_t)(BLK >> 4) << 16);
        const uint32_t v_lo0 = ((smem_u32(stages + 2 * BLK) & 0x3FFFF) >> 4) | ((uint32_t)(128 >> 4) << 16);   // LBO = next 8 keys
        const int g = warp - W_MMA;
        const bool leader = elect_one();
        auto qk = [&](int j, int sb) {      // S(j) -> buffer sb
            const int s = j % NSTAGE;
            mbar_wait_warp<0>(&kv_full[s], (j / NSTAGE) & 1);
            tc_fence_after();
            if (leader) {
                umma_bf16(tmem_base + sb * TK, desc_q, ((uint64_t)DESC_HI_K << 32) | (k_lo0 + s * (STAGE_BYTES >> 4)), IDESC_QK, 0u);
                umma_commit(&s_full[sb]);
            }
            __syncwarp();
        };
        for (int i = 0; i < NBUF; ++i)
            if (g + NWG * i < T) qk(g + NWG * i, g + NWG * i);
        int it = 0;
        for (int j = g; j < T; j += NWG, ++it) {
            const int sb = g + NWG * (it % NBUF), s = j % NSTAGE;
            // the warpgroup holds S(j) in registers: refill its buffer with the tile NBUF rounds ahead while it exponentiates
            if (j + NWG * NBUF < T) {
                mbar_wait_warp<0>(&s_free[sb], (it / NBUF) & 1);
                qk(j + NWG * NBUF, sb);
            }
            mbar_wait_warp<100>(&p_full[g], it & 1);    // P arrives most of a tile later: sleep between probes
            tc_fence_after();
            if (leader) {
                const uint32_t v_lo = v_lo0 + s * (STAGE_BYTES >> 4);
#pragma unroll
                for (int ks = 0; ks < TK / 16; ++ks)
                    umma_bf16_ts(tmem_o, tmem_p + g * (TK / 2) + ks * 8, ((uint64_t)DESC_HI_V << 32) | (v_lo + ks * 16), IDESC_PV, 1u);
                umma_commit(&p_free[g]);        // P may be overwritten once these MMAs retire
                umma_commit(&kv_free[s]);       // ... and K and V of this stage are consumed
                if (j + NWG >= T) umma_commit(&o_full[0]);   // this warp's last tile
            }
            __syncwarp();
        }
    } else {
        // ---- softmax warpgroups: thread = query row -------------------------------------------------------------------
        const int g = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        auto exp_pack = [&](const uint32_t (&sv)[32], uint32_t (&pk)[16]) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int i8 = i & 7;
                const bool poly = POLY == 1 ? i8 == 5 : POLY == 2 ? (i8 & 3) == 3 : POLY == 3 ? (i8 == 2 || i8 == 5 || i8 == 7)
                                : POLY == 4 ? (i8 & 1) == 1 : POLY == 5 ? (i8 != 0 && i8 != 3 && i8 != 6) : false;
                const float x0 = __uint_as_float(sv[2 * i]), x1 = __uint_as_float(sv[2 * i + 1]);
                if (poly && DEG == 1) { pk[i] = ex2_pair_bf16(x0, x1); continue; }
                const float e0 = poly ? ex2_poly_bounded<DEG == 1 ? 2 : DEG>(x0) : ex2f(x0);
                const float e1 = poly ? ex2_poly_bounded<DEG == 1 ? 2 : DEG>(x1) : ex2f(x1);
                pk[i] = pack_bf16_trunc(e0, e1);
            }
        };
        const uint32_t p_addr = lane_base + (tmem_p - tmem_base) + g * (TK / 2);
        int it = 0;
        for (int j = g; j < T; j += NWG, ++it) {
            const int sb = g + NWG * (it % NBUF);
            const uint32_t s_addr = lane_base + sb * TK;
            mbar_wait_warp<0>(&s_full[sb], (it / NBUF) & 1);
            tc_fence_after();
            uint32_t sv[TK / 32][32], pk[TK / 32][16];
#pragma unroll
            for (int c = 0; c < TK / 32; ++c) tmem_ld32_nowait(s_addr + c * 32, sv[c]);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(&s_free[sb]);                 // S is in registers: the MMA warp may refill the buffer
#pragma unroll
            for (int c = 0; c < TK / 32; ++c) exp_pack(sv[c], pk[c]);
            if (it > 0) {                             // PV of the previous tile has read P
                mbar_wait_warp<0>(&p_free[g], (it - 1) & 1);
                tc_fence_after();
            }
#pragma unroll
            for (int c = 0; c < TK / 32; ++c) tmem_st16(p_addr + c * 16, pk[c]);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[g]);
        }
        if (g == 0) {
            // ---- epilogue: O / row sum -> bf16 ---------------------------------------------------------------------------
            mbar_wait(o_full, 0);
            tc_fence_after();
            const uint32_t o_addr = tmem_o + ((uint32_t)((warp & 3) * 32) << 16);
            float o[NO];
            {
                float v[16];
                tmem_ld16(o_addr, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = v[i];
                if (NO == 32) {
                    tmem_ld16(o_addr + 16, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[(NO == 32 ? 16 : 0) + i] = v[i];
                }
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

template <int HD, int POLY, int DEG>
int launch(const CUtensorMap& tm, void* out, const float* kmax, int* flags, int B, int L, int C, int heads, int redo, cudaStream_t st) {
    constexpr int NO = HD == 8 ? 16 : 32;
    // a CTA owns 256 of the 512 TMEM columns: ask for more than a third of the shared memory so that no third CTA lands on the SM
    constexpr int need = NSTAGE * (2 + NO / 8) * BLK + 2 * QBLK + (2 * NSTAGE + 2 * LAG + 2 * NWG + 2) * 8 + 16 + 128;
    constexpr int floor_bytes = (227 * 1024) / (CTAS_PER_SM + 1) + 1024;     // no more than CTAS_PER_SM CTAs fit an SM
    constexpr int smem = need > floor_bytes ? need : floor_bytes;
    static_assert(LAG * TK + NWG * (TK / 2) + 32 <= TMEM_COLS, "TMEM budget");
    auto kern = attn_tc_kernel<HD, POLY, DEG>;
    static PerDevice attr_set;
    if (int& done = attr_set.cur(); !done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { ddpmir_set_error("attention_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
        done = 1;
    }
    dim3 grid(L / TQ, heads, B);
    kern<<<grid, NTHREADS, smem, st>>>(tm, (bf16*)out, kmax, flags, L, C, redo);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

}  // namespace

// qkv [B, L, 3C] bf16 with pre-scaled q; kmax [B*heads] from the key-norm pre-pass; flags [B*heads*L/128].
// sel: bits 0-2 = score pairs of 8 on the FMA pipe, bits 3-4 = how: 0 degree-3 fp32 polynomial, 1 degree 2, 2 packed bf16 pairs.
// redo != 0: flags holds the half-precision tier's verdicts; only CTAs with flag 1 run (and overwrite it with their own).
int ddpmir_attention_tc(const void* qkv, void* out, const float* kmax, int* flags, int B, int L, int C, int heads, int sel, int redo, cudaStream_t st) {
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
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ddpmir_set_error("attention_tc: tensor map failed (%d)", (int)r); return DDPMIR_ERR_CUDA; }
    }
    const int poly = sel & 7, mode = (sel >> 3) & 3;       // mode 0: degree-3 fp32 polynomial, 1: degree 2, 2: packed bf16 pairs
#define GP(HD, DEG) (poly == 0 ? launch<HD, 0, DEG>(tm, out, kmax, flags, B, L, C, heads, redo, st) : \
                     poly == 1 ? launch<HD, 1, DEG>(tm, out, kmax, flags, B, L, C, heads, redo, st) : \
                     poly == 2 ? launch<HD, 2, DEG>(tm, out, kmax, flags, B, L, C, heads, redo, st) : \
                     poly == 3 ? launch<HD, 3, DEG>(tm, out, kmax, flags, B, L, C, heads, redo, st) : \
                     poly == 4 ? launch<HD, 4, DEG>(tm, out, kmax, flags, B, L, C, heads, redo, st) : \
                                 launch<HD, 5, DEG>(tm, out, kmax, flags, B, L, C, heads, redo, st))
#define GT(HD) (mode == 2 ? GP(HD, 1) : mode == 1 ? GP(HD, 2) : GP(HD, 3))
    return hd == 8 ? GT(8) : GT(16);
#undef GT
#undef GP
}
