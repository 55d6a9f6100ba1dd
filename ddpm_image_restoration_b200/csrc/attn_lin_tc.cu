// Polynomial-kernel tier of the bounded-softmax self-attention on tcgen05 / TMEM (see attn_lin.cu for the algebra and the
// pre-pass).  Per (image, head) whose logit bound fits a polynomial set, with phi = the monomial feature map in binary16:
//
//   attn_lin_state_tc_kernel   S^T += Phi(K~)^T [w v | w]     features x keys x (head_dim + 1): the M = 128 rows of a tcgen05.mma are
//                              128 FEATURES, K = 16 keys, N = 16 / 32 value columns.  A warp = 32 keys: every thread walks the
//                              monomial tree of its key in registers (fp32 products, one rounding to f16), writes octets of 8
//                              features as 16-byte rows into the canonical MN-major layout ([feature octet][key][16 B] -- conflict
//                              free, no transpose needed: "M-major A" is exactly "one key's features contiguous"), and after every
//                              128 features its elected lane issues the two MMAs of that chunk itself (no issuer warp, no
//                              cross-warp barrier; the accumulators of all warps add up in TMEM).  Two chunk buffers per warp.
//   attn_lin_reduce_tc_kernel  sums the key slices, applies c_n n!/a! / L, rounds to f16 and lays S^T out as the K-major B operand.
//   attn_lin_out_tc_kernel     O = Phi(Q~) S: thread = query row = TMEM lane.  The features never touch shared memory: octets go
//                              straight into the row's TMEM lane (tcgen05.st) as the A operand of M128 N16/32 K16 MMAs against the
//                              resident S^T; an issuer warp trails the four generator warps by one 128-feature chunk.
// The feature ORDER is whatever the generator below emits (depth first; for the big maps -- head_dim 16 degree 4, head_dim 8
// degree 6 -- first variable in a runtime loop and every (d1, d2) block padded to 8); the coefficient table is produced by
// running the SAME generator in a bookkeeping mode, so the three kernels cannot disagree about it.
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "attn_lin.cuh"

namespace {
using namespace attn_lin;

constexpr int CH_OCT = 16;                 // octets per chunk: 128 features
constexpr int MAXG = 4;                    // feature groups (TMEM holds 512 accumulator columns per CTA)

struct Groups {
    int n;
    int d1lo[MAXG], d1hi[MAXG];            // first-variable range of the group
    int chunks[MAXG];                      // 128-feature chunks of the group
    int base[MAXG];                        // first chunk of the group in the global feature order
    int total;                             // chunks of all groups
};

template <int HD, int DEG> struct Lay {
    static constexpr int F = nfeat(HD, DEG);
    static constexpr bool FULL = F <= 1300;                // everything unrolled, no padding inside the map
    static constexpr int NB = HD == 8 ? 16 : 32;           // value columns: v | 1 | zeros
    static constexpr int HDR = 1 + HD / 8;                 // !FULL: header octets [1, 0 x 7], [x_0 .. x_7] (, [x_8 .. x_15])
    static constexpr int MAXCH = 512 / NB;                 // chunks whose accumulators fit TMEM
    static constexpr int blk(int d2) { return (nfeat(HD - d2, DEG - 2) + 7) / 8; }     // octets of block (., d2)
    static constexpr int sub(int d1) { int s = 0; for (int d2 = d1; d2 < HD; ++d2) s += blk(d2); return s; }
    static constexpr Groups groups() {
        Groups g{};
        if (FULL) {
            g.n = 1; g.d1lo[0] = 0; g.d1hi[0] = HD; g.chunks[0] = ((F + 7) / 8 + CH_OCT - 1) / CH_OCT; g.base[0] = 0;
            g.total = g.chunks[0];
            return g;
        }
        int d1 = 0, base = 0;
        g.n = 0;
        while (d1 < HD) {
            int oct = g.n == 0 ? HDR : 0, lo = d1;
            while (d1 < HD && (oct + sub(d1) + CH_OCT - 1) / CH_OCT <= MAXCH) { oct += sub(d1); ++d1; }
            g.d1lo[g.n] = lo; g.d1hi[g.n] = d1; g.chunks[g.n] = (oct + CH_OCT - 1) / CH_OCT; g.base[g.n] = base;
            base += g.chunks[g.n]; ++g.n;
        }
        g.total = base;
        return g;
    }
    static constexpr int maxchunks() { Groups g = groups(); int m = 0; for (int i = 0; i < g.n; ++i) m = g.chunks[i] > m ? g.chunks[i] : m; return m; }
};

// ---- the feature generator ---------------------------------------------------------------------------------------------------
// Emits octets of 8 consecutive features to `sink.octet(v, lv)`; in bookkeeping mode (COEF) v is n!/a! and lv the degree.
// The monomial tree is walked by TEMPLATE recursion, not loops: every node's position in the feature order is a template
// parameter, so the slot of the octet buffer it lands in and the points where an octet is complete are compile-time facts by
// construction (a loop-carried position counter only becomes constant if the compiler fully unrolls four to six nested
// triangular loops, which it stops doing once the sink's code makes the bodies large).  xrt(d) reads x[d] for a RUNTIME d: the
// big maps keep the first variable in a runtime loop to bound the code size.
template <int HD, int DEG, bool COEF, class Sink>
struct Gen {
    const float (&x)[HD];
    Sink& sink;
    float ov[8];
    int olv[8];

    template <int POS> __device__ __forceinline__ void put(float val, float mult, int lev) {
        ov[POS & 7] = COEF ? mult : val;
        if (COEF) olv[POS & 7] = lev;
        if constexpr ((POS & 7) == 7) sink.octet(ov, olv);
    }
    template <int POS> __device__ __forceinline__ void pad() {          // zeros up to the next octet boundary
        if constexpr ((POS & 7) != 0) { put<POS>(0.f, 0.f, 0); pad<POS + 1>(); }
    }
    // the node "parent * x[D]" at level LEV and position POS, its subtree (extensions by variables >= D, depth first), and -- with
    // SIB -- its siblings D + 1 .. at the same level.  RUN + run_extra = multiplicity of x[D] in the node (bookkeeping mode).
    template <int LEV, int D, int POS, int RUN, bool SIB>
    __device__ __forceinline__ void node(float mparent, float multparent, int run_extra) {
        const float m = mparent * x[D];
        const float mult = COEF ? multparent * (float)LEV / (float)(RUN + run_extra) : 0.f;
        put<POS>(m, mult, LEV);
        if constexpr (LEV < DEG) node<LEV + 1, D, POS + 1, RUN + 1, true>(m, mult, run_extra);
        if constexpr (SIB && D + 1 < HD) node<LEV, D + 1, POS + nfeat(HD - D, DEG - LEV), 1, true>(mparent, multparent, 0);
    }
    template <int D> __device__ __forceinline__ void header_vars() {
        put<8 + D>(x[D], 1.f, 1);
        if constexpr (D + 1 < HD) header_vars<D + 1>();
    }
    // !FULL: the blocks (d1, D2), D2 >= d1, of first variable d1 (runtime, warp-uniform); every block starts on an octet boundary
    template <int D2> __device__ __forceinline__ void blocks(int d1, float m1) {
        if (D2 >= d1) {
            node<2, D2, 0, 1, false>(m1, 1.f, D2 == d1 ? 1 : 0);
            pad<nfeat(HD - D2, DEG - 2)>();
        }
        if constexpr (D2 + 1 < HD) blocks<D2 + 1>(d1, m1);
    }
};

template <int HD, int DEG, bool COEF, class Sink, class XRt>
__device__ __forceinline__ void generate(const float (&x)[HD], XRt&& xrt, int d1lo, int d1hi, bool with_header, Sink& sink) {
    Gen<HD, DEG, COEF, Sink> g{x, sink, {}, {}};
    if constexpr (Lay<HD, DEG>::FULL) {
        g.template put<0>(1.f, 1.f, 0);
        g.template node<1, 0, 1, 1, true>(1.f, 1.f, 0);
        g.template pad<nfeat(HD, DEG)>();
    } else {
        if (with_header) {
            g.template put<0>(1.f, 1.f, 0);
            g.template pad<1>();
            g.template header_vars<0>();
        }
#pragma unroll 1
        for (int d1 = d1lo; d1 < d1hi; ++d1) g.template blocks<0>(d1, xrt(d1));
    }
    sink.finish();
}

__device__ __forceinline__ uint4 pack_octet(const float (&v)[8]) {
    uint4 r;
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
    r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
    return r;
}
__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

// chunk boundary of the state kernel: this warp's 128 features x 32 keys are in shared memory -> the elected lane multiplies them
// into the accumulators (two K = 16 MMAs), commits to the buffer's barrier, and the warp waits until the OTHER buffer (the one
// it fills next) has been read by the MMAs issued from it two chunks ago.
template <int NB>
__device__ __noinline__ void state_publish(unsigned char* chunk, uint32_t vaddr, uint32_t tmem_d, uint64_t* bar_this, uint64_t* bar_next,
                                           int uses_next, int lane) {
    constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        const uint32_t a0 = smem_u32(chunk);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tmem_d, make_desc_noswz(a0 + ks * 256, 128, 512), make_desc_noswz(vaddr + ks * 256, 128, 512), IDESC, 1u);
        umma_commit(bar_this);
    }
    __syncwarp();
    if (uses_next > 0) mbar_wait_warp<0>(bar_next, (uses_next - 1) & 1);
}
// chunk boundary of the output kernel: the thread's 128 features are in its TMEM lane -> arrive on the chunk's barrier, then wait
// until the MMAs of chunk n - 2 have read the buffer that is filled next.  n = chunks published before this one.
__device__ __noinline__ void out_publish(uint64_t* full_this, uint64_t* free_next, int n) {
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(full_this);
    if (n >= 1) mbar_wait_warp<0>(free_next, ((n - 1) >> 1) & 1);
}

// ---- coefficient table: n!/a! and degree of every (padded) feature, in the generator's order -------------------------------------
template <int HD, int DEG> struct Table {
    static constexpr int FP = Lay<HD, DEG>::groups().total * CH_OCT * 8;
};
template <int HD, int DEG> __device__ float g_mult[Table<HD, DEG>::FP];
template <int HD, int DEG> __device__ unsigned char g_deg[Table<HD, DEG>::FP];

template <int HD, int DEG>
__global__ void attn_lin_table_kernel(Groups gi) {
    const int g = threadIdx.x;
    if (g >= gi.n) return;
    struct Sink {
        float* mult; unsigned char* deg; int oct, end;
        __device__ void octet(const float (&v)[8], const int (&lv)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { mult[oct * 8 + i] = v[i]; deg[oct * 8 + i] = (unsigned char)lv[i]; }
            ++oct;
        }
        __device__ void finish() {
            for (; oct < end; ++oct)
                for (int i = 0; i < 8; ++i) { mult[oct * 8 + i] = 0.f; deg[oct * 8 + i] = 0; }
        }
    } sink{g_mult<HD, DEG>, g_deg<HD, DEG>, gi.base[g] * CH_OCT, (gi.base[g] + gi.chunks[g]) * CH_OCT};
    float x[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) x[d] = 1.f;
    generate<HD, DEG, true>(x, [](int) { return 1.f; }, gi.d1lo[g], gi.d1hi[g], g == 0, sink);
}

// ---- S^T partial sums: features x keys on the tensor core ---------------------------------------------------------------------------
template <int HD, int DEG> struct StateTc {
    using L_ = Lay<HD, DEG>;
    static constexpr int NB = L_::NB, NBO = NB / 8;
    static constexpr int CHUNK_BYTES = CH_OCT * 32 * 16;             // [16 feature octets][32 keys][16 B]
    static constexpr int V_BYTES = NBO * 32 * 16;                    // [value octets][32 keys][16 B]
    static constexpr int XS_BYTES = L_::FULL ? 0 : 32 * HD * 4;
    static constexpr int WARP_BYTES = 2 * CHUNK_BYTES + 2 * V_BYTES + XS_BYTES;
    static constexpr int SMEM = 4 * WARP_BYTES + 128 + 128;
    static constexpr int tmem_cols() { int c = L_::maxchunks() * NB, p = 32; while (p < c) p *= 2; return p; }
};

template <int HD, int DEG>
__global__ void __launch_bounds__(128)
attn_lin_state_tc_kernel(const __half* __restrict__ qkv, const int* __restrict__ tier, const float* __restrict__ params,
                         float* __restrict__ spart, int L, int C, int keys_per_split, Groups gi) {
    using T = StateTc<HD, DEG>;
    constexpr int NB = T::NB, NBO = T::NBO, TCOLS = T::tmem_cols();
    const int H = gridDim.y / gi.n, h = blockIdx.y / gi.n, g = blockIdx.y % gi.n, b = blockIdx.z, split = blockIdx.x;
    const int set = tier[b * H + h];
    if (set < 0 || set_degree(set) != DEG) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    unsigned char* wbase = smem + warp * T::WARP_BYTES;
    unsigned char* chunk0 = wbase;                                   // 2 chunk buffers
    unsigned char* vbuf0 = wbase + 2 * T::CHUNK_BYTES;               // 2 value buffers (key tile parity)
    float* xs = reinterpret_cast<float*>(vbuf0 + 2 * T::V_BYTES);    // !FULL: this warp's scaled keys for runtime-indexed reads
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * T::WARP_BYTES);      // [4 warps][2 buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nch = gi.chunks[g];
    {   // clear this warp's lane quadrant of the accumulators
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        for (int c = 0; c < nch * NB / 16; ++c) tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 16, z);
        tmem_wait_st();
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();

    const long long rstride = 3LL * C;
    const __half* base = qkv + (long long)b * L * rstride + (long long)h * HD;
    const int j0 = split * keys_per_split, j1 = min(L, j0 + keys_per_split);
    const float* par = params + (long long)(b * H + h) * PSTRIDE;
    uint64_t* my_bar = bars + warp * 2;

    // (the chunk-boundary work is a separate, non-inlined function: inlined at every octet it pushes the unrolled feature walk
    //  over the compiler's full-unroll size limit, and the walk then falls back to runtime loops over register arrays)
    struct Sink {
        unsigned char* chunk0; uint64_t* bar; uint32_t tmem_base, vaddr; int lane, oct, buf, uses0, uses1, nch_oct;
        __device__ __forceinline__ void octet(const float (&v)[8], const int (&)[8]) {
            *reinterpret_cast<uint4*>(chunk0 + buf * StateTc<HD, DEG>::CHUNK_BYTES + (oct & (CH_OCT - 1)) * 512 + lane * 16) = pack_octet(v);
            ++oct;
            if ((oct & (CH_OCT - 1)) == 0) {
                const int uses = buf == 0 ? ++uses0 : ++uses1;
                (void)uses;
                const int nb = buf ^ 1;
                state_publish<NB>(chunk0 + buf * StateTc<HD, DEG>::CHUNK_BYTES, vaddr, tmem_base + (uint32_t)((oct / CH_OCT - 1) * NB), &bar[buf], &bar[nb],
                                  nb == 0 ? uses0 : uses1, lane);
                buf = nb;
            }
        }
        __device__ __forceinline__ void finish() {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int dummy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
            while (oct < nch_oct) octet(z, dummy);
        }
    } sink{chunk0, my_bar, tmem_base, 0u, lane, 0, 0, 0, 0, nch * CH_OCT};

    uint4 kraw[HD / 8], vraw[HD / 8];
    auto fetch = [&](int jt) {
        const int j = jt + warp * 32 + lane;
        if (j < j1) {
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                kraw[c] = __ldg(reinterpret_cast<const uint4*>(base + (long long)j * rstride + C) + c);
                vraw[c] = __ldg(reinterpret_cast<const uint4*>(base + (long long)j * rstride + 2 * C) + c);
            }
        }
    };
    fetch(j0);
    int tile = 0;
    for (int jt = j0; jt < j1; jt += 128, ++tile) {
        const bool live = jt + warp * 32 + lane < j1;
        float x[HD];
        {
            float k[HD], v[HD];
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) { unpack8h(kraw[c], k + 8 * c); unpack8h(vraw[c], v + 8 * c); }
            float e = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) {
                const float kc = live ? k[d] - par[P_B + d] : 0.f;
                e = fmaf(par[P_A + d], kc, e);
                x[d] = kc * par[P_DI + d];
            }
            const float w = live ? exp2f(e - par[P_EOFF]) : 0.f;        // <= 1: w v cannot overflow binary16
            unsigned char* vb = vbuf0 + (tile & 1) * T::V_BYTES;
            float wv[8];
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
#pragma unroll
                for (int i = 0; i < 8; ++i) wv[i] = w * v[8 * c + i];
                *reinterpret_cast<uint4*>(vb + c * 512 + lane * 16) = pack_octet(wv);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) wv[i] = i == 0 ? w : 0.f;
            *reinterpret_cast<uint4*>(vb + (HD / 8) * 512 + lane * 16) = pack_octet(wv);
            if (NBO > HD / 8 + 1) *reinterpret_cast<uint4*>(vb + (HD / 8 + 1) * 512 + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
            if constexpr (!Lay<HD, DEG>::FULL) {
#pragma unroll
                for (int d = 0; d < HD; ++d) xs[lane * HD + d] = x[d];
            }
            sink.vaddr = smem_u32(vb);
        }
        fetch(jt + 128);
        sink.oct = 0;
        generate<HD, DEG, false>(x, [&](int d) { return xs[lane * HD + d]; }, gi.d1lo[g], gi.d1hi[g], g == 0, sink);
    }
    // drain: the last MMAs of both buffers
    if (sink.uses0 > 0) mbar_wait_warp<0>(&my_bar[0], (sink.uses0 - 1) & 1);
    if (sink.uses1 > 0) mbar_wait_warp<0>(&my_bar[1], (sink.uses1 - 1) & 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // epilogue: thread = feature row of every chunk
    const int FPAD = gi.total * 128;
    float* dst = spart + (((long long)(b * H + h)) * gridDim.x + split) * ((long long)FPAD * NB);
    for (int c = 0; c < nch; ++c) {
        const int f = (gi.base[g] + c) * 128 + warp * 32 + lane;
#pragma unroll
        for (int q = 0; q < NB / 16; ++q) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c * NB + q * 16, v);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(dst + (long long)f * NB + q * 16 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

// S^T (f16, K-major B operand: [feature octet][value octet][8 values][8 features]) = c_n n!/a! / L * sum over the key slices
template <int HD, int DEG>
__global__ void __launch_bounds__(256)
attn_lin_reduce_tc_kernel(const float* __restrict__ spart, const int* __restrict__ tier, __half* __restrict__ ST, int splits, float inv_l,
                          int total_chunks) {
    constexpr int NB = Lay<HD, DEG>::NB;
    const int bh = blockIdx.y, chunk = blockIdx.x;
    const int set = tier[bh];
    if (set < 0 || set_degree(set) != DEG) return;
    const long long FPAD = (long long)total_chunks * 128;
    const float* src = spart + (long long)bh * splits * (FPAD * NB) + (long long)chunk * 128 * NB;
    __half* dst = ST + (long long)bh * (FPAD * NB) + (long long)chunk * 128 * NB;
    for (int i = threadIdx.x; i < 128 * NB; i += 256) {
        const int fl = i / NB, col = i % NB, f = chunk * 128 + fl;
        float a = 0.f;
        for (int s = 0; s < splits; ++s) a += src[(long long)s * (FPAD * NB) + i];
        a *= g_mult<HD, DEG>[f] * set_coef(set, g_deg<HD, DEG>[f]) * inv_l;
        a = fminf(fmaxf(a, -65504.f), 65504.f);
        dst[((fl >> 3) * (NB / 8) + (col >> 3)) * 64 + (col & 7) * 8 + (fl & 7)] = __float2half_rn(a);
    }
}

// ---- output: rows x features on the tensor core, features generated straight into TMEM -----------------------------------------------
template <int HD, int DEG> struct OutTc {
    using L_ = Lay<HD, DEG>;
    static constexpr int NB = L_::NB;
    static constexpr int ST_BYTES = L_::maxchunks() * 128 * NB * 2;
    static constexpr int QS_BYTES = L_::FULL ? 0 : 128 * HD * 4;
    static constexpr int SMEM = ST_BYTES + QS_BYTES + 128 + 128;
    static constexpr int TCOLS = 256;                                 // 2 x 64 (feature chunks) + NB accumulator columns
};

template <int HD, int DEG>
__global__ void __launch_bounds__(160, 2)      // 256 TMEM columns per CTA: two CTAs per SM
attn_lin_out_tc_kernel(const __half* __restrict__ qkv, const int* __restrict__ tier, const float* __restrict__ params,
                       const __half* __restrict__ ST, bf16* __restrict__ out, int* __restrict__ flags, int L, int C, Groups gi) {
    using T = OutTc<HD, DEG>;
    constexpr int NB = T::NB, TCOLS = T::TCOLS;
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    const int set = tier[b * H + h];
    if (set < 0 || set_degree(set) != DEG) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char* sT = smem;
    float* qs = reinterpret_cast<float*>(smem + T::ST_BYTES);          // !FULL: [HD][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + T::ST_BYTES + T::QS_BYTES);   // [2]
    uint64_t* freeb = full + 2;                                        // [2]
    uint64_t* gdone = freeb + 2;                                       // [1] all MMAs of a group have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gdone + 1);
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        mbar_init(&full[0], 128); mbar_init(&full[1], 128);
        mbar_init(&freeb[0], 1); mbar_init(&freeb[1], 1);
        mbar_init(gdone, 1);
        fence_barrier_init();
        flags[((long long)(b * H + h)) * (L / 128) + blockIdx.x] = 0;      // the quadratic tiers skip this tile
    }
    if (warp == 4) tmem_alloc(tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base + 128;
    const float* par = params + (long long)(b * H + h) * PSTRIDE;
    const long long FPAD = (long long)gi.total * 128;
    const __half* STg = ST + (long long)(b * H + h) * (FPAD * NB);
    const int row = blockIdx.x * 128 + tid;

    float x[HD];
    if (warp < 4) {
        const uint4* qp = reinterpret_cast<const uint4*>(qkv + ((long long)b * L + row) * (3LL * C) + (long long)h * HD);
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) unpack8h(__ldg(qp + c), &x[c * 8]);
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            x[d] = (x[d] - par[P_A + d]) * par[P_D + d];
            if constexpr (!Lay<HD, DEG>::FULL) qs[d * 128 + tid] = x[d];
        }
    }
    struct Sink {
        uint64_t* full; uint64_t* freeb; uint32_t taddr0; int oct, buf, n;      // n: chunks published so far (all groups)
        __device__ __forceinline__ void octet(const float (&v)[8], const int (&)[8]) {
            tmem_st4(taddr0 + buf * 64 + (oct & (CH_OCT - 1)) * 4, pack_octet(v));
            ++oct;
            if ((oct & (CH_OCT - 1)) == 0) {
                out_publish(&full[buf], &freeb[buf ^ 1], n);       // chunk n is complete; chunk n + 1 goes to the other buffer
                ++n;
                buf ^= 1;
            }
        }
        int end_oct;
        __device__ __forceinline__ void finish() {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int dummy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
            while (oct < end_oct) octet(z, dummy);
        }
    } sink{full, freeb, tmem_base + ((uint32_t)((warp & 3) * 32) << 16), 0, 0, 0, 0};

    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int n_issued = 0;
    for (int g = 0; g < gi.n; ++g) {
        // the previous group's MMAs have read S^T: load this group's slice (generic proxy -> fence -> async proxy)
        if (g > 0) mbar_wait(gdone, (g - 1) & 1);
        {
            const uint4* src = reinterpret_cast<const uint4*>(STg + (long long)gi.base[g] * 128 * NB);
            uint4* dst = reinterpret_cast<uint4*>(sT);
            const int n16 = gi.chunks[g] * 128 * NB * 2 / 16;
            for (int i = tid; i < n16; i += 160) dst[i] = __ldg(src + i);
        }
        fence_proxy_async();
        __syncthreads();
        if (warp < 4) {
            sink.oct = 0;
            sink.end_oct = gi.chunks[g] * CH_OCT;
            generate<HD, DEG, false>(x, [&](int d) { return qs[d * 128 + tid]; }, gi.d1lo[g], gi.d1hi[g], g == 0, sink);
        } else {
            const bool leader = elect_one();
            const uint32_t s0 = smem_u32(sT);
            for (int c = 0; c < gi.chunks[g]; ++c, ++n_issued) {
                const int buf = n_issued & 1;
                mbar_wait_warp<0>(&full[buf], (n_issued >> 1) & 1);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
                        umma_bf16_ts(tmem_d, tmem_base + buf * 64 + ks * 8,
                                     make_desc_noswz(s0 + (uint32_t)((c * 16 + ks * 2) * (NB / 8) * 128), (NB / 8) * 128, 128), IDESC,
                                     (n_issued > 0 || ks > 0) ? 1u : 0u);
                    umma_commit(&freeb[buf]);
                    if (c == gi.chunks[g] - 1) umma_commit(gdone);
                }
                __syncwarp();
            }
        }
    }
    if (warp < 4) {
        mbar_wait(gdone, (gi.n - 1) & 1);
        tc_fence_after();
        float o[NB];
#pragma unroll
        for (int q = 0; q < NB / 16; ++q) {
            float v[16];
            tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + q * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[q * 16 + i] = v[i];
        }
        const float inv = 1.f / o[HD];
        bf16* op = out + ((long long)b * L + row) * C + (long long)h * HD;
#pragma unroll
        for (int c8 = 0; c8 < HD / 8; ++c8) {
            uint4 w;
            __nv_bfloat162* wp = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
            for (int i = 0; i < 4; ++i) wp[i] = __floats2bfloat162_rn(o[c8 * 8 + 2 * i] * inv, o[c8 * 8 + 2 * i + 1] * inv);
            *reinterpret_cast<uint4*>(op + c8 * 8) = w;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, TCOLS);
}

int tc_splits(int B, int L, int heads, int ngroups) {
    int s = 592 / (B * heads * ngroups);
    const int cap = L / 2048;
    if (s > cap) s = cap;
    return s < 1 ? 1 : s;
}
template <int HD, int DEG> constexpr size_t tc_bytes_per_bh(int splits) {
    constexpr int FP = Lay<HD, DEG>::groups().total * 128, NB = Lay<HD, DEG>::NB;
    return (size_t)FP * NB * 2 + (size_t)splits * FP * NB * 4;
}

template <int HD, int DEG>
int launch_tc(const __half* qkv, const int* tier, const float* params, void* ws, bf16* out, int* flags, int B, int L, int C, int heads,
              cudaStream_t st) {
    constexpr Groups gi = Lay<HD, DEG>::groups();
    constexpr int FP = gi.total * 128, NB = Lay<HD, DEG>::NB;
    const int splits = tc_splits(B, L, heads, gi.n);
    const int kps = (ceil_div(L, splits) + 127) / 128 * 128;
    __half* ST = (__half*)ws;
    float* spart = (float*)((char*)ws + ((size_t)B * heads * FP * NB * 2 + 255) / 256 * 256);
    static PerDevice ready;
    if (int& done = ready.cur(); !done) {
        cudaError_t e = cudaFuncSetAttribute(attn_lin_state_tc_kernel<HD, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, StateTc<HD, DEG>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_lin_out_tc_kernel<HD, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, OutTc<HD, DEG>::SMEM);
        if (e != cudaSuccess) { ddpmir_set_error("attention_lin_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
        attn_lin_table_kernel<HD, DEG><<<1, 32, 0, st>>>(gi);          // once per device: stream-ordered before its first reader
        DDPMIR_LAUNCH_CHECK();
        done = 1;
    }
    attn_lin_state_tc_kernel<HD, DEG><<<dim3(splits, heads * gi.n, B), 128, StateTc<HD, DEG>::SMEM, st>>>(qkv, tier, params, spart, L, C, kps, gi);
    DDPMIR_LAUNCH_CHECK();
    attn_lin_reduce_tc_kernel<HD, DEG><<<dim3(gi.total, B * heads), 256, 0, st>>>(spart, tier, ST, splits, 1.f / (float)L, gi.total);
    DDPMIR_LAUNCH_CHECK();
    attn_lin_out_tc_kernel<HD, DEG><<<dim3(L / 128, heads, B), 160, OutTc<HD, DEG>::SMEM, st>>>(qkv, tier, params, ST, out, flags, L, C, gi);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

}  // namespace

size_t ddpmir_attention_lin_tc_workspace(int B, int L, int hd, int heads) {
    size_t m = 0;
    auto upd = [&](size_t st_bytes, size_t part_bytes) {
        const size_t t = (st_bytes + 255) / 256 * 256 + part_bytes + 256;
        if (t > m) m = t;
    };
    const size_t bh = (size_t)B * heads;
#define SZ(HD, DEG) { constexpr Groups gi = Lay<HD, DEG>::groups(); constexpr size_t FP = (size_t)gi.total * 128, NB = Lay<HD, DEG>::NB; \
                      upd(bh * FP * NB * 2, bh * tc_splits(B, L, heads, gi.n) * FP * NB * 4); }
    if (hd == 8) { SZ(8, 3) SZ(8, 4) SZ(8, 5) SZ(8, 6) }
    else if (hd == 16) { SZ(16, 3) SZ(16, 4) }
#undef SZ
    return m;
}

int ddpmir_attention_lin_tc(const void* qkv, void* out, const int* tier, const float* params, void* ws, int* flags, int B, int L, int C,
                            int heads, int max_set, cudaStream_t st) {
    const int hd = C / heads;
    if ((hd != 8 && hd != 16) || L % 128 != 0) return DDPMIR_ERR_UNSUPPORTED;
    const __half* q = (const __half*)qkv;
    int rc = DDPMIR_OK;
#define LT(HD, DEG) if (rc == DDPMIR_OK && set_degree(max_set) >= DEG) rc = launch_tc<HD, DEG>(q, tier, params, ws, (bf16*)out, flags, B, L, C, heads, st)
    if (hd == 8) { LT(8, 3); LT(8, 4); LT(8, 5); LT(8, 6); }
    else { LT(16, 3); LT(16, 4); }
#undef LT
    return rc;
}
