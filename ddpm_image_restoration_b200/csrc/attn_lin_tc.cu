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
    int pad[MAXG];                         // zero octets that fill the group's last chunk
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
            g.pad[0] = g.chunks[0] * CH_OCT - (F + 7) / 8;
            g.total = g.chunks[0];
            return g;
        }
        int d1 = 0, base = 0;
        g.n = 0;
        while (d1 < HD) {
            int oct = g.n == 0 ? HDR : 0, lo = d1;
            while (d1 < HD && (oct + sub(d1) + CH_OCT - 1) / CH_OCT <= MAXCH) { oct += sub(d1); ++d1; }
            g.d1lo[g.n] = lo; g.d1hi[g.n] = d1; g.chunks[g.n] = (oct + CH_OCT - 1) / CH_OCT; g.base[g.n] = base;
            g.pad[g.n] = g.chunks[g.n] * CH_OCT - oct;
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
// %laneid / %tid.x through volatile asm: the compiler treats special-register reads as free to rematerialise and re-reads them (S2R,
// ~25 cycles on a slow pipe) at every use inside the unrolled feature walk; an opaque value stays in its register
__device__ __forceinline__ int opaque_lane() { int v; asm volatile("mov.u32 %0, %%laneid;" : "=r"(v)); return v; }
__device__ __forceinline__ int opaque_tid() { int v; asm volatile("mov.u32 %0, %%tid.x;" : "=r"(v)); return v; }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
}

// ---- chunk boundaries, out of line -------------------------------------------------------------------------------------------------
// The unrolled feature walk is thousands of instructions long; whatever is inlined at each of its ~100 octet sites is paid for
// in instruction-cache misses (measured: 3.5 stall cycles per issued instruction with the boundary work inlined).  So an octet
// site is: pack, one 16-byte store through a 32-bit address, add, compare, predicated call.  Everything a boundary needs lives
// in a small per-warp context in shared memory; the non-inlined boundary function gets its address and returns the address of
// the next octet.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 q, c, 0x400000;\n\t"
        "@!q trap;\n\t"
        "nanosleep.u32 20;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t"
        "}"
        :: "r"(bar), "r"(parity) : "memory");
}

// state kernel, per-warp context (8 words): chunk buffers | barriers | accumulator base | value tile | n | c
//   n = chunks published so far (whole kernel): chunk n lives in buffer n % NBUF;  c = chunks published for the current key tile
constexpr int SC_CHUNK0 = 0, SC_BAR = 4, SC_TMEM = 8, SC_VADDR = 12, SC_N = 16, SC_C = 20, SC_BYTES = 32;
// This warp's 128 features x 32 keys are in shared memory: the elected lane multiplies them into the accumulators (two K = 16
// MMAs) and commits to the buffer's barrier; then the warp waits until the buffer it fills next has been read.
template <int NB, int NBUF, int CHUNK_BYTES>
__device__ __noinline__ uint32_t state_boundary(uint32_t ctx, int lane) {
    constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t chunk0 = lds32(ctx + SC_CHUNK0), bar = lds32(ctx + SC_BAR);
    const int n = (int)lds32(ctx + SC_N);
    const int buf = n % NBUF, nxt = (n + 1) % NBUF;
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        const uint32_t a0 = chunk0 + buf * CHUNK_BYTES, vaddr = lds32(ctx + SC_VADDR);
        const int c = (int)lds32(ctx + SC_C);
        const uint32_t d = lds32(ctx + SC_TMEM) + (uint32_t)(c * NB);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            umma_bf16(d, make_desc_noswz(a0 + ks * 256, 128, 512), make_desc_noswz(vaddr + ks * 256, 128, 512), IDESC, 1u);
        umma_commit_addr(bar + buf * 8);
        sts32(ctx + SC_N, (uint32_t)(n + 1));
        sts32(ctx + SC_C, (uint32_t)(c + 1));
    }
    __syncwarp();
    // chunk n + 1 reuses the buffer of chunk n + 1 - NBUF: wait for that chunk's commit (the ((n + 1) / NBUF)-th use of the buffer)
    if (n + 1 >= NBUF) mbar_wait_addr(bar + nxt * 8, (uint32_t)((((n + 1) / NBUF) - 1) & 1));
    return chunk0 + nxt * CHUNK_BYTES + lane * 16;
}

// output kernel, per-warp context: full barriers | free barriers | lane-quadrant TMEM base | n (chunks published so far)
constexpr int OC_FULL = 0, OC_FREE = 4, OC_TADDR0 = 8, OC_N = 12, OC_BYTES = 16;
// The thread's 128 features are in its TMEM lane: arrive on the chunk's barrier, then wait until the MMAs of chunk n - 1 have read
// the buffer that is filled next.  Returns the TMEM address of the next octet.
__device__ __noinline__ uint32_t out_boundary(uint32_t ctx, int lane) {
    const int n = (int)lds32(ctx + OC_N);
    const uint32_t full = lds32(ctx + OC_FULL), freeb = lds32(ctx + OC_FREE);
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive_addr(full + (n & 1) * 8);
    __syncwarp();
    if (lane == 0) sts32(ctx + OC_N, (uint32_t)(n + 1));
    __syncwarp();
    if (n >= 1) mbar_wait_addr(freeb + ((n + 1) & 1) * 8, (uint32_t)(((n - 1) >> 1) & 1));
    return lds32(ctx + OC_TADDR0) + ((n + 1) & 1) * 64;
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

// Key slices per (image, head): chosen ON THE DEVICE from how many (image, head) pairs the pre-pass gave this degree, so that a
// handful of pairs still fills the machine (every thread of the state and reduce kernels derives the same number).
constexpr int TC_ITEMS_TARGET = 592;
__host__ __device__ inline int tc_splits_for(int count, int ngroups, int L) {
    if (count <= 0) return 1;
    int s = TC_ITEMS_TARGET / (count * ngroups);
    const int cap = L / 2048 < 32 ? L / 2048 : 32;
    if (s > cap) s = cap;
    return s < 1 ? 1 : s;
}
// partial-sum slots a call can need: count * splits <= count + TC_ITEMS_TARGET / ngroups
__host__ __device__ inline int tc_slots(int n_bh, int ngroups) { return n_bh + TC_ITEMS_TARGET / ngroups + 1; }

// ---- S^T partial sums: features x keys on the tensor core ---------------------------------------------------------------------------
template <int HD, int DEG> struct StateTc {
    using L_ = Lay<HD, DEG>;
    static constexpr int NB = L_::NB, NBO = NB / 8;
    static constexpr int CHUNK_BYTES = CH_OCT * 32 * 16;             // [16 feature octets][32 keys][16 B]
    static constexpr int V_BYTES = NBO * 32 * 16;                    // [value octets][32 keys][16 B]
    static constexpr int XS_BYTES = L_::FULL ? 0 : 32 * HD * 4;
    // chunk buffers per warp: two when several CTAs share the SM (their warps hide the latency between an MMA and its commit),
    // four when the accumulators take all of TMEM and the SM holds a single CTA of four warps
    static constexpr int NBUF = L_::maxchunks() * NB > 256 ? 4 : 2;
    static constexpr int WARP_BYTES = NBUF * CHUNK_BYTES + 2 * V_BYTES + XS_BYTES;
    static constexpr int SMEM = 4 * WARP_BYTES + 4 * SC_BYTES + 256 + 128;
    static constexpr int tmem_cols() { int c = L_::maxchunks() * NB, p = 32; while (p < c) p *= 2; return p; }
};

// Persistent over the work items (list entry of this degree, feature group, key slice).
template <int HD, int DEG>
__global__ void __launch_bounds__(128)
attn_lin_state_tc_kernel(const __half* __restrict__ qkv, const int* __restrict__ counts, const int* __restrict__ lists, int n_bh_total,
                         const float* __restrict__ params, float* __restrict__ spart, int L, int C, int H, Groups gi) {
    using T = StateTc<HD, DEG>;
    constexpr int NB = T::NB, NBO = T::NBO, TCOLS = T::tmem_cols();
    const int count = counts[DEG];
    const int splits = tc_splits_for(count, gi.n, L);
    const int n_items = count * gi.n * splits;
    if ((int)blockIdx.x >= n_items) return;
    const int* list = lists + (long long)DEG * n_bh_total;
    const int keys_per_split = ((L + splits - 1) / splits + 127) / 128 * 128;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int tid = opaque_tid(), lane = opaque_lane();
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    unsigned char* wbase = smem + warp * T::WARP_BYTES;
    unsigned char* chunk0 = wbase;                                   // 2 chunk buffers
    unsigned char* vbuf0 = wbase + T::NBUF * T::CHUNK_BYTES;         // 2 value buffers (key tile parity)
    float* xs = reinterpret_cast<float*>(vbuf0 + 2 * T::V_BYTES);    // !FULL: this warp's scaled keys for runtime-indexed reads
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * T::WARP_BYTES);      // [4 warps][NBUF buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * T::NBUF);
    const uint32_t ctx = smem_u32(tmem_slot + 4) + warp * SC_BYTES;            // this warp's boundary context

    if (tid == 0) {
        for (int i = 0; i < 4 * T::NBUF; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint64_t* my_bar = bars + warp * T::NBUF;

    if (lane == 0) {
        sts32(ctx + SC_CHUNK0, smem_u32(chunk0)); sts32(ctx + SC_BAR, smem_u32(my_bar)); sts32(ctx + SC_TMEM, tmem_base);
        sts32(ctx + SC_N, 0u); sts32(ctx + SC_C, 0u);
    }
    __syncwarp();
    struct Sink {
        uint32_t wptr, wend, ctx; int lane, pad;        // where the next octet goes / end of the chunk being filled
        __device__ __forceinline__ void octet(const float (&v)[8], const int (&)[8]) {
            sts128(wptr, pack_octet(v));
            wptr += 512;
            if (wptr == wend) {
                wptr = state_boundary<NB, StateTc<HD, DEG>::NBUF, StateTc<HD, DEG>::CHUNK_BYTES>(ctx, lane);
                wend = wptr + CH_OCT * 512;
            }
        }
        __device__ __forceinline__ void finish() {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int dummy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
            for (int i = 0; i < pad; ++i) octet(z, dummy);
        }
    } sink{0u, 0u, ctx, lane, 0};

    const long long rstride = 3LL * C;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int li = item / (gi.n * splits), g = (item / splits) % gi.n, split = item % splits;
    const int bh = list[li], b = bh / H, h = bh % H;
    const int nch = gi.chunks[g];
    sink.pad = gi.pad[g];
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
    const __half* base = qkv + (long long)b * L * rstride + (long long)h * HD;
    const int j0 = split * keys_per_split, j1 = min(L, j0 + keys_per_split);
    const float* par = params + (long long)bh * PSTRIDE;

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
            if (lane == 0) { sts32(ctx + SC_VADDR, smem_u32(vb)); sts32(ctx + SC_C, 0u); }
            __syncwarp();
        }
        fetch(jt + 128);
        sink.wptr = smem_u32(chunk0) + ((int)lds32(ctx + SC_N) % T::NBUF) * T::CHUNK_BYTES + lane * 16;
        sink.wend = sink.wptr + CH_OCT * 512;
        generate<HD, DEG, false>(x, [&](int d) { return xs[lane * HD + d]; }, gi.d1lo[g], gi.d1hi[g], g == 0, sink);
    }
    // drain: the last commit of every buffer
    {
        const int n = (int)lds32(ctx + SC_N);
        for (int i = 0; i < T::NBUF; ++i) {
            const int last = n - 1 - i;                 // chunk numbers n-1, n-2, ... cover all buffers
            if (last >= 0) mbar_wait_warp<0>(&my_bar[last % T::NBUF], (last / T::NBUF) & 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // epilogue: thread = feature row of every chunk
    const int FPAD = gi.total * 128;
    float* dst = spart + ((long long)li * splits + split) * ((long long)FPAD * NB);      // slot = (list position, slice)
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
    __syncthreads();            // every warp has read its accumulators: the next item may clear them
    tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

// S^T (f16, K-major B operand: [feature octet][value octet][8 values][8 features]) = c_n n!/a! / L * sum over the key slices
template <int HD, int DEG>
__global__ void __launch_bounds__(256)
attn_lin_reduce_tc_kernel(const float* __restrict__ spart, const int* __restrict__ tier, const int* __restrict__ counts,
                          const int* __restrict__ lists, int n_bh_total, __half* __restrict__ ST, int L, int ngroups, int total_chunks) {
    constexpr int NB = Lay<HD, DEG>::NB;
    const int li = blockIdx.y, chunk = blockIdx.x;
    const int count = counts[DEG];
    if (li >= count) return;
    const int splits = tc_splits_for(count, ngroups, L);
    const int bh = lists[(long long)DEG * n_bh_total + li];
    const int set = tier[bh];
    const float inv_l = 1.f / (float)L;
    const long long FPAD = (long long)total_chunks * 128;
    const float* src = spart + (long long)li * splits * (FPAD * NB) + (long long)chunk * 128 * NB;
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
    static constexpr int SMEM = ST_BYTES + QS_BYTES + 4 * OC_BYTES + 128 + 128;
    static constexpr int TCOLS = 256;                                 // 2 x 64 (feature chunks) + NB accumulator columns
};

// Persistent: the grid is a fixed two CTAs per SM; every CTA walks the work items (an (image, head) of this degree from the
// pre-pass's list x a 128-row tile) with stride gridDim.x.  A degree nobody uses costs a few hundred CTAs that leave at once,
// not a CTA per tile of the whole batch.
template <int HD, int DEG>
__global__ void __launch_bounds__(160, 2)      // 256 TMEM columns per CTA: two CTAs per SM
attn_lin_out_tc_kernel(const __half* __restrict__ qkv, const int* __restrict__ counts, const int* __restrict__ lists, int n_bh_total,
                       const float* __restrict__ params, const __half* __restrict__ ST, bf16* __restrict__ out, int* __restrict__ flags,
                       int L, int C, int H, Groups gi) {
    using T = OutTc<HD, DEG>;
    constexpr int NB = T::NB, TCOLS = T::TCOLS;
    const int tiles = L / 128;
    const long long n_items = (long long)counts[DEG] * tiles;
    if ((long long)blockIdx.x >= n_items) return;
    const int* list = lists + (long long)DEG * n_bh_total;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char* sT = smem;
    float* qs = reinterpret_cast<float*>(smem + T::ST_BYTES);          // !FULL: [HD][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + T::ST_BYTES + T::QS_BYTES);   // [2]
    uint64_t* freeb = full + 2;                                        // [2]
    uint64_t* gdone = freeb + 2;                                       // [1] all MMAs of a group have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gdone + 1);
    const uint32_t ctx0 = smem_u32(tmem_slot + 4);                     // [4 generator warps] boundary contexts
    const int tid = opaque_tid();
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        mbar_init(&full[0], 128); mbar_init(&full[1], 128);
        mbar_init(&freeb[0], 1); mbar_init(&freeb[1], 1);
        mbar_init(gdone, 1);
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base + 128;
    const long long FPAD = (long long)gi.total * 128;

    const int lane = opaque_lane();
    const uint32_t ctx = ctx0 + (warp & 3) * OC_BYTES;
    if (warp < 4 && lane == 0) {
        sts32(ctx + OC_FULL, smem_u32(full)); sts32(ctx + OC_FREE, smem_u32(freeb));
        sts32(ctx + OC_TADDR0, tmem_base + ((uint32_t)(warp * 32) << 16)); sts32(ctx + OC_N, 0u);
    }
    __syncwarp();
    struct Sink {
        uint32_t taddr, ctx; int lane, pad;     // where the next octet goes (4 columns each; a chunk buffer = 64 columns, aligned)
        __device__ __forceinline__ void octet(const float (&v)[8], const int (&)[8]) {
            tmem_st4(taddr, pack_octet(v));
            taddr += 4;
            if ((taddr & 63u) == 0u) taddr = out_boundary(ctx, lane);
        }
        __device__ __forceinline__ void finish() {
            const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int dummy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
            for (int i = 0; i < pad; ++i) octet(z, dummy);
        }
    } sink{0u, ctx, lane, 0};

    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int n_issued = 0;          // issuer warp: chunks multiplied so far
    int gcount = 0;            // groups completed so far (every thread counts them the same way)
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int bh = list[item / tiles], tile = (int)(item % tiles);
        const int b = bh / H, h = bh % H;
        const float* par = params + (long long)bh * PSTRIDE;
        const __half* STg = ST + (long long)bh * (FPAD * NB);
        const int row = tile * 128 + tid;
        if (tid == 0) flags[(long long)bh * tiles + tile] = 0;          // the quadratic tiers skip this tile
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
        bool first_mma = true;
        for (int g = 0; g < gi.n; ++g, ++gcount) {
            // the previous group's MMAs have read S^T: load this group's slice (generic proxy -> fence -> async proxy)
            if (gcount > 0) mbar_wait(gdone, (gcount - 1) & 1);
            {
                const uint4* src = reinterpret_cast<const uint4*>(STg + (long long)gi.base[g] * 128 * NB);
                uint4* dst = reinterpret_cast<uint4*>(sT);
                const int n16 = gi.chunks[g] * 128 * NB * 2 / 16;
                for (int i = tid; i < n16; i += 160) dst[i] = __ldg(src + i);
            }
            fence_proxy_async();
            __syncthreads();
            if (warp < 4) {
                sink.pad = gi.pad[g];
                sink.taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + ((int)lds32(ctx + OC_N) & 1) * 64;
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
                                         (first_mma && ks == 0) ? 0u : 1u);
                        umma_commit(&freeb[buf]);
                        if (c == gi.chunks[g] - 1) umma_commit(gdone);
                    }
                    first_mma = false;
                    __syncwarp();
                }
            }
        }
        if (warp < 4) {
            mbar_wait(gdone, (gcount - 1) & 1);
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
            tc_fence_before();
        }
        __syncthreads();            // the accumulator has been read: the next item's first MMA may overwrite it
        tc_fence_after();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, TCOLS);
}

template <int HD, int DEG>
int launch_tc(const __half* qkv, const int* tier, const int* counts, const int* lists, const float* params, void* ws, bf16* out, int* flags,
              int B, int L, int C, int heads, cudaStream_t st) {
    constexpr Groups gi = Lay<HD, DEG>::groups();
    constexpr int FP = gi.total * 128, NB = Lay<HD, DEG>::NB;
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
    {
        // two to three CTAs per SM where TMEM and shared memory allow it, one where the accumulators take all 512 columns
        const int per_sm = StateTc<HD, DEG>::tmem_cols() > 256 ? 1 : StateTc<HD, DEG>::tmem_cols() > 128 ? 2 : 3;
        const int max_items = B * heads * gi.n * 32;
        const int grid = max_items < per_sm * 148 ? max_items : per_sm * 148;
        attn_lin_state_tc_kernel<HD, DEG><<<grid, 128, StateTc<HD, DEG>::SMEM, st>>>(qkv, counts, lists, B * heads, params, spart, L, C, heads, gi);
        DDPMIR_LAUNCH_CHECK();
    }
    attn_lin_reduce_tc_kernel<HD, DEG><<<dim3(gi.total, B * heads), 256, 0, st>>>(spart, tier, counts, lists, B * heads, ST, L, gi.n, gi.total);
    DDPMIR_LAUNCH_CHECK();
    const long long items = (long long)B * heads * (L / 128);
    const int grid = (int)(items < 2 * 148 ? items : 2 * 148);
    attn_lin_out_tc_kernel<HD, DEG><<<grid, 160, OutTc<HD, DEG>::SMEM, st>>>(qkv, counts, lists, B * heads, params, ST, out, flags, L, C, heads, gi);
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
                      upd(bh * FP * NB * 2, (size_t)tc_slots((int)bh, gi.n) * FP * NB * 4); }
    if (hd == 8) { SZ(8, 3) SZ(8, 4) SZ(8, 5) SZ(8, 6) }
    else if (hd == 16) { SZ(16, 3) SZ(16, 4) }
#undef SZ
    return m;
}

int ddpmir_attention_lin_tc(const void* qkv, void* out, const int* tier, const int* counts, const int* lists, const float* params, void* ws,
                            int* flags, int B, int L, int C, int heads, int max_degree, cudaStream_t st) {
    const int hd = C / heads;
    if ((hd != 8 && hd != 16) || L % 128 != 0) return DDPMIR_ERR_UNSUPPORTED;
    const __half* q = (const __half*)qkv;
    int rc = DDPMIR_OK;
#define LT(HD, DEG) if (rc == DDPMIR_OK && max_degree >= DEG) rc = launch_tc<HD, DEG>(q, tier, counts, lists, params, ws, (bf16*)out, flags, B, L, C, heads, st)
    if (hd == 8) { LT(8, 3); LT(8, 4); LT(8, 5); LT(8, 6); }
    else { LT(16, 3); LT(16, 4); }
#undef LT
    return rc;
}
