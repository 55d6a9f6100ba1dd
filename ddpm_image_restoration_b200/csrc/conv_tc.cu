// tcgen05 / TMEM / TMA implicit-GEMM for the UNet's 3x3 convolutions (9 taps) and 1x1 convolutions / linears
// (1 tap) on NHWC bf16 activations, sm_100a only.
//
//   out[m, n] = epilogue( sum_{tap, c} X[pixel(m) + off(tap), c] * Wt[n, tap*Cin + c] )
//
// * A operand: no im2col buffer.  The activation tensor is described to TMA as a 4-D tensor (C, W, H, B); the 128
//   output pixels of a CTA tile form a box (64 ch, bw, bh, bb) with bw*bh*bb = 128, and each filter tap is the SAME box
//   shifted by (kw-1, kh-1).  Out-of-image coordinates (the zero padding of Conv2d(padding=1)) are zero-filled by TMA,
//   also at image boundaries inside a multi-image box because W/H/B are separate tensor dimensions.
// * The box lands in shared memory as 128 rows x 128 B with the hardware 128B swizzle = a K-major SWIZZLE_128B UMMA
//   operand tile; the weights [N, 9*Cin] (K-major) are loaded the same way.  4-stage mbarrier ring.
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16, bf16 x bf16 -> fp32) into TMEM; tcgen05.commit
//   releases the shared-memory stage / signals the epilogue.
// * Epilogue warps read the accumulator with tcgen05.ld (one TMEM lane = one output pixel per thread), apply the
//   fused epilogue (bias, time-embedding row bias, activation, frequency gates, residual; epilogue.cuh) and store the
//   fp32 stream and/or bf16 operand copy.
#include <cuda.h>
#include "epilogue.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;      // 64 bf16 = 128 B = one swizzle row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int EPI_GROUPS = 2;      // epilogue warp groups of four (one warp per TMEM lane quarter); each takes a share of the columns
constexpr int NUM_THREADS = (2 + 4 * EPI_GROUPS) * 32;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, then the epilogue warps
constexpr int STREAM_EPI_GROUPS = 4;   // persistent kernel: 4 x 4 epilogue warps (2..17), four per TMEM lane quarter, each group
constexpr int STREAM_THREADS = (2 + 4 * STREAM_EPI_GROUPS) * 32;   // taking a share of the output columns -- the epilogue is bound
                                      // by loads in flight per SM

struct TcParams {
    long long M;
    int H, W, Cin, N;
    int bw, bh;            // pixel box = (bw, bh, 128/(bw*bh)) over (W, H, B)
    int taps;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS)
igemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, void* __restrict__ out,
                EpiDev ep, TcParams p) {
    constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // SWIZZLE_128B tiles must start on 1024-byte boundaries of the shared address space
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sa = smem;                                   // STAGES x 16 KB (1024-B aligned tiles)
    unsigned char* sb = smem + STAGES * A_STAGE_BYTES;          // STAGES x B tiles
    uint64_t* full = reinterpret_cast<uint64_t*>(sb + STAGES * B_STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long m0 = (long long)blockIdx.x * BLOCK_M;
    const int n0 = blockIdx.y * BLOCK_N;
    const int kc = p.Cin / BLOCK_K;          // k-blocks per tap
    const int num_kb = p.taps * kc;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer ----
            const int hw = p.H * p.W;
            const int b0 = (int)(m0 / hw);
            const int rem = (int)(m0 - (long long)b0 * hw);
            const int h0 = rem / p.W, w0 = rem - h0 * p.W;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int tap = kb / kc, c0 = (kb - tap * kc) * BLOCK_K;
                const int dh = p.taps == 9 ? tap / 3 - 1 : 0, dw = p.taps == 9 ? tap % 3 - 1 : 0;
                mbar_expect_tx(&full[s], A_STAGE_BYTES + B_STAGE_BYTES);
                tma_load_4d(sa + s * A_STAGE_BYTES, &tmap_a, &full[s], c0, w0 + dw, h0 + dh, b0);
                tma_load_2d(sb + s * B_STAGE_BYTES, &tmap_b, &full[s], tap * p.Cin + c0, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer ----
            // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
            // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sa + s * A_STAGE_BYTES), b_addr = smem_u32(sb + s * B_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k) {
                    const uint64_t da = make_desc_sw128(a_addr + k * 32), db = make_desc_sw128(b_addr + k * 32);
                    umma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty[s]);          // stage reusable once these MMAs have read it
            }
            umma_commit(acc_full);               // accumulator complete
        }
    } else {
        // ---- epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 ----
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int lane_base = 32 * (warp & 3);
        const long long m = m0 + lane_base + lane;
        const bool m_ok = m < p.M;
        EpiRow row;
        if (m_ok) row = epi_row(ep, m);
        constexpr int PER = (BLOCK_N / 16 + EPI_GROUPS - 1) / EPI_GROUPS * 16;     // columns per epilogue group
        const int grp = (warp - 2) >> 2;
#pragma unroll 1
        for (int c0 = grp * PER; c0 < min(BLOCK_N, (grp + 1) * PER); c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);
            if (m_ok) {
                const int n = n0 + c0;
                if (n + 16 <= p.N && (p.N & 15) == 0) {
                    epi_chunk16(ep, row, v, m, n, out);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < p.N) epi_store(ep, out, m, n + j, epi_apply(ep, row, v[j], m, n + j));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

// ---- persistent, weight-resident variant -------------------------------------------------------------------
// For the full-resolution layers (M = B*H*W is huge, N*K small) the whole weight matrix fits in shared memory.  One
// CTA per SM loads it ONCE, then streams activation boxes through a deep TMA ring while looping over its pixel tiles;
// the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct StreamParams {
    TcParams t;
    int num_tiles;     // M tiles of 128 pixels
    int nb;            // weight rows per k-block tile in smem (N rounded up to 8)
    int a_stages;      // depth of the activation ring
};

template <int TCOLS>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
igemm_tc_stream_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       void* __restrict__ out, EpiDev ep, StreamParams sp) {
    const TcParams& p = sp.t;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kc = p.Cin / BLOCK_K;
    const int num_kb = p.taps * kc;
    const int w_tile_bytes = sp.nb * BLOCK_K * 2;
    unsigned char* sw = smem;                                    // num_kb weight tiles
    unsigned char* sa = smem + num_kb * w_tile_bytes;             // a_stages x 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sa + sp.a_stages * A_STAGE_BYTES);
    uint64_t* w_full = bars;
    uint64_t* full = bars + 1;
    uint64_t* empty = full + sp.a_stages;
    uint64_t* acc_full = empty + sp.a_stages;    // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        mbar_init(w_full, 1);
        for (int s = 0; s < sp.a_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4 * STREAM_EPI_GROUPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TCOLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w_full, (uint32_t)(num_kb * w_tile_bytes));
            for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(sw + kb * w_tile_bytes, &tmap_b, w_full, kb * BLOCK_K, 0);
            const int hw = p.H * p.W;
            int it = 0;
            for (int tile = blockIdx.x; tile < sp.num_tiles; tile += gridDim.x) {
                const long long m0 = (long long)tile * BLOCK_M;
                const int b0 = (int)(m0 / hw);
                const int rem = (int)(m0 - (long long)b0 * hw);
                const int h0 = rem / p.W, w0 = rem - h0 * p.W;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % sp.a_stages, ph = (it / sp.a_stages) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    const int tap = kb / kc, c0 = (kb - tap * kc) * BLOCK_K;
                    const int dh = p.taps == 9 ? tap / 3 - 1 : 0, dw = p.taps == 9 ? tap % 3 - 1 : 0;
                    mbar_expect_tx(&full[s], A_STAGE_BYTES);
                    tma_load_4d(sa + s * A_STAGE_BYTES, &tmap_a, &full[s], c0, w0 + dw, h0 + dh, b0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            mbar_wait(w_full, 0);
            int it = 0, i = 0;
            for (int tile = blockIdx.x; tile < sp.num_tiles; tile += gridDim.x, ++i) {
                const int buf = i & 1;
                mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (TCOLS / 2);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % sp.a_stages, ph = (it / sp.a_stages) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sa + s * A_STAGE_BYTES), b_addr = smem_u32(sw + kb * w_tile_bytes);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k)
                        umma_bf16(d_tmem, make_desc_sw128(a_addr + k * 32), make_desc_sw128(b_addr + k * 32), idesc,
                                  (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else {
        const int lane_base = 32 * (warp & 3);
        int i = 0;
        for (int tile = blockIdx.x; tile < sp.num_tiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            mbar_wait(&acc_full[buf], (i >> 1) & 1);
            tc_fence_after();
            const long long m = (long long)tile * BLOCK_M + lane_base + lane;
            const bool m_ok = m < p.M;
            EpiRow row;
            if (m_ok) row = epi_row(ep, m);
            const uint32_t t_addr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * (TCOLS / 2));
            // epilogue group (warp - 2) / 4 takes its share of the 16-column chunks
            const int chunks = p.N / 16, per = (chunks + STREAM_EPI_GROUPS - 1) / STREAM_EPI_GROUPS, grp = (warp - 2) >> 2;
            const int c_lo = min(p.N, grp * per * 16), c_hi = min(p.N, (grp + 1) * per * 16);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                float v[16];
                tmem_ld16(t_addr + (uint32_t)c0, v);
                if (m_ok) epi_chunk16(ep, row, v, m, c0, out);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TCOLS);
}

// ---- host side ---------------------------------------------------------------------------------------------
bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

template <int BLOCK_N, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, void* out, const EpiDev& ep, const TcParams& p, cudaStream_t st) {
    constexpr int smem = STAGES * (A_STAGE_BYTES + BLOCK_N * BLOCK_K * 2) + (2 * STAGES + 1) * 8 + 16 + 1024;
    static PerDevice attr_set;
    if (int& done = attr_set.cur(); !done) {
        cudaError_t e = cudaFuncSetAttribute(igemm_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { ddpmir_set_error("igemm_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
        done = 1;
    }
    dim3 grid(ceil_div(p.M, BLOCK_M), ceil_div(p.N, BLOCK_N));
    igemm_tc_kernel<BLOCK_N, STAGES><<<grid, NUM_THREADS, smem, st>>>(ta, tb, out, ep, p);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

template <int TCOLS>
int launch_stream(const CUtensorMap& ta, const CUtensorMap& tb, void* out, const EpiDev& ep, const StreamParams& sp, int smem,
                  int grid, cudaStream_t st) {
    static PerDevice attr_smem_dev;
    if (int& attr_smem = attr_smem_dev.cur(); smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(igemm_tc_stream_kernel<TCOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { ddpmir_set_error("igemm_tc_stream: smem opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
        attr_smem = smem;
    }
    igemm_tc_stream_kernel<TCOLS><<<grid, STREAM_THREADS, smem, st>>>(ta, tb, out, ep, sp);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

int g_num_sms = 0;
int g_tc_variant = 0;   // 0 = auto, 1 = tiled kernel only, 2 = streaming kernel whenever legal, 3 = tiled kernel with the widest N tile

}  // namespace

extern "C" int ddpmir_igemm_set_variant(int v) { g_tc_variant = v; return DDPMIR_OK; }

int ddpmir_igemm_tc(int taps, const void* x, int B, int H, int W, int Cin, const void* w, int N,
                    const ddpmir_epilogue_t* epi, void* out, cudaStream_t st) {
    // shape constraints of this kernel; anything else goes to the generic path
    if (Cin % BLOCK_K != 0 || N % 8 != 0) return DDPMIR_ERR_UNSUPPORTED;
    if (!pow2(H) || !pow2(W)) return DDPMIR_ERR_UNSUPPORTED;
    if (((uintptr_t)x & 15) || ((uintptr_t)w & 15)) return DDPMIR_ERR_UNSUPPORTED;
    EncodeTiledFn enc = get_encode();
    if (!enc) { ddpmir_set_error("igemm_tc: cuTensorMapEncodeTiled not available"); return DDPMIR_ERR_CUDA; }

    TcParams p;
    p.M = (long long)B * H * W; p.H = H; p.W = W; p.Cin = Cin; p.N = N; p.taps = taps;
    p.bw = W < BLOCK_M ? W : BLOCK_M;
    p.bh = (BLOCK_M / p.bw) < H ? (BLOCK_M / p.bw) : H;
    const int bb = BLOCK_M / (p.bw * p.bh);

    CUtensorMap ta, tb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)bb};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ddpmir_set_error("igemm_tc: activation tensor map failed (%d)", (int)r); return DDPMIR_ERR_CUDA; }
    }
    // persistent weight-resident kernel: whole [N, K] weight in shared memory, >= 2 tiles per SM
    const long long w_bytes = (long long)N * taps * Cin * 2;
    const int num_tiles = ceil_div(p.M, BLOCK_M);
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const bool stream_ok = N % 16 == 0 && N <= 256 && w_bytes <= 112 * 1024;
    if (g_tc_variant != 1 && g_tc_variant != 3 && stream_ok && (g_tc_variant == 2 || num_tiles >= 2 * g_num_sms)) {
        const cuuint64_t K = (cuuint64_t)taps * Cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)N};
        cuuint64_t strides[1] = {K * 2};
        cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)N};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ddpmir_set_error("igemm_tc: weight tensor map failed (%d)", (int)r); return DDPMIR_ERR_CUDA; }
        StreamParams sp;
        sp.t = p; sp.num_tiles = num_tiles; sp.nb = N;
        const int budget = 220 * 1024 - (int)w_bytes - 1024 - 256;
        sp.a_stages = budget / A_STAGE_BYTES;
        if (sp.a_stages > 8) sp.a_stages = 8;
        const int smem = (int)w_bytes + sp.a_stages * A_STAGE_BYTES + (2 * sp.a_stages + 5) * 8 + 16 + 1024;
        const int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
        EpiDev ep = make_epi(epi, H, W, N, DDPMIR_BF16);
        if (N <= 64) return launch_stream<128>(ta, tb, out, ep, sp, smem, grid, st);
        if (N <= 128) return launch_stream<256>(ta, tb, out, ep, sp, smem, grid, st);
        return launch_stream<512>(ta, tb, out, ep, sp, smem, grid, st);
    }
    // widest N tile that divides N: per k-block a CTA moves (128 + BLOCK_N) x 128 B for 128 x BLOCK_N x 64 MACs, and the
    // L2->SM path (~42 B/clk/SM) is what bounds the tensor pipe, so wider is better
    // Which divisor?  A CTA's k loop is bound by its L2->SM traffic, (128 + BLOCK_N) x 128 B per k-block, and the grid runs in
    // waves of one CTA per SM: cost ~ waves x (128 + BLOCK_N).  Large M fills the SMs at any width (widest wins, as before);
    // the 8x8 / 16x16 levels of a 16-image micro-batch have only 8..32 pixel tiles, so narrower tiles put 2-4x more SMs to work.
    int block_n = 64;
    {
        const int cand[3] = {256, 128, 64};
        long long best = -1;
        for (int i = 0; i < 3; ++i) {
            const int bn = cand[i];
            if (bn != 64 && N % bn != 0) continue;
            if (bn == 256 && taps * (Cin / BLOCK_K) <= 2) continue;
            const long long ctas = (long long)num_tiles * ceil_div(N, bn);
            const long long cost = ((ctas + g_num_sms - 1) / g_num_sms) * (128 + bn);
            if (g_tc_variant == 3) { block_n = bn; break; }        // tuning: the old rule (widest divisor)
            if (best < 0 || cost < best) { best = cost; block_n = bn; }
        }
    }
    {
        const cuuint64_t K = (cuuint64_t)taps * Cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)N};
        cuuint64_t strides[1] = {K * 2};
        cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)block_n};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ddpmir_set_error("igemm_tc: weight tensor map failed (%d)", (int)r); return DDPMIR_ERR_CUDA; }
    }
    EpiDev ep = make_epi(epi, H, W, N, DDPMIR_BF16);
    // short K loops (1x1 convolutions with K <= 128) are bandwidth-bound: a 2-stage ring keeps the CTA small so that
    // 3-4 CTAs share an SM and one CTA's epilogue overlaps the others' loads; long K loops get a 4-stage ring
    const int num_kb = taps * (Cin / BLOCK_K);
    if (num_kb <= 2)
        return block_n == 128 ? launch<128, 2>(ta, tb, out, ep, p, st) : launch<64, 2>(ta, tb, out, ep, p, st);
    if (block_n == 256) return launch<256, 4>(ta, tb, out, ep, p, st);
    return block_n == 128 ? launch<128, 4>(ta, tb, out, ep, p, st) : launch<64, 4>(ta, tb, out, ep, p, st);
}
