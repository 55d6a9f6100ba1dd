// Tensor-core weight gradient for the training step:  dW[n, k] += sum_m dY[m, n] * im2col(X)[m, k]  on bf16 operands
// (mma.sync m16n8k16, fp32 accumulate).  The reduction runs over pixels m, which is the contiguous-row dimension of both
// operand tensors ([m][n] and [m][c]), so both fragments come from ldmatrix.trans on [pixel][channel] shared-memory tiles
// filled with cp.async (zero-fill for the 3x3 halo and the tile edges).  One CTA owns a 128 (n) x 64 (k) block of dW for
// a slice of the pixels; slices are combined with fp32 atomics directly in the checkpoint layout.
#include "common.cuh"

namespace {

constexpr int BN = 128, BKO = 64, PM = 32;   // dW tile and pixels per pipeline stage

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256)
wgrad_mma_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, float* __restrict__ out, long long M, int H, int W, int Cin,
                 int N, int taps, int n_begin, int n_count, int k_begin, int k_count, int out_ld, int layout, long long m_per_split) {
    // [stage][pixel][channel] with the 16-byte chunk index XOR-swizzled by (pixel & 7)
    __shared__ __align__(128) bf16 Ys[2][PM][BN];
    __shared__ __align__(128) bf16 Xs[2][PM][BKO];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = n_begin + blockIdx.y * BN, k0 = k_begin + blockIdx.x * BKO;
    const int n_end = n_begin + n_count, k_end = k_begin + k_count;
    const int tap = taps == 9 ? k0 / Cin : 0;
    const int c0 = k0 - tap * Cin;                    // first input channel of this k tile (tile lies inside one tap)
    const int dh = taps == 9 ? tap / 3 - 1 : 0, dw = taps == 9 ? tap % 3 - 1 : 0;
    const long long m_lo = (long long)blockIdx.z * m_per_split;
    const long long m_hi = min(M, m_lo + m_per_split);
    const int hw = H * W;

    auto load_stage = [&](long long mb, int stage) {
        // dY tile: PM rows x 16 chunks
        for (int i = tid; i < PM * (BN / 8); i += 256) {
            const int r = i / (BN / 8), ch = i % (BN / 8);
            const long long m = mb + r;
            const int n = n0 + ch * 8;
            const bool ok = m < m_hi && n < n_end;        // N % 8 == 0 and n_end % 8 == 0 are checked by the host
            cp_async16_zfill(&Ys[stage][r][(ch ^ (r & 7)) * 8], dy + (ok ? m * N + n : 0), ok);
        }
        // X tile (shifted pixels of this tap): PM rows x 8 chunks
        {
            const int r = tid / (BKO / 8), ch = tid % (BKO / 8);
            const long long m = mb + r;
            bool ok = m < m_hi && (c0 + ch * 8) < Cin && (k0 + ch * 8) < k_end;
            long long src = 0;
            if (ok) {
                const int b = (int)(m / hw);
                const int rem = (int)(m - (long long)b * hw);
                const int h = rem / W + dh, w = rem % W + dw;
                ok = h >= 0 && h < H && w >= 0 && w < W;
                src = (((long long)b * H + h) * W + w) * Cin + c0 + ch * 8;
            }
            cp_async16_zfill(&Xs[stage][r][(ch ^ (r & 7)) * 8], x + (ok ? src : 0), ok);
        }
    };

    const int wn = warp & 3, wk = warp >> 2;          // warp tile: 32 n x 32 k
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;

    const long long nchunks = (m_hi - m_lo + PM - 1) / PM;
    if (nchunks > 0) { load_stage(m_lo, 0); }
    cp_async_commit();
    const uint32_t ys = (uint32_t)__cvta_generic_to_shared(&Ys[0][0][0]), xs = (uint32_t)__cvta_generic_to_shared(&Xs[0][0][0]);
    const int lr = lane & 7, lj = lane >> 3;
    for (long long c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) { load_stage(m_lo + (c + 1) * PM, (int)((c + 1) & 1)); cp_async_commit(); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const int st = (int)(c & 1);
#pragma unroll
        for (int s = 0; s < PM / 16; ++s) {
            uint32_t af[2][4], bfr[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                // matrices: (m0, n), (m0, n+8), (m0+8, n), (m0+8, n+8) -> a0, a1, a2, a3
                const int row = s * 16 + (lj >> 1) * 8 + lr;
                const int ch = (wn * 32 + i * 16 + (lj & 1) * 8) / 8;
                ldsm_x4_t(af[i], ys + ((st * PM + row) * BN + ((ch ^ (row & 7)) * 8)) * 2);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                // matrices: (m0, k), (m0+8, k), (m0, k+8), (m0+8, k+8) -> (b0, b1) of k-tile 2j, (b0, b1) of k-tile 2j+1
                const int row = s * 16 + (lj & 1) * 8 + lr;
                const int ch = (wk * 32 + j * 16 + (lj >> 1) * 8) / 8;
                ldsm_x4_t(bfr[j], xs + ((st * PM + row) * BKO + ((ch ^ (row & 7)) * 8)) * 2);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma_16816(acc[i][2 * j], af[i], bfr[j][0], bfr[j][1]);
                    mma_16816(acc[i][2 * j + 1], af[i], bfr[j][2], bfr[j][3]);
                }
        }
        __syncthreads();
    }
    // accumulator (n rows, k cols) -> atomics into the gradient
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = n0 + wn * 32 + i * 16 + (lane >> 2) + (q >> 1) * 8;
                const int k = k0 + wk * 32 + j * 8 + (lane & 3) * 2 + (q & 1);
                if (n < n_end && k < k_end && (k - tap * Cin) < Cin) {
                    long long idx;
                    if (layout == 1) idx = ((long long)(n - n_begin) * Cin + (k - tap * Cin)) * 9 + tap;
                    else idx = (long long)(n - n_begin) * out_ld + (k - k_begin);
                    atomicAdd(&out[idx], acc[i][j][q]);
                }
            }
}

}  // namespace

// returns DDPMIR_ERR_UNSUPPORTED when the shape needs the generic kernel
int ddpmir_wgrad_mma(const void* dy, const void* x, float* out, int B, int H, int W, int Cin, int N, int taps, int n_begin, int n_count,
                     int k_begin, int k_count, int out_ld, int oihw, cudaStream_t st) {
    if (N % 8 != 0 || Cin % 8 != 0 || n_begin % 8 != 0 || n_count % 8 != 0 || k_begin % 8 != 0 || k_count % 8 != 0) return DDPMIR_ERR_UNSUPPORTED;
    if (taps == 9 && (Cin % BKO != 0 || k_begin % BKO != 0)) return DDPMIR_ERR_UNSUPPORTED;   // a k tile must stay inside one tap
    if (((uintptr_t)dy & 15) || ((uintptr_t)x & 15)) return DDPMIR_ERR_UNSUPPORTED;
    const long long M = (long long)B * H * W;
    const int tiles = ceil_div(k_count, BKO) * ceil_div(n_count, BN);
    int splits = (148 * 3 + tiles - 1) / tiles;
    const long long max_splits = (M + 4 * PM - 1) / (4 * PM);
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    long long per = (M + splits - 1) / splits;
    per = (per + PM - 1) / PM * PM;
    splits = (int)((M + per - 1) / per);
    dim3 grid(ceil_div(k_count, BKO), ceil_div(n_count, BN), splits);
    wgrad_mma_kernel<<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, out, M, H, W, Cin, N, taps, n_begin, n_count, k_begin, k_count,
                                           out_ld, oihw, per);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
