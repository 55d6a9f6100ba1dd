// svd_structure_preservation (0409_method.ipynb#c0:L321-346): rank-k truncation of every [H, W] plane.
// The reference calls torch.linalg.svd B*3 times from a Python loop; here one CTA owns one plane and runs a
// one-sided (Hestenes) Jacobi SVD on the ROWS of the plane, all planes concurrently:
//     A <- G X   with G orthogonal (the product of the rotations), rows of A mutually orthogonal  =>  A_i = sigma_i v_i^T
// G is never formed (round 2; it doubled the rotation work and the working set): with the k rows of largest norm kept,
//     X_k = U_k S_k V_k^T = X V_k V_k^T = (X A_k^T diag(1 / sigma^2)) A_k ,
// two small batched GEMMs on the converged A (T = X A_s^T, X_k = T A).  Rows whose sigma^2 is below fp32 noise carry nothing
// and are skipped (they would divide by ~0).  Rotations of a round act on disjoint rows (round-robin tournament ordering over
// blocks of four rows; a warp holds the eight rows of a block pair in registers for its 16 rotations -- see
// jacobi_sweeps_regs; planes wider than 256 columns take the plain row-pair loop); the working set (A, H*W floats per plane)
// lives in global memory and stays L2-resident (50 MB for 192 planes of 256x256).  256 threads per CTA, two planes per SM:
// 192 planes run in ONE wave on 148 SMs.  fp32 throughout; converges to |<a_i,a_j>| <= 1e-6 |a_i||a_j|, i.e. fp32 round-off.
#include "common.cuh"

namespace {

constexpr int SVD_THREADS = 256;   // 8 warps per plane, two planes per SM, up to 128 registers per thread (eight rows of a block pair)
constexpr int EPL_MAX = 8;      // row elements per lane that the register-resident path holds: planes up to 256 columns

// round-robin tournament: pair pi of round r among n players (n even); every pair of rows meets once per sweep
__device__ __forceinline__ void tournament_pair(int pi, int r, int n, int& i, int& j) {
    if (pi == 0) { i = n - 1; j = r; }
    else { i = (r + pi) % (n - 1); j = (r - pi + (n - 1)) % (n - 1); }
}

// Jacobi rotation that makes rows with <a_i,a_i> = alpha, <a_j,a_j> = beta, <a_i,a_j> = gamma orthogonal; false = leave them
__device__ __forceinline__ bool jacobi_rotation(float alpha, float beta, float gamma, float& cs, float& sn) {
    if (alpha < 1e-30f || beta < 1e-30f) return false;
    if (fabsf(gamma) <= 1e-6f * sqrtf(alpha * beta)) return false;
    const float zeta = (beta - alpha) / (2.f * gamma);
    const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
    cs = rsqrtf(1.f + t * t); sn = cs * t;
    return true;
}

// All sweeps for planes of at most 32 * EPL_MAX columns: BLOCK Jacobi.  The kernel is bound by moving rows between L2 and the
// SM (every rotation of the plain tournament reads and writes two whole rows: the plane crosses the L2 <-> SM path once per
// round, ~60 B/clk/SM), so rows are grouped in blocks of RB = 4 and the tournament runs over BLOCKS: a warp loads the eight
// rows of a block pair into registers (lane = column mod 32), performs all 16 cross rotations there -- four "diagonals" of
// four independent rotations, so the shuffle reductions of one hide behind the others -- and writes back only the rows that
// changed.  Sixteen rotations per sixteen row transfers instead of four: half the traffic per sweep.  The six pairs inside each
// block are rotated in the first round of a sweep (every block takes part in every round exactly once), so a sweep still
// visits every pair of rows once.
constexpr int RB = 4;

__device__ __forceinline__ void rotate_rows(float (&x)[EPL_MAX], float (&y)[EPL_MAX], bool& cx, bool& cy, int& rotated) {
    float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
    for (int e = 0; e < EPL_MAX; ++e) { al = fmaf(x[e], x[e], al); be = fmaf(y[e], y[e], be); ga = fmaf(x[e], y[e], ga); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        al += __shfl_xor_sync(0xffffffffu, al, o);
        be += __shfl_xor_sync(0xffffffffu, be, o);
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
    }
    float cs, sn;
    if (!jacobi_rotation(al, be, ga, cs, sn)) return;
    rotated = 1; cx = true; cy = true;
#pragma unroll
    for (int e = 0; e < EPL_MAX; ++e) {
        const float u = x[e], v = y[e];
        x[e] = cs * u - sn * v;
        y[e] = sn * u + cs * v;
    }
}

__device__ __forceinline__ void jacobi_sweeps_regs(float* __restrict__ A, int H, int W, int max_sweeps, int warp, int lane, int nwarps) {
    const int nblk = (H + RB - 1) / RB;
    const int n = (nblk + 1) & ~1;            // even number of players (a dummy block if needed)
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        int rotated = 0;
        for (int r = 0; r < (n > 1 ? n - 1 : 1); ++r) {
            for (int pi = warp; pi < n / 2; pi += nwarps) {
                int bi, bj;
                tournament_pair(pi, r, n, bi, bj);
                float a[2 * RB][EPL_MAX];
                bool ch[2 * RB];
                int row[2 * RB];
#pragma unroll
                for (int q = 0; q < 2 * RB; ++q) {
                    const int blk = q < RB ? bi : bj;
                    row[q] = blk < nblk ? blk * RB + (q % RB) : H;      // H = "no such row"
                    ch[q] = false;
#pragma unroll
                    for (int e = 0; e < EPL_MAX; ++e) {
                        const int c = lane + 32 * e;
                        a[q][e] = (row[q] < H && c < W) ? A[(long long)row[q] * W + c] : 0.f;
                    }
                }
                if (r == 0) {     // pairs inside each block, once per sweep: (0,1)(2,3) | (0,2)(1,3) | (0,3)(1,2)
#pragma unroll
                    for (int g = 0; g < 2 * RB; g += RB) {
                        rotate_rows(a[g + 0], a[g + 1], ch[g + 0], ch[g + 1], rotated); rotate_rows(a[g + 2], a[g + 3], ch[g + 2], ch[g + 3], rotated);
                        rotate_rows(a[g + 0], a[g + 2], ch[g + 0], ch[g + 2], rotated); rotate_rows(a[g + 1], a[g + 3], ch[g + 1], ch[g + 3], rotated);
                        rotate_rows(a[g + 0], a[g + 3], ch[g + 0], ch[g + 3], rotated); rotate_rows(a[g + 1], a[g + 2], ch[g + 1], ch[g + 2], rotated);
                    }
                }
#pragma unroll
                for (int d = 0; d < RB; ++d)          // diagonal d: rows (x, RB + (x + d) % RB), independent of each other
#pragma unroll
                    for (int x = 0; x < RB; ++x) {
                        const int y = RB + (x + d) % RB;
                        rotate_rows(a[x], a[y], ch[x], ch[y], rotated);
                    }
#pragma unroll
                for (int q = 0; q < 2 * RB; ++q)
                    if (ch[q] && row[q] < H) {
#pragma unroll
                        for (int e = 0; e < EPL_MAX; ++e) {
                            const int c = lane + 32 * e;
                            if (c < W) A[(long long)row[q] * W + c] = a[q][e];
                        }
                    }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(rotated)) break;
    }
}

__global__ void __launch_bounds__(SVD_THREADS, 2)
svd_jacobi_kernel(const float* __restrict__ x, float* __restrict__ ws, int H, int W, int k, int max_sweeps) {
    const int p = blockIdx.x;
    const float* X = x + (long long)p * H * W;
    float* A = ws + (long long)p * ((long long)H * W + (long long)H * H + 2 * H);
    float* scale = A + (long long)H * W + (long long)H * H;      // [H]: 1 / sigma_i^2 for the kept rows, 0 otherwise
    float* nrm = scale + H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SVD_THREADS / 32;

    for (int i = tid; i < H * W; i += SVD_THREADS) A[i] = X[i];
    __syncthreads();

    const int n = (H + 1) & ~1;  // even number of players (a dummy row if H is odd)
    if (W <= 32 * EPL_MAX) jacobi_sweeps_regs(A, H, W, max_sweeps, warp, lane, nwarps);
    else
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        int rotated = 0;
        for (int r = 0; r < n - 1; ++r) {
            for (int pi = warp; pi < n / 2; pi += nwarps) {
                int i, j;
                tournament_pair(pi, r, n, i, j);
                if (i >= H || j >= H) continue;
                float* ai = A + (long long)i * W;
                float* aj = A + (long long)j * W;
                float alpha = 0.f, beta = 0.f, gamma = 0.f;
                for (int c = lane; c < W; c += 32) {
                    const float u = ai[c], v = aj[c];
                    alpha = fmaf(u, u, alpha); beta = fmaf(v, v, beta); gamma = fmaf(u, v, gamma);
                }
                alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
                float cs, sn;
                if (!jacobi_rotation(alpha, beta, gamma, cs, sn)) continue;
                rotated = 1;
                for (int c = lane; c < W; c += 32) {
                    const float u = ai[c], v = aj[c];
                    ai[c] = cs * u - sn * v;
                    aj[c] = sn * u + cs * v;
                }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(rotated)) break;
    }

    // singular values = row norms; rank them (ties by index) and keep the k largest
    for (int i = warp; i < H; i += nwarps) {
        float s = 0.f;
        for (int c = lane; c < W; c += 32) { const float u = A[(long long)i * W + c]; s = fmaf(u, u, s); }
        s = warp_sum(s);
        if (lane == 0) nrm[i] = s;
    }
    __syncthreads();
    float top = 0.f;
    for (int j = 0; j < H; ++j) top = fmaxf(top, nrm[j]);
    for (int i = tid; i < H; i += SVD_THREADS) {
        const float s = nrm[i];
        int rank = 0;
        for (int j = 0; j < H; ++j) { const float q = nrm[j]; rank += (q > s || (q == s && j < i)) ? 1 : 0; }
        // sigma_i^2 below 1e-12 sigma_max^2: the component is at fp32 round-off of the plane and contributes nothing
        scale[i] = (rank < k && s > 1e-12f * top && s > 0.f) ? 1.f / s : 0.f;
    }
}

// Batched 64x64-tile SIMT GEMMs of the reconstruction (2 x 2 H H W FLOP per plane -- 13 GFLOP for 192 planes of 256x256).
//   NT: T[h][i] = scale[i] * sum_w X[h][w] A[i][w]        NN: O[h][w] = sum_i T[h][i] A[i][w]
template <bool NT>
__global__ void __launch_bounds__(256)
svd_gemm_kernel(const float* __restrict__ x, const float* __restrict__ ws, float* __restrict__ out_or_null, int H, int W) {
    const int p = blockIdx.z;
    const long long stride = (long long)H * W + (long long)H * H + 2 * H;
    const float* A = ws + (long long)p * stride;
    float* T = const_cast<float*>(A) + (long long)H * W;
    const float* scale = T + (long long)H * H;
    // C [M x N] = L [M x K] * R,  NT: L = X (H x W), R[k][n] = A[n][k] (N = H rows of A), K = W;  NN: L = T (H x H), R = A (H x W), K = H
    const float* Lm = NT ? x + (long long)p * H * W : T;
    const int M = H, N = NT ? H : W, K = NT ? W : H;
    float* Cm = NT ? T : out_or_null + (long long)p * H * W;
    __shared__ float ls[16][64 + 1], rs[16][64 + 1];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int kk = i & 15, mm = i >> 4;
            ls[kk][mm] = (m0 + mm < M && k0 + kk < K) ? Lm[(long long)(m0 + mm) * K + k0 + kk] : 0.f;
            if (NT) rs[kk][mm] = (n0 + mm < N && k0 + kk < K) ? A[(long long)(n0 + mm) * W + k0 + kk] : 0.f;
        }
        if (!NT)
            for (int i = threadIdx.x; i < 64 * 16; i += 256) {
                const int nn = i & 63, kk = i >> 6;
                rs[kk][nn] = (n0 + nn < N && k0 + kk < K) ? A[(long long)(k0 + kk) * W + n0 + nn] : 0.f;
            }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = ls[kk][ty * 4 + i]; b[i] = rs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, nn = n0 + tx * 4 + j;
            if (m < M && nn < N) Cm[(long long)m * N + nn] = NT ? acc[i][j] * scale[nn] : acc[i][j];
        }
}

}  // namespace

extern "C" int ddpmir_svd_lowrank(const float* x, int planes, int H, int W, int k, float* out, float* ws, int sweeps,
                                  ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && ws, "svd_lowrank: null pointer");
    DDPMIR_CHECK_ARG(planes > 0 && H > 0 && W > 0 && k >= 1, "svd_lowrank: bad shape");
    DDPMIR_CHECK_ARG(H <= 1024 && W <= 4096, "svd_lowrank: plane too large (%d x %d)", H, W);
    DDPMIR_CHECK_ARG(planes <= 65535, "svd_lowrank: too many planes");
    if (sweeps <= 0) sweeps = 30;
    cudaStream_t st = (cudaStream_t)stream;
    svd_jacobi_kernel<<<planes, SVD_THREADS, 0, st>>>(x, ws, H, W, k, sweeps);
    DDPMIR_LAUNCH_CHECK();
    svd_gemm_kernel<true><<<dim3(ceil_div(H, 64), ceil_div(H, 64), planes), 256, 0, st>>>(x, ws, nullptr, H, W);
    DDPMIR_LAUNCH_CHECK();
    svd_gemm_kernel<false><<<dim3(ceil_div(W, 64), ceil_div(H, 64), planes), 256, 0, st>>>(x, ws, out, H, W);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
