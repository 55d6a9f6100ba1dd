// svd_structure_preservation (0409_method.ipynb#c0:L321-346): rank-k truncation of every [H, W] plane.
// The reference calls torch.linalg.svd B*3 times from a Python loop; here one CTA owns one plane and runs a
// one-sided (Hestenes) Jacobi SVD on the ROWS of the plane, all planes concurrently:
//     A <- G X   with G orthogonal (the product of the rotations), rows of A mutually orthogonal  =>  A_i = sigma_i v_i^T
// G is never formed (round 2; it doubled the rotation work and the working set): with the k rows of largest norm kept,
//     X_k = U_k S_k V_k^T = X V_k V_k^T = (X A_k^T diag(1 / sigma^2)) A_k ,
// two small batched GEMMs on the converged A (T = X A_s^T, X_k = T A).  Rows whose sigma^2 is below fp32 noise carry nothing
// and are skipped (they would divide by ~0).  Rotations of a round act on disjoint row pairs (round-robin tournament
// ordering), one warp per pair; the working set (A, H*W floats per plane) lives in global memory and stays L2-resident
// (50 MB for 192 planes of 256x256).  512 threads per CTA so that two planes share an SM: 192 planes run in ONE wave on
// 148 SMs instead of 1.3.  fp32 throughout; converges to |<a_i,a_j>| <= 1e-6 |a_i||a_j|, i.e. fp32 round-off.
#include "common.cuh"

namespace {

constexpr int SVD_THREADS = 512;

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
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        int rotated = 0;
        for (int r = 0; r < n - 1; ++r) {
            for (int pi = warp; pi < n / 2; pi += nwarps) {
                int i, j;
                if (pi == 0) { i = n - 1; j = r; }
                else { i = (r + pi) % (n - 1); j = (r - pi + (n - 1)) % (n - 1); }
                if (i >= H || j >= H) continue;
                float* ai = A + (long long)i * W;
                float* aj = A + (long long)j * W;
                float alpha = 0.f, beta = 0.f, gamma = 0.f;
                for (int c = lane; c < W; c += 32) {
                    const float u = ai[c], v = aj[c];
                    alpha = fmaf(u, u, alpha); beta = fmaf(v, v, beta); gamma = fmaf(u, v, gamma);
                }
                alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
                if (alpha < 1e-30f || beta < 1e-30f) continue;
                if (fabsf(gamma) <= 1e-6f * sqrtf(alpha * beta)) continue;
                rotated = 1;
                const float zeta = (beta - alpha) / (2.f * gamma);
                const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
                const float cs = rsqrtf(1.f + t * t), sn = cs * t;
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
