// svd_structure_preservation (0409_method.ipynb#c0:L321-346): rank-k truncation of every [H, W] plane.
// The reference calls torch.linalg.svd B*3 times from a Python loop; here one CTA owns one plane and runs a
// one-sided (Hestenes) Jacobi SVD on the ROWS of the plane, all planes concurrently:
//     A <- G X   with G orthogonal (accumulated rotations), rows of A mutually orthogonal  =>  |A_i| = sigma_i
//     X_k = G^T diag(keep) A ,  keep_i = 1 for the k rows of largest norm.
// Rotations of a round act on disjoint row pairs (round-robin tournament ordering), one warp per pair; the working
// set (A and G, (H*W + H*H) floats per plane) lives in global memory and stays L2-resident (50 MB for 192 planes of
// 256x256).  fp32 throughout; converges to |<a_i,a_j>| <= 1e-6 |a_i||a_j|, i.e. fp32 round-off.
#include "common.cuh"

namespace {

constexpr int SVD_THREADS = 1024;

__global__ void __launch_bounds__(SVD_THREADS)
svd_lowrank_kernel(const float* __restrict__ x, float* __restrict__ out, float* __restrict__ ws, int H, int W, int k,
                   int max_sweeps) {
    const int p = blockIdx.x;
    const float* X = x + (long long)p * H * W;
    float* O = out + (long long)p * H * W;
    float* A = ws + (long long)p * ((long long)H * W + (long long)H * H + 2 * H);
    float* G = A + (long long)H * W;
    float* nrm = G + (long long)H * H;
    int* keep = reinterpret_cast<int*>(nrm + H);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = SVD_THREADS / 32;

    for (int i = tid; i < H * W; i += SVD_THREADS) A[i] = X[i];
    for (int i = tid; i < H * H; i += SVD_THREADS) G[i] = (i / H == i % H) ? 1.f : 0.f;
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
                float* gi = G + (long long)i * H;
                float* gj = G + (long long)j * H;
                for (int c = lane; c < H; c += 32) {
                    const float u = gi[c], v = gj[c];
                    gi[c] = cs * u - sn * v;
                    gj[c] = sn * u + cs * v;
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
    for (int i = tid; i < H; i += SVD_THREADS) {
        const float s = nrm[i];
        int rank = 0;
        for (int j = 0; j < H; ++j) { const float q = nrm[j]; rank += (q > s || (q == s && j < i)) ? 1 : 0; }
        keep[i] = rank < k;
    }
    __syncthreads();

    // X_k[a][w] = sum_i keep_i * G[i][a] * A[i][w]
    const int ta = tid & 31, tw = tid >> 5;  // 32 rows x (32 * 4) columns per pass
    for (int a0 = 0; a0 < H; a0 += 32)
        for (int w0 = 0; w0 < W; w0 += 128) {
            const int a = a0 + ta, w = w0 + tw * 4;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if (a < H && w < W) {
                for (int i = 0; i < H; ++i) {
                    if (!keep[i]) continue;
                    const float gv = G[(long long)i * H + a];
                    const float* ar = A + (long long)i * W + w;
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (w + q < W) acc[q] = fmaf(gv, ar[q], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) if (w + q < W) O[(long long)a * W + w + q] = acc[q];
            }
        }
}

}  // namespace

extern "C" int ddpmir_svd_lowrank(const float* x, int planes, int H, int W, int k, float* out, float* ws, int sweeps,
                                  ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && out && ws, "svd_lowrank: null pointer");
    DDPMIR_CHECK_ARG(planes > 0 && H > 0 && W > 0 && k >= 1, "svd_lowrank: bad shape");
    DDPMIR_CHECK_ARG(H <= 1024 && W <= 4096, "svd_lowrank: plane too large (%d x %d)", H, W);
    if (sweeps <= 0) sweeps = 30;
    svd_lowrank_kernel<<<planes, SVD_THREADS, 0, (cudaStream_t)stream>>>(x, out, ws, H, W, k, sweeps);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
