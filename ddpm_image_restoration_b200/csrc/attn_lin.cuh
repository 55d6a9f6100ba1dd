// Shared by attn_lin.cu (pre-pass, SIMT kernels) and attn_lin_tc.cu (tcgen05 kernels): polynomial sets, workspace layout of the
// per-(image, head) parameters.
#pragma once
#include <cuda_runtime.h>

namespace attn_lin {

constexpr int MAXDEG = 6;
constexpr int NSETS = 7;
// minimax polynomials of 2^x in relative error on [-B, B] (tools/minimax_exp2.py): degree, window B, max relative error
//   0: 3, 0.75, 3.8e-4   1: 3, 1.0, 1.2e-3   2: 3, 1.25, 2.8e-3   3: 4, 1.5, 6.1e-4   4: 4, 2.0, 2.5e-3   5: 5, 2.5, 1.1e-3   6: 6, 3.5, 1.3e-3
// Which sets a call may use is a bit mask (set_mask below): head_dim 8 skips set 2 (its degree-4 map is nearly as cheap and four
// times more accurate), head_dim 16 stops at degree 4, and short sequences stop where the quadratic tier becomes cheaper.
__host__ __device__ constexpr int set_degree(int s) { return s <= 2 ? 3 : s <= 4 ? 4 : s == 5 ? 5 : 6; }
__host__ __device__ constexpr float set_bound(int s) {
    return s == 0 ? 0.75f : s == 1 ? 1.0f : s == 2 ? 1.25f : s == 3 ? 1.5f : s == 4 ? 2.0f : s == 5 ? 2.5f : 3.5f;
}
__device__ __forceinline__ float set_coef(int s, int n) {
    // (a select chain over literals: a table indexed with runtime s, n would live in local memory)
    const float c0[NSETS] = {0.999655739f, 0.998997116f, 0.997807596f, 0.999535858f, 0.997719925f, 1.00009398f, 1.00111371f};
    const float c1[NSETS] = {0.693718363f, 0.694930421f, 0.697431525f, 0.691511522f, 0.689048622f, 0.690660092f, 0.692568291f};
    const float c2[NSETS] = {0.245536242f, 0.249528671f, 0.254488064f, 0.241847765f, 0.24514987f, 0.238338385f, 0.23741251f};
    const float c3[NSETS] = {0.0547586426f, 0.0541850512f, 0.0534554379f, 0.0590221947f, 0.0614503155f, 0.0571867887f, 0.0548192962f};
    const float c4[NSETS] = {0.f, 0.f, 0.f, 0.00919362656f, 0.00887524488f, 0.0108539063f, 0.0103227623f};
    const float c5[NSETS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.00119714351f, 0.00158192741f};
    const float c6[NSETS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.000127974665f};
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < NSETS; ++i)
        if (s == i) r = n == 0 ? c0[i] : n == 1 ? c1[i] : n == 2 ? c2[i] : n == 3 ? c3[i] : n == 4 ? c4[i] : n == 5 ? c5[i] : c6[i];
    return r;
}
// Sets a call may use.  Cost model (measured on B200, per (image, head)): the polynomial tier is ~8.4e-13 s per feature and token
// at head_dim 16 (3.6e-13 at head_dim 8), the quadratic half-precision tier ~1.4e-13 s per score, so a map of F features pays
// off from L ~ 6 F (2.8 F) tokens on.
__host__ __device__ constexpr unsigned set_mask(int hd, int L, bool simt) {
    if (hd == 8) {
        unsigned m = 0x1Bu;                                   // sets 0, 1, 3, 4 (degree 3 and 4)
        if (!simt && L >= 4096) m |= 1u << 5;                 // degree 5: 1408 features
        if (!simt && L >= 16384) m |= 1u << 6;                // degree 6: 3200 features
        return m;
    }
    unsigned m = 0x07u;                                       // head_dim 16: sets 0, 1, 2 (degree 3: 1024 features)
    if (!simt && L >= 32768) m |= 0x18u;                      // degree 4: 5504 features
    return m;
}

// params per (image, head): a[16] | b[16] | D[16] | 1/D[16] | e_off (offset of the key weights' exponent) | pad
constexpr int PSTRIDE = 72;
constexpr int P_A = 0, P_B = 16, P_D = 32, P_DI = 48, P_EOFF = 64;

__host__ __device__ constexpr long long binom(int n, int k) {
    if (k < 0 || k > n) return 0;
    long long r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}
// monomials of degree <= deg in n variables
__host__ __device__ constexpr int nfeat(int n, int deg) { return deg < 0 ? 0 : (int)binom(n + deg, deg); }

}  // namespace attn_lin

// attn_lin_tc.cu
size_t ddpmir_attention_lin_tc_workspace(int B, int L, int hd, int heads);
// counts [MAXDEG + 1], lists [MAXDEG + 1][B*heads]: the (image, head) pairs of every degree, compacted by the pre-pass
int ddpmir_attention_lin_tc(const void* qkv, void* out, const int* tier, const int* counts, const int* lists, const float* params, void* ws,
                            int* flags, int B, int L, int C, int heads, int max_degree, cudaStream_t st);
