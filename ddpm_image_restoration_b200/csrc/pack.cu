// Per-step weight packing of the training step (webp_training.py:476-537: the weights change every optimizer step, so the GEMM
// operand copies do too).  One pass over a checkpoint-layout weight [N, Cin, kh, kw] (fp32, taps = kh*kw = 9 or 1) writes BOTH
// operand forms the step needs, in the operand dtype:
//     fwd[n, tap*Cin + c]          = w[n, c, tap]        forward / weight-gradient layout  [N, (kh, kw, cin)]
//     bwd[c, (taps-1-tap)*N + n]   = w[n, c, tap]        data-gradient layout: transposed, taps flipped  [Cin, (kh, kw, cout)]
// Leading dimensions and column offsets let several checkpoint tensors land in one stacked operand (the low/high gate MLP).
// CTA = 32 output channels x 32 input channels x taps through shared memory: reads are contiguous runs of 32*taps floats,
// writes are 32-element runs in both layouts (the torch version was permute + flip + cat + cast: five passes per weight).
#include "epilogue.cuh"

namespace {

template <int TAPS>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, int N, int Cin, void* __restrict__ fwd, long long fwd_ld, long long fwd_off,
                   void* __restrict__ bwd, long long bwd_ld, long long bwd_off, int dtype) {
    constexpr int RUN = 32 * TAPS;
    __shared__ float T[32][RUN + 1];
    const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.x; i < 32 * RUN; i += 256) {
        const int nl = i / RUN, j = i - nl * RUN;
        const int n = n0 + nl, c = c0 + j / TAPS;
        T[nl][j] = (n < N && c < Cin) ? w[((long long)n * Cin + c0) * TAPS + j] : 0.f;
    }
    __syncthreads();
    if (fwd)
        for (int i = threadIdx.x; i < 32 * RUN; i += 256) {
            const int nl = i / RUN, j = i - nl * RUN;
            const int tap = j >> 5, cl = j & 31;
            const int n = n0 + nl, c = c0 + cl;
            if (n < N && c < Cin) st_any(fwd, dtype, (long long)n * fwd_ld + fwd_off + (long long)tap * Cin + c, T[nl][cl * TAPS + tap]);
        }
    if (bwd)
        for (int i = threadIdx.x; i < 32 * RUN; i += 256) {
            const int cl = i / RUN, j = i - cl * RUN;
            const int tapo = j >> 5, nl = j & 31;
            const int n = n0 + nl, c = c0 + cl;
            if (n < N && c < Cin)
                st_any(bwd, dtype, (long long)c * bwd_ld + bwd_off + (long long)tapo * N + n, T[nl][cl * TAPS + (TAPS - 1 - tapo)]);
        }
}

}  // namespace

extern "C" int ddpmir_pack_weight(const float* w, int N, int Cin, int taps, void* fwd, long long fwd_ld, long long fwd_off, void* bwd,
                                  long long bwd_ld, long long bwd_off, int dtype, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(w && (fwd || bwd) && N > 0 && Cin > 0, "pack_weight: bad arguments");
    DDPMIR_CHECK_ARG(taps == 1 || taps == 9, "pack_weight: taps %d", taps);
    DDPMIR_CHECK_ARG(dtype == DDPMIR_F32 || dtype == DDPMIR_BF16, "pack_weight: dtype %d", dtype);
    DDPMIR_CHECK_ARG(fwd_ld >= 0 && bwd_ld >= 0 && fwd_off >= 0 && bwd_off >= 0, "pack_weight: bad layout");
    dim3 grid(ceil_div(Cin, 32), ceil_div(N, 32));
    DDPMIR_CHECK_ARG(grid.y <= 65535, "pack_weight: N too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (taps == 9) pack_weight_kernel<9><<<grid, 256, 0, st>>>(w, N, Cin, fwd, fwd_ld, fwd_off, bwd, bwd_ld, bwd_off, dtype);
    else pack_weight_kernel<1><<<grid, 256, 0, st>>>(w, N, Cin, fwd, fwd_ld, fwd_off, bwd, bwd_ld, bwd_off, dtype);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
