// tcgen05 / TMA implicit-GEMM convolution (placeholder until the kernel lands: reports "unsupported" so the
// router falls back to the generic kernel).
#include "epilogue.cuh"
int ddpmir_igemm_tc(int taps, const void* x, int B, int H, int W, int Cin, const void* w, int N,
                    const ddpmir_epilogue_t* epi, void* out, cudaStream_t st) {
    return DDPMIR_ERR_UNSUPPORTED;
}
