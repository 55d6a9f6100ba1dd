// C-ABI entry points for the GEMM-shaped ops; routes to the tcgen05 kernels (conv_tc.cu) or the generic SIMT
// kernel (gemm_simt.cu).
#include "epilogue.cuh"

int ddpmir_conv3x3_simt(const void*, int, int, int, int, int, const void*, int, const ddpmir_epilogue_t*, void*, cudaStream_t);
int ddpmir_gemm_simt(const void*, int, int, int, int, int, const void*, int, const ddpmir_epilogue_t*, void*, cudaStream_t);
// return DDPMIR_ERR_UNSUPPORTED when the shape does not fit the tensor-core kernel
int ddpmir_igemm_tc(int taps, const void* x, int B, int H, int W, int Cin, const void* w, int N,
                    const ddpmir_epilogue_t* epi, void* out, cudaStream_t st);

static int route(int taps, const void* x, int dtype, int B, int H, int W, int Cin, const void* w, int N,
                 const ddpmir_epilogue_t* epi, void* out, int impl, cudaStream_t st) {
    DDPMIR_CHECK_ARG(x && w && out, "igemm: null pointer");
    DDPMIR_CHECK_ARG(dtype == DDPMIR_F32 || dtype == DDPMIR_BF16, "igemm: bad dtype %d", dtype);
    DDPMIR_CHECK_ARG(B > 0 && H > 0 && W > 0 && N > 0, "igemm: bad shape");
    DDPMIR_CHECK_ARG(Cin > 0 && Cin % 16 == 0, "igemm: Cin/K must be a multiple of 16 (got %d)", Cin);
    int rc = check_epi(epi, N);
    if (rc) return rc;
    if (impl == DDPMIR_IMPL_TENSOR || impl == DDPMIR_IMPL_AUTO) {
        if (dtype == DDPMIR_BF16) {
            rc = ddpmir_igemm_tc(taps, x, B, H, W, Cin, w, N, epi, out, st);
            if (rc != DDPMIR_ERR_UNSUPPORTED) return rc;
        }
        if (impl == DDPMIR_IMPL_TENSOR) {
            ddpmir_set_error("igemm: tensor-core kernel does not support this shape/dtype");
            return DDPMIR_ERR_UNSUPPORTED;
        }
    }
    return taps == 9 ? ddpmir_conv3x3_simt(x, dtype, B, H, W, Cin, w, N, epi, out, st)
                     : ddpmir_gemm_simt(x, dtype, B, H, W, Cin, w, N, epi, out, st);
}

extern "C" int ddpmir_conv3x3(const void* x, int dtype, int B, int H, int W, int Cin, const void* w, int N,
                              const ddpmir_epilogue_t* epi, void* out, int impl, ddpmir_stream_t stream) {
    return route(9, x, dtype, B, H, W, Cin, w, N, epi, out, impl, (cudaStream_t)stream);
}
extern "C" int ddpmir_gemm(const void* a, int dtype, int B, int H, int W, int K, const void* w, int N,
                           const ddpmir_epilogue_t* epi, void* out, int impl, ddpmir_stream_t stream) {
    return route(1, a, dtype, B, H, W, K, w, N, epi, out, impl, (cudaStream_t)stream);
}
