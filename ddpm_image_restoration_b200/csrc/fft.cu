// phase_consistency (webp_inference.py:531-550) with hand-written shared-memory FFTs (no cuFFT):
//     out = alpha*x + (1-alpha) * Re ifft2( |fft2(x)| * exp(i*angle(fft2(ref))) )
// 2-D transforms are row passes and column passes of a batched radix-2 Stockham FFT held in shared memory.
// The reference recomputes angle(fft2(ref)) at every call although ref (= y) never changes during a trajectory;
// here its unit phasors are computed once (ddpmir_phase_reference) and the per-call work is
//     row FFT(x) -> [column FFT -> magnitude * phasor -> inverse column FFT] -> inverse row FFT + blend,
// three kernels, with the element-wise complex arithmetic fused into the column pass.
#include "common.cuh"

namespace {

constexpr int FFT_THREADS = 256;
constexpr int FFT_ELEMS = 4096;  // complex elements per shared-memory buffer (two buffers = 64 KB)

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// nb sequences of length N, contiguous in `a`; result (natural order) is returned in the pointer `a` after swaps.
__device__ __forceinline__ void fft_batch(float2*& a, float2*& b, int nb, int N, float sign) {
    const int half = N >> 1;
    for (int Ns = 1; Ns < N; Ns <<= 1) {
        const float inv = sign / (float)Ns;
        for (int idx = threadIdx.x; idx < nb * half; idx += FFT_THREADS) {
            const int q = idx / half, j = idx - q * half;
            const int k = j & (Ns - 1);
            float s, c;
            sincospif((float)k * inv, &s, &c);
            const float2 u0 = a[q * N + j], u1 = a[q * N + j + half];
            const float2 t = make_float2(u1.x * c - u1.y * s, u1.x * s + u1.y * c);
            const int j0 = ((j - k) << 1) + k;
            b[q * N + j0] = cadd(u0, t);
            b[q * N + j0 + Ns] = csub(u0, t);
        }
        __syncthreads();
        float2* tmp = a; a = b; b = tmp;
    }
}

// forward row FFT of real rows: x [rows_total, W] -> ws [rows_total, W] complex
__global__ void __launch_bounds__(FFT_THREADS)
row_fft_kernel(const float* __restrict__ x, float2* __restrict__ ws, long long rows_total, int W, int rows_per_cta) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const int nb = (int)min((long long)rows_per_cta, rows_total - r0);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) a[i] = make_float2(x[r0 * W + i], 0.f);
    __syncthreads();
    fft_batch(a, b, nb, W, -1.f);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) ws[r0 * W + i] = a[i];
}

// column pass.  MODE 0: phasor = unit(colFFT(ws));  MODE 1: ws = colIFFT(|colFFT(ws)| * phasor)
template <int MODE>
__global__ void __launch_bounds__(FFT_THREADS)
col_pass_kernel(float2* __restrict__ ws, float2* __restrict__ phasor, int H, int W, int cols_per_cta) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const int p = blockIdx.y;
    const int c0 = blockIdx.x * cols_per_cta;
    const int nb = min(cols_per_cta, W - c0);
    float2* plane = ws + (long long)p * H * W;
    float2* pplane = phasor + (long long)p * H * W;
    // gather: consecutive threads read consecutive columns of one row (nb*8 contiguous bytes)
    for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
        const int h = i / nb, c = i - h * nb;
        a[c * H + h] = plane[(long long)h * W + c0 + c];
    }
    __syncthreads();
    fft_batch(a, b, nb, H, -1.f);
    if (MODE == 0) {
        for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
            const int h = i / nb, c = i - h * nb;
            const float2 v = a[c * H + h];
            const float m = sqrtf(v.x * v.x + v.y * v.y);
            // cos/sin(angle(v)); torch.angle(0) = 0 -> (1, 0)
            pplane[(long long)h * W + c0 + c] = m > 0.f ? make_float2(v.x / m, v.y / m) : make_float2(1.f, 0.f);
        }
    } else {
        for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
            const int h = i / nb, c = i - h * nb;
            const float2 v = a[c * H + h];
            const float m = sqrtf(v.x * v.x + v.y * v.y);
            const float2 ph = pplane[(long long)h * W + c0 + c];
            a[c * H + h] = make_float2(m * ph.x, m * ph.y);
        }
        __syncthreads();
        fft_batch(a, b, nb, H, 1.f);
        for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
            const int h = i / nb, c = i - h * nb;
            plane[(long long)h * W + c0 + c] = a[c * H + h];
        }
    }
}

// inverse row FFT, real part, 1/(H*W) normalisation and the alpha blend
__global__ void __launch_bounds__(FFT_THREADS)
row_ifft_blend_kernel(const float2* __restrict__ ws, const float* __restrict__ x, float* __restrict__ out,
                      long long rows_total, int W, int rows_per_cta, float alpha, float inv_n) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const int nb = (int)min((long long)rows_per_cta, rows_total - r0);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) a[i] = ws[r0 * W + i];
    __syncthreads();
    fft_batch(a, b, nb, W, 1.f);
    const float beta = 1.f - alpha;
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS)
        out[r0 * W + i] = alpha * x[r0 * W + i] + beta * (a[i].x * inv_n);
}

// frequency terms of frequency_aware_loss (webp_training.py:114-126): column FFT of the row-transformed pred and
// target planes over the rfft2 half spectrum (columns 0..W/2), accumulating sum (|P|-|T|)^2 and
// sum (angle P - angle T)^2.  The row pass transformed x*0.5+0.5 (see row_fft_affine_kernel).
__global__ void __launch_bounds__(FFT_THREADS)
col_loss_kernel(const float2* __restrict__ wp, const float2* __restrict__ wt, int H, int W, int cols_per_cta,
                double* __restrict__ acc, int full) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const int p = blockIdx.y;
    const int half = W / 2 + 1;
    const int c0 = blockIdx.x * cols_per_cta;
    const int nb = min(cols_per_cta, half - c0);     // columns of pred; the same columns of target follow
    const float2* pp = wp + (long long)p * H * W;
    const float2* tp = wt + (long long)p * H * W;
    for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
        const int h = i / nb, c = i - h * nb;
        a[c * H + h] = pp[(long long)h * W + c0 + c];
        a[(nb + c) * H + h] = tp[(long long)h * W + c0 + c];
    }
    __syncthreads();
    fft_batch(a, b, 2 * nb, H, -1.f);
    float sm2 = 0.f, sp2 = 0.f;
    for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
        const float2 P = a[i], T = a[nb * H + i];
        const float dm = sqrtf(P.x * P.x + P.y * P.y) - sqrtf(T.x * T.x + T.y * T.y);
        const float dp = atan2f(P.y, P.x) - atan2f(T.y, T.x);
        // full = 1: the sums run over the whole fft2 spectrum (avif.py:150-158); its columns W/2+1 .. W-1 mirror 1 .. W/2-1
        // (conjugate symmetry of a real image: same magnitudes, negated angles), so those columns count twice
        const int col = c0 + i / H;
        const float wgt = (full && col > 0 && 2 * col < W) ? 2.f : 1.f;
        sm2 = fmaf(wgt * dm, dm, sm2); sp2 = fmaf(wgt * dp, dp, sp2);
    }
    sm2 = warp_sum(sm2); sp2 = warp_sum(sp2);
    __shared__ float red[2][FFT_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = sm2; red[1][wid] = sp2; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0;
        for (int w = 0; w < FFT_THREADS / 32; ++w) v += (double)red[threadIdx.x][w];
        atomicAdd(&acc[threadIdx.x], v);
    }
}

// backward of the frequency terms: G_k = d/dP_k [ wm (|P|-|T|)^2 + wp (angle P - angle T)^2 ] on the rfft2 half spectrum,
// then the inverse column transform; the inverse row transform + real part (row_ifft_accum_kernel) finishes
// d/dp = Re IDFT_unnormalised(G).  Columns > W/2 of wg stay zero (memset by the host wrapper).
__global__ void __launch_bounds__(FFT_THREADS)
col_loss_bwd_kernel(const float2* __restrict__ wp, const float2* __restrict__ wt, float2* __restrict__ wg, int H, int W,
                    int cols_per_cta, float w_mag, float w_phase, int full) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const int p = blockIdx.y;
    const int half = W / 2 + 1;
    const int c0 = blockIdx.x * cols_per_cta;
    const int nb = min(cols_per_cta, half - c0);
    const float2* pp = wp + (long long)p * H * W;
    const float2* tp = wt + (long long)p * H * W;
    for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
        const int h = i / nb, c = i - h * nb;
        a[c * H + h] = pp[(long long)h * W + c0 + c];
        a[(nb + c) * H + h] = tp[(long long)h * W + c0 + c];
    }
    __syncthreads();
    fft_batch(a, b, 2 * nb, H, -1.f);
    for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
        const float2 P = a[i], T = a[nb * H + i];
        const float mp = sqrtf(P.x * P.x + P.y * P.y), mt = sqrtf(T.x * T.x + T.y * T.y);
        float gx = 0.f, gy = 0.f;
        const int col = c0 + i / H;
        const float wgt = (full && col > 0 && 2 * col < W) ? 2.f : 1.f;            // mirrored columns of the full spectrum
        if (mp > 0.f) {
            const float cm = wgt * 2.f * w_mag * (mp - mt) / mp;                                  // d|P| = (Re, Im)/|P|
            const float cp = wgt * 2.f * w_phase * (atan2f(P.y, P.x) - atan2f(T.y, T.x)) / (mp * mp);   // d angle = (-Im, Re)/|P|^2
            gx = cm * P.x - cp * P.y;
            gy = cm * P.y + cp * P.x;
        }
        a[i] = make_float2(gx, gy);
    }
    __syncthreads();
    fft_batch(a, b, nb, H, 1.f);
    float2* gp = wg + (long long)p * H * W;
    for (int i = threadIdx.x; i < nb * H; i += FFT_THREADS) {
        const int h = i / nb, c = i - h * nb;
        gp[(long long)h * W + c0 + c] = a[c * H + h];
    }
}

// out[r, :] += coef * Re(inverse row FFT of ws[r, :])
__global__ void __launch_bounds__(FFT_THREADS)
row_ifft_accum_kernel(const float2* __restrict__ ws, float* __restrict__ out, long long rows_total, int W, int rows_per_cta, float coef) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const int nb = (int)min((long long)rows_per_cta, rows_total - r0);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) a[i] = ws[r0 * W + i];
    __syncthreads();
    fft_batch(a, b, nb, W, 1.f);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) out[r0 * W + i] += coef * a[i].x;
}

// forward row FFT of x*0.5+0.5 (the [0,1] images of webp_training.py:111-112)
__global__ void __launch_bounds__(FFT_THREADS)
row_fft_affine_kernel(const float* __restrict__ x, float2* __restrict__ ws, long long rows_total, int W, int rows_per_cta) {
    extern __shared__ float2 sm[];
    float2* a = sm;
    float2* b = sm + FFT_ELEMS;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const int nb = (int)min((long long)rows_per_cta, rows_total - r0);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) a[i] = make_float2(__fadd_rn(__fmul_rn(x[r0 * W + i], 0.5f), 0.5f), 0.f);
    __syncthreads();
    fft_batch(a, b, nb, W, -1.f);
    for (int i = threadIdx.x; i < nb * W; i += FFT_THREADS) ws[r0 * W + i] = a[i];
}

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int setup_smem() {
    static PerDevice done_dev;
    int& done = done_dev.cur();
    if (done) return DDPMIR_OK;
    const int bytes = 2 * FFT_ELEMS * sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(row_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(col_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(col_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(row_ifft_blend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(col_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(col_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(row_ifft_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(row_fft_affine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { ddpmir_set_error("fft: shared-memory opt-in failed: %s", cudaGetErrorString(e)); return DDPMIR_ERR_CUDA; }
    done = 1;
    return DDPMIR_OK;
}

int check_shape(int planes, int H, int W) {
    if (planes <= 0 || !pow2(H) || !pow2(W) || H < 2 || W < 2 || H > 1024 || W > 1024) {
        ddpmir_set_error("phase_consistency: H and W must be powers of two in [2, 1024] (got %d x %d)", H, W);
        return DDPMIR_ERR_UNSUPPORTED;
    }
    if (planes > 65535) { ddpmir_set_error("phase_consistency: too many planes"); return DDPMIR_ERR_INVALID; }
    return DDPMIR_OK;
}

constexpr int SMEM_BYTES = 2 * FFT_ELEMS * sizeof(float2);

}  // namespace

extern "C" int ddpmir_phase_reference(const float* ref, int planes, int H, int W, float* phasor, float* ws,
                                      ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(ref && phasor && ws, "phase_reference: null pointer");
    int rc = check_shape(planes, H, W);
    if (rc) return rc;
    if ((rc = setup_smem())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)planes * H;
    const int rpc = FFT_ELEMS / W, cpc = min(16, FFT_ELEMS / H);
    row_fft_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>(ref, (float2*)ws, rows, W, rpc);
    DDPMIR_LAUNCH_CHECK();
    col_pass_kernel<0><<<dim3(ceil_div(W, cpc), planes), FFT_THREADS, SMEM_BYTES, st>>>((float2*)ws, (float2*)phasor, H, W, cpc);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

extern "C" int ddpmir_phase_consistency(const float* x, const float* phasor, float alpha, int planes, int H, int W,
                                        float* out, float* ws, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(x && phasor && out && ws, "phase_consistency: null pointer");
    int rc = check_shape(planes, H, W);
    if (rc) return rc;
    if ((rc = setup_smem())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)planes * H;
    const int rpc = FFT_ELEMS / W, cpc = min(16, FFT_ELEMS / H);
    row_fft_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>(x, (float2*)ws, rows, W, rpc);
    DDPMIR_LAUNCH_CHECK();
    col_pass_kernel<1><<<dim3(ceil_div(W, cpc), planes), FFT_THREADS, SMEM_BYTES, st>>>((float2*)ws, (float2*)const_cast<float*>(phasor), H, W, cpc);
    DDPMIR_LAUNCH_CHECK();
    row_ifft_blend_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>((const float2*)ws, x, out, rows, W, rpc, alpha,
                                                                                1.f / ((float)H * (float)W));
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}

static int freq_loss_terms_impl(const float* pred, const float* target, int planes, int H, int W, float* ws_pred,
                                float* ws_target, double* acc2, int full, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(pred && target && ws_pred && ws_target && acc2, "freq_loss_terms: null pointer");
    int rc = check_shape(planes, H, W);
    if (rc) return rc;
    if ((rc = setup_smem())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st);
    const long long rows = (long long)planes * H;
    const int rpc = FFT_ELEMS / W;
    int cpc = FFT_ELEMS / H / 2;      // pred + target columns share one shared-memory batch
    if (cpc > 8) cpc = 8;
    if (cpc < 1) { ddpmir_set_error("freq_loss_terms: H too large"); return DDPMIR_ERR_UNSUPPORTED; }
    row_fft_affine_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>(pred, (float2*)ws_pred, rows, W, rpc);
    row_fft_affine_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>(target, (float2*)ws_target, rows, W, rpc);
    DDPMIR_LAUNCH_CHECK();
    col_loss_kernel<<<dim3(ceil_div(W / 2 + 1, cpc), planes), FFT_THREADS, SMEM_BYTES, st>>>((const float2*)ws_pred, (const float2*)ws_target,
                                                                                          H, W, cpc, acc2, full);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
extern "C" int ddpmir_freq_loss_terms(const float* pred, const float* target, int planes, int H, int W, float* ws_pred,
                                      float* ws_target, double* acc2, ddpmir_stream_t stream) {
    return freq_loss_terms_impl(pred, target, planes, H, W, ws_pred, ws_target, acc2, 0, stream);
}
extern "C" int ddpmir_fft2_loss_terms(const float* pred, const float* target, int planes, int H, int W, float* ws_pred,
                                      float* ws_target, double* acc2, ddpmir_stream_t stream) {
    return freq_loss_terms_impl(pred, target, planes, H, W, ws_pred, ws_target, acc2, 1, stream);
}

// dpred += d/dpred [ w_mag * sum (|P|-|T|)^2 + w_phase * sum (angle P - angle T)^2 ],  P = rfft2(pred*0.5+0.5)
static int freq_loss_backward_impl(const float* pred, const float* target, int planes, int H, int W, float w_mag, float w_phase,
                                   float* ws_pred, float* ws_target, float* ws_grad, float* dpred, int full, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(pred && target && ws_pred && ws_target && ws_grad && dpred, "freq_loss_backward: null pointer");
    int rc = check_shape(planes, H, W);
    if (rc) return rc;
    if ((rc = setup_smem())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)planes * H;
    const int rpc = FFT_ELEMS / W;
    int cpc = FFT_ELEMS / H / 2;
    if (cpc > 8) cpc = 8;
    if (cpc < 1) { ddpmir_set_error("freq_loss_backward: H too large"); return DDPMIR_ERR_UNSUPPORTED; }
    cudaMemsetAsync(ws_grad, 0, sizeof(float2) * rows * W, st);
    row_fft_affine_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>(pred, (float2*)ws_pred, rows, W, rpc);
    row_fft_affine_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>(target, (float2*)ws_target, rows, W, rpc);
    col_loss_bwd_kernel<<<dim3(ceil_div(W / 2 + 1, cpc), planes), FFT_THREADS, SMEM_BYTES, st>>>((const float2*)ws_pred, (const float2*)ws_target,
                                                                                              (float2*)ws_grad, H, W, cpc, w_mag, w_phase, full);
    // chain rule of p01 = 0.5 * pred + 0.5
    row_ifft_accum_kernel<<<ceil_div(rows, rpc), FFT_THREADS, SMEM_BYTES, st>>>((const float2*)ws_grad, dpred, rows, W, rpc, 0.5f);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
extern "C" int ddpmir_freq_loss_backward(const float* pred, const float* target, int planes, int H, int W, float w_mag, float w_phase,
                                         float* ws_pred, float* ws_target, float* ws_grad, float* dpred, ddpmir_stream_t stream) {
    return freq_loss_backward_impl(pred, target, planes, H, W, w_mag, w_phase, ws_pred, ws_target, ws_grad, dpred, 0, stream);
}
extern "C" int ddpmir_fft2_loss_backward(const float* pred, const float* target, int planes, int H, int W, float w_mag, float w_phase,
                                         float* ws_pred, float* ws_target, float* ws_grad, float* dpred, ddpmir_stream_t stream) {
    return freq_loss_backward_impl(pred, target, planes, H, W, w_mag, w_phase, ws_pred, ws_target, ws_grad, dpred, 1, stream);
}
