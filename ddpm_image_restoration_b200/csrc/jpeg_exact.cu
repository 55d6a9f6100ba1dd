// Bit-exact baseline-JPEG round trip on the GPU: what `Image.save(format="JPEG", quality=q, subsampling=...)` followed by
// `Image.open(...)` does to the pixels (jpeg_compress, svd.ipynb#c1:L20-44 and 0409_method.ipynb#c0:L44-62), without the
// entropy coder, which is lossless.  The arithmetic is libjpeg-turbo's (the library behind Pillow; its SIMD paths are
// bit-identical to its C code by design), all of it integer:
//   encoder: RGB -> YCbCr (jccolor.c, 16-bit fixed point), h2v2 chroma downsampling with the alternating 1,2 bias
//            (jcsample.c), forward DCT jpeg_fdct_islow (jfdctint.c: LL&M, CONST_BITS 13, PASS1_BITS 2, output scaled by 8),
//            quantisation round-half-away division by 8*Q (jcdctmgr.c), tables from jpeg_quality_scaling (jcparam.c);
//   decoder: dequantise + jpeg_idct_islow (jidctint.c), "fancy" triangle-filter h2v2 upsampling (jdsample.c),
//            YCbCr -> RGB (jdcolor.c).
// oracle/jpeg_exact.py restates the same pipeline in numpy and is pinned bit-exactly against Pillow over all qualities.
// Any image size: planes are padded to whole MCUs (16 with 4:2:0, 8 with 4:4:4) the way libjpeg pads them.
#include "common.cuh"
#include <stdint.h>

namespace {

constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633, F_1_501 = 12299,
              F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// jfdctint.c, one 8-point pass.  FIRST: rows (outputs scaled up by 2^PASS1_BITS), else columns (scaled back down)
template <bool FIRST> __device__ __forceinline__ void fdct8(const int (&d)[8], int (&o)[8]) {
    int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
    int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    constexpr int SH = FIRST ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS;
    if (FIRST) { o[0] = (tmp10 + tmp11) << PASS1_BITS; o[4] = (tmp10 - tmp11) << PASS1_BITS; }
    else { o[0] = descale(tmp10 + tmp11, PASS1_BITS); o[4] = descale(tmp10 - tmp11, PASS1_BITS); }
    int z1 = (tmp12 + tmp13) * F_0_541;
    o[2] = descale(z1 + tmp13 * F_0_765, SH);
    o[6] = descale(z1 + tmp12 * (-F_1_847), SH);
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * F_1_175;
    tmp4 *= F_0_298; tmp5 *= F_2_053; tmp6 *= F_3_072; tmp7 *= F_1_501;
    z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
    z3 += z5; z4 += z5;
    o[7] = descale(tmp4 + z1 + z3, SH); o[5] = descale(tmp5 + z2 + z4, SH);
    o[3] = descale(tmp6 + z2 + z3, SH); o[1] = descale(tmp7 + z1 + z4, SH);
}

// jidctint.c, one 8-point pass.  FIRST: columns of dequantised coefficients, else rows (final descale includes the /8)
template <bool FIRST> __device__ __forceinline__ void idct8(const int (&c)[8], int (&o)[8]) {
    int z2 = c[2], z3 = c[6];
    int z1 = (z2 + z3) * F_0_541;
    int tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    z2 = c[0]; z3 = c[4];
    int tmp0 = (z2 + z3) << CONST_BITS, tmp1 = (z2 - z3) << CONST_BITS;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = c[7]; tmp1 = c[5]; tmp2 = c[3]; tmp3 = c[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298; tmp1 *= F_2_053; tmp2 *= F_3_072; tmp3 *= F_1_501;
    z1 *= -F_0_899; z2 *= -F_2_562; z3 *= -F_1_961; z4 *= -F_0_390;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    constexpr int SH = FIRST ? CONST_BITS - PASS1_BITS : CONST_BITS + PASS1_BITS + 3;
    o[0] = descale(tmp10 + tmp3, SH); o[7] = descale(tmp10 - tmp3, SH);
    o[1] = descale(tmp11 + tmp2, SH); o[6] = descale(tmp11 - tmp2, SH);
    o[2] = descale(tmp12 + tmp1, SH); o[5] = descale(tmp12 - tmp1, SH);
    o[3] = descale(tmp13 + tmp0, SH); o[4] = descale(tmp13 - tmp0, SH);
}

struct QTables { int q[2][64]; };   // [0] luma, [1] chroma, natural (row-major) order

inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > 148LL * 64 ? 148LL * 64 : g));
}

// RGB (uint8 HWC, H x W) -> Y plane (Hp x Wp) and Cb / Cr planes (Hp x Wp, or Hp/2 x Wp/2 when SUB), Hp / Wp = size rounded up
// to the MCU.  One thread per 2x2 quad of the PADDED image.  Padding as libjpeg does it: columns are replicated at full
// resolution before downsampling (jcsample.c expand_right_edge); rows are replicated at full resolution only up to an even
// height, beyond that every plane repeats ITS last row (jcprepct.c expand_bottom_edge on the downsampled output).
template <bool SUB>
__global__ void __launch_bounds__(256)
rgb_to_ycc_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ yp, uint8_t* __restrict__ cbp, uint8_t* __restrict__ crp,
                  int H, int W, int Hp, int Wp, long long total) {
    constexpr int FIX_299 = 19595, FIX_587 = 38470, FIX_114 = 7471, FIX_16874 = 11059, FIX_33126 = 21709, FIX_5 = 32768,
                  FIX_41869 = 27439, FIX_08131 = 5329, HALF = 1 << 15, OFF = 128 << 16;
    const int W2 = Wp >> 1, H2 = Hp >> 1, ch = (H + 1) >> 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int qx = (int)(i % W2);
        long long r = i / W2;
        const int qy = (int)(r % H2);
        const long long b = r / H2;
        const uint8_t* img = rgb + b * H * W * 3;
        int cbs = 0, crs = 0;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int yo = 2 * qy + dy, xo = 2 * qx + dx;                 // position in the padded planes
                const int x = min(xo, W - 1);
                {
                    const uint8_t* p = img + ((long long)min(yo, H - 1) * W + x) * 3;
                    const int R = p[0], G = p[1], B = p[2];
                    yp[(b * Hp + yo) * Wp + xo] = (uint8_t)((FIX_299 * R + FIX_587 * G + FIX_114 * B + HALF) >> 16);
                    if (!SUB) {
                        cbp[(b * Hp + yo) * Wp + xo] = (uint8_t)((-FIX_16874 * R - FIX_33126 * G + FIX_5 * B + OFF + HALF - 1) >> 16);
                        crp[(b * Hp + yo) * Wp + xo] = (uint8_t)((FIX_5 * R - FIX_41869 * G - FIX_08131 * B + OFF + HALF - 1) >> 16);
                    }
                }
                if (SUB) {      // chroma quad: rows of the last REAL chroma row when this quad lies below the image
                    const int yc = min(2 * min(qy, ch - 1) + dy, H - 1);
                    const uint8_t* p = img + ((long long)yc * W + x) * 3;
                    const int R = p[0], G = p[1], B = p[2];
                    cbs += (-FIX_16874 * R - FIX_33126 * G + FIX_5 * B + OFF + HALF - 1) >> 16;
                    crs += (FIX_5 * R - FIX_41869 * G - FIX_08131 * B + OFF + HALF - 1) >> 16;
                }
            }
        if (SUB) {
            const int bias = (qx & 1) ? 2 : 1;       // jcsample.c h2v2_downsample: bias = 1, 2, 1, 2, ... along the row
            cbp[(b * H2 + qy) * W2 + qx] = (uint8_t)((cbs + bias) >> 2);
            crp[(b * H2 + qy) * W2 + qx] = (uint8_t)((crs + bias) >> 2);
        }
    }
}

// forward DCT -> quantise -> dequantise -> inverse DCT of every 8x8 block of a set of uint8 planes, in place.
// 8 threads per block (a row, then a column, then a row again), 32 blocks per CTA.
__global__ void __launch_bounds__(256)
block_roundtrip_kernel(uint8_t* __restrict__ planes, long long nblk_y, int ybw, int yW, long long nblk_c, int cbw, int cW,
                       long long y_bytes, long long c_bytes, const __grid_constant__ QTables Q) {
    __shared__ int tile[32][8][9];
    const int lb = threadIdx.x >> 3, k = threadIdx.x & 7;
    const long long blk = (long long)blockIdx.x * 32 + lb;
    const long long total = nblk_y + 2 * nblk_c;
    const bool ok = blk < total;
    // which plane set, which block
    uint8_t* base = planes;
    long long bi = blk;
    int bw = ybw, Wp = yW, table = 0;
    if (ok && blk >= nblk_y) {
        bi = blk - nblk_y;
        base = planes + y_bytes;
        if (bi >= nblk_c) { bi -= nblk_c; base += c_bytes; }
        bw = cbw; Wp = cW; table = 1;
    }
    // blocks are numbered row-major over (image*block_rows, block_cols); planes of one kind are contiguous [B, Hp, Wp]
    const long long brow = ok ? bi / bw : 0;
    const int bcol = ok ? (int)(bi % bw) : 0;
    uint8_t* p = base + (brow * 8) * Wp + bcol * 8;
    int d[8], o[8];
    if (ok) {
        const uint2 v = *reinterpret_cast<const uint2*>(p + (long long)k * Wp);       // row k
        d[0] = (int)(v.x & 255) - 128; d[1] = (int)((v.x >> 8) & 255) - 128; d[2] = (int)((v.x >> 16) & 255) - 128; d[3] = (int)(v.x >> 24) - 128;
        d[4] = (int)(v.y & 255) - 128; d[5] = (int)((v.y >> 8) & 255) - 128; d[6] = (int)((v.y >> 16) & 255) - 128; d[7] = (int)(v.y >> 24) - 128;
        fdct8<true>(d, o);
#pragma unroll
        for (int c = 0; c < 8; ++c) tile[lb][k][c] = o[c];
    }
    __syncwarp();
    if (ok) {
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] = tile[lb][r][k];                              // column k
        fdct8<false>(d, o);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int q = Q.q[table][r * 8 + k], dv = q << 3;
            int t = o[r];
            const int a = (t < 0 ? -t : t) + (dv >> 1);
            const int c = a / dv;
            d[r] = (t < 0 ? -c : c) * q;                                                 // quantised, then dequantised
        }
        idct8<true>(d, o);
#pragma unroll
        for (int r = 0; r < 8; ++r) tile[lb][r][k] = o[r];
    }
    __syncwarp();
    if (ok) {
#pragma unroll
        for (int c = 0; c < 8; ++c) d[c] = tile[lb][k][c];                              // row k
        idct8<false>(d, o);
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int v = min(255, max(0, o[c] + 128));
            w[c >> 2] |= (uint32_t)v << (8 * (c & 3));
        }
        *reinterpret_cast<uint2*>(p + (long long)k * Wp) = make_uint2(w[0], w[1]);
    }
}

// decoded planes (padded, see above) -> RGB uint8 HWC of the real H x W image; SUB: chroma is half resolution and goes through
// jdsample.c's h2v2 fancy upsampling over the REAL ceil(H/2) x ceil(W/2) samples (edge cases at the real edge), or through
// plain replication when the real chroma width is <= 2 (jinit_upsampler)
template <bool SUB>
__global__ void __launch_bounds__(256)
ycc_to_rgb_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ cbp, const uint8_t* __restrict__ crp,
                  uint8_t* __restrict__ rgb, int H, int W, int Hp, int Wp, long long total) {
    constexpr int FIX_1402 = 91881, FIX_1772 = 116130, FIX_71414 = 46802, FIX_34414 = 22554, HALF = 1 << 15;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int y = (int)(r % H);
        const long long b = r / H;
        int cb, cr;
        if (SUB) {
            const int W2 = Wp >> 1, H2 = Hp >> 1, cw = (W + 1) >> 1, chh = (H + 1) >> 1;
            const int cy = y >> 1, cx = x >> 1;
            const int fy = (y & 1) ? min(cy + 1, chh - 1) : max(cy - 1, 0);      // the farther of the two nearest chroma rows
            auto up = [&](const uint8_t* pl) {
                const uint8_t* near = pl + (b * H2 + cy) * W2;
                if (cw <= 2) return (int)near[cx];
                const uint8_t* far = pl + (b * H2 + fy) * W2;
                const int cs = 3 * near[cx] + far[cx];                                   // thiscolsum
                if (x & 1) {
                    if (cx == cw - 1) return (cs * 4 + 7) >> 4;
                    return (3 * cs + (3 * near[cx + 1] + far[cx + 1]) + 7) >> 4;
                }
                if (cx == 0) return (cs * 4 + 8) >> 4;
                return (3 * cs + (3 * near[cx - 1] + far[cx - 1]) + 8) >> 4;
            };
            cb = up(cbp); cr = up(crp);
        } else {
            cb = cbp[(b * Hp + y) * Wp + x]; cr = crp[(b * Hp + y) * Wp + x];
        }
        const int Y = yp[(b * Hp + y) * Wp + x], xb = cb - 128, xr = cr - 128;
        const int R = Y + ((FIX_1402 * xr + HALF) >> 16);
        const int B = Y + ((FIX_1772 * xb + HALF) >> 16);
        const int G = Y + ((-FIX_34414 * xb + HALF - FIX_71414 * xr) >> 16);
        uint8_t* o = rgb + i * 3;
        o[0] = (uint8_t)min(255, max(0, R)); o[1] = (uint8_t)min(255, max(0, G)); o[2] = (uint8_t)min(255, max(0, B));
    }
}

}  // namespace

extern "C" size_t ddpmir_jpeg_roundtrip_workspace(int B, int H, int W) {
    return (size_t)B * ((H + 15) / 16 * 16) * ((W + 15) / 16 * 16) * 3;
}

extern "C" int ddpmir_jpeg_roundtrip_u8(const uint8_t* rgb, uint8_t* out, int B, int H, int W, int quality, int subsample_420,
                                        void* workspace, ddpmir_stream_t stream) {
    DDPMIR_CHECK_ARG(rgb && out && workspace && B > 0 && H > 0 && W > 0, "jpeg_roundtrip: bad arguments");
    const int mcu = subsample_420 ? 16 : 8;
    const int Hp = (H + mcu - 1) / mcu * mcu, Wp = (W + mcu - 1) / mcu * mcu;       // planes padded to whole MCUs
    static const int QY[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56,
                               14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const int QC[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99,
                               47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                               99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    // jcparam.c: jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
    int q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
    QTables T;
    for (int i = 0; i < 64; ++i) {
        long ty = ((long)QY[i] * scale + 50) / 100, tc = ((long)QC[i] * scale + 50) / 100;
        T.q[0][i] = (int)(ty < 1 ? 1 : (ty > 255 ? 255 : ty));
        T.q[1][i] = (int)(tc < 1 ? 1 : (tc > 255 ? 255 : tc));
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    const long long y_bytes = (long long)B * Hp * Wp;
    const int cH = subsample_420 ? Hp / 2 : Hp, cW = subsample_420 ? Wp / 2 : Wp;
    const long long c_bytes = (long long)B * cH * cW;
    uint8_t *yp = ws, *cbp = ws + y_bytes, *crp = cbp + c_bytes;
    const long long quads = (long long)B * (Hp / 2) * (Wp / 2);
    if (subsample_420) rgb_to_ycc_kernel<true><<<grid_for(quads, 256), 256, 0, st>>>(rgb, yp, cbp, crp, H, W, Hp, Wp, quads);
    else rgb_to_ycc_kernel<false><<<grid_for(quads, 256), 256, 0, st>>>(rgb, yp, cbp, crp, H, W, Hp, Wp, quads);
    DDPMIR_LAUNCH_CHECK();
    const long long nblk_y = (long long)B * (Hp / 8) * (Wp / 8), nblk_c = (long long)B * (cH / 8) * (cW / 8);
    const long long nblk = nblk_y + 2 * nblk_c;
    DDPMIR_CHECK_ARG((nblk + 31) / 32 <= 2147483647LL, "jpeg_roundtrip: too many blocks");
    block_roundtrip_kernel<<<(unsigned)((nblk + 31) / 32), 256, 0, st>>>(ws, nblk_y, Wp / 8, Wp, nblk_c, cW / 8, cW, y_bytes, c_bytes, T);
    DDPMIR_LAUNCH_CHECK();
    const long long px = (long long)B * H * W;
    if (subsample_420) ycc_to_rgb_kernel<true><<<grid_for(px, 256), 256, 0, st>>>(yp, cbp, crp, out, H, W, Hp, Wp, px);
    else ycc_to_rgb_kernel<false><<<grid_for(px, 256), 256, 0, st>>>(yp, cbp, crp, out, H, W, Hp, Wp, px);
    DDPMIR_LAUNCH_CHECK();
    return DDPMIR_OK;
}
