"""tcgen05/TMA implicit-GEMM kernel (IMPL_TENSOR) against PyTorch fp32 on bf16-rounded operands and against the
generic SIMT kernel, over the shape classes of the UNet: pixel boxes that are part of a row (W >= 128), several rows,
a whole image, several images, ragged N tiles, every epilogue mode."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import restated as R
from util import bf16_round, nchw, nhwc, pack1, pack3, rel

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from ddpm_image_restoration_b200 import ops as o
    return o


def g(seed):
    return torch.Generator().manual_seed(seed)


CONV_SHAPES = [  # B, Cin, Cout, H, W
    (2, 64, 128, 16, 16),    # 8 rows per tile
    (1, 128, 64, 8, 8),      # M = 64 < 128: half-empty tile, batch OOB
    (4, 64, 64, 4, 4),       # 8 images per tile
    (3, 64, 192, 8, 8),      # ragged M (192 pixels), N = 3 x 64
    (1, 64, 64, 256, 256),   # W >= 128: half-row boxes
    (2, 1024, 512, 8, 8),    # long K loop (144 k-blocks)
    (2, 512, 1024, 2, 2),    # tiny spatial, 32 images per box
    (1, 128, 32, 16, 16),    # N smaller than the N tile
    (1, 64, 64, 32, 128),    # non-square, W = 128
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_tensor_core(ops, shape):
    B, Ci, Co, H, W = shape
    x = bf16_round(torch.randn(B, Ci, H, W, generator=g(1)))
    w = bf16_round(torch.randn(Co, Ci, 3, 3, generator=g(2)) / math.sqrt(9 * Ci))
    b = torch.randn(Co, generator=g(3))
    tb = torch.randn(B, Co, generator=g(4))
    res = torch.randn(B, Co, H, W, generator=g(5))
    ref = F.conv2d(x, w, b, padding=1) + tb[:, :, None, None] + res
    out, out2 = ops.conv3x3(nhwc(x, BF), pack3(w, BF), Co, ops.IMPL_TENSOR, out_dtype=torch.float32, out2_dtype=BF,
                            bias=b.cuda(), row_bias=tb.cuda(), res=nhwc(res))
    assert rel(nchw(out), ref) < 2e-5
    assert torch.equal(out2, out.to(BF))
    simt = ops.conv3x3(nhwc(x, BF), pack3(w, BF), Co, ops.IMPL_SIMT, out_dtype=torch.float32, bias=b.cuda(),
                       row_bias=tb.cuda(), res=nhwc(res))
    assert rel(out, simt) < 2e-5


@pytest.mark.parametrize("shape", [(2, 64, 192, 16, 16), (1, 1024, 1024, 8, 8), (2, 128, 64, 64, 64), (1, 64, 3072 // 16, 4, 4),
                                   (5, 256, 256, 2, 2), (1, 64, 64, 256, 256)])
def test_gemm_tensor_core(ops, shape):
    B, K, N, H, W = shape
    x = bf16_round(torch.randn(B, K, H, W, generator=g(1)))
    w = bf16_round(torch.randn(N, K, 1, 1, generator=g(2)) / math.sqrt(K))
    b = torch.randn(N, generator=g(3))
    scale = torch.rand(B, generator=g(4)) + 0.5
    mul = bf16_round(torch.randn(B, N, H, W, generator=g(5)))
    ref = torch.sigmoid(F.conv2d(x, w, b)) * scale.view(-1, 1, 1, 1) * mul
    out = ops.gemm(nhwc(x, BF), pack1(w, BF), N, ops.IMPL_TENSOR, out_dtype=torch.float32, bias=b.cuda(), act=ops.ACT_SIGMOID,
                   img_scale=scale.cuda(), mul=nhwc(mul, BF))
    assert rel(nchw(out), ref) < 2e-5


@pytest.mark.parametrize("fam,hw", [("webp", (16, 16)), ("jpeg", (16, 32)), ("webp", (4, 4))])
def test_freq_gate_epilogues_tensor_core(ops, fam, hw):
    C, B = 64, 2
    H, W = hw
    f = R.FAMILY[fam]
    gen = g(11)
    rnd = lambda *s: torch.randn(*s, generator=gen)
    d = bf16_round(rnd(B, C, H, W))
    h3 = rnd(B, C, H, W)
    w1 = bf16_round(rnd(C, C) / 8); b1 = rnd(C) * 0.3
    w2 = bf16_round(rnd(C, C) / 5); b2l = rnd(C) * 0.3; b2h = rnd(C) * 0.3
    boost = torch.tensor([0.8, 0.15])
    kw = dict(bs=f["bs"], low=f["low"])
    for impl in (ops.IMPL_SIMT, ops.IMPL_TENSOR):
        g1 = ops.gemm(nhwc(d, BF), w1.to(BF).cuda(), C, impl, bias=b1.cuda(), act=ops.ACT_LRELU02, freq_mode=1, **kw)
        e = ops.gemm(g1, w2.to(BF).cuda(), C, impl, out_dtype=torch.float32, bias=b2l.cuda(), bias2=b2h.cuda(),
                     act=ops.ACT_SIGMOID, freq_mode=2, img_scale=boost.cuda(), mul=nhwc(d, BF), res=nhwc(h3), **kw)
        if impl == ops.IMPL_SIMT:
            base_g1, base_e = g1, e
        else:
            assert rel(g1, base_g1) < 1e-2 and torch.equal(g1 == 0, base_g1 == 0)
            assert rel(e, base_e) < 1e-5
    # and against the closed form
    m = R.low_mask(H, W, f["bs"], f["low"]).float()
    hid = F.leaky_relu(F.conv2d(d, w1.view(C, C, 1, 1), b1), 0.2)
    hid = bf16_round(torch.cat([hid[:, :C // 2] * m, hid[:, C // 2:] * (1 - m)], 1))
    z = F.conv2d(hid, w2.view(C, C, 1, 1))
    gate = torch.sigmoid(z + torch.where(m.bool(), b2l.view(1, C, 1, 1), b2h.view(1, C, 1, 1)))
    sc = torch.where(m.bool(), torch.ones(1), boost.view(B, 1, 1, 1))
    assert rel(nchw(base_e), h3 + gate * sc * d) < 2e-5


def test_auto_dispatch_prefers_tensor_core_and_falls_back(ops):
    # Cin = 32 is not a multiple of the 64-wide K block -> generic kernel through IMPL_AUTO, explicit TENSOR refuses
    from ddpm_image_restoration_b200._lib import DdpmirError
    x = bf16_round(torch.randn(1, 32, 8, 8, generator=g(1)))
    w = bf16_round(torch.randn(64, 32, 3, 3, generator=g(2)) / 17)
    out = ops.conv3x3(nhwc(x, BF), pack3(w, BF), 64, ops.IMPL_AUTO, out_dtype=torch.float32)
    assert rel(nchw(out), F.conv2d(x, w, padding=1)) < 2e-5
    with pytest.raises(DdpmirError):
        ops.conv3x3(nhwc(x, BF), pack3(w, BF), 64, ops.IMPL_TENSOR)


@pytest.mark.parametrize("shape", [(2, 64, 64, 64, 64, 9), (1, 64, 192, 64, 64, 1), (3, 128, 64, 32, 32, 1), (1, 64, 32, 32, 64, 9),
                                   (2, 64, 128, 16, 16, 9), (1, 64, 256, 256, 256, 1)])
def test_streaming_weight_resident_kernel(ops, shape):
    """The persistent, weight-resident tcgen05 kernel (forced with variant 2) against the tiled one and PyTorch."""
    from ddpm_image_restoration_b200 import _lib
    B, Ci, Co, H, W, taps = shape
    x = bf16_round(torch.randn(B, Ci, H, W, generator=g(1)))
    ks = 3 if taps == 9 else 1
    w = bf16_round(torch.randn(Co, Ci, ks, ks, generator=g(2)) / math.sqrt(taps * Ci))
    b = torch.randn(Co, generator=g(3))
    res = torch.randn(B, Co, H, W, generator=g(5))
    ref = F.gelu(F.conv2d(x, w, b, padding=ks // 2)) + res
    fn = ops.conv3x3 if taps == 9 else ops.gemm
    pk = pack3(w, BF) if taps == 9 else pack1(w, BF)
    outs = {}
    for variant in (1, 2):
        _lib.lib().ddpmir_igemm_set_variant(variant)
        try:
            outs[variant] = fn(nhwc(x, BF), pk, Co, ops.IMPL_TENSOR, out_dtype=torch.float32, bias=b.cuda(), act=ops.ACT_GELU,
                               res=nhwc(res))
        finally:
            _lib.lib().ddpmir_igemm_set_variant(0)
    assert rel(nchw(outs[2]), ref) < 2e-5
    assert rel(outs[2], outs[1]) < 1e-6
