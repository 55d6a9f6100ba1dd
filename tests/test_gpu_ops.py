"""GPU parity tests, one per C-ABI kernel, against plain PyTorch fp32 on the CPU / the restated oracle.
Tolerances: fp32 check mode <= 1e-5 relative L2 (north_star); bf16 mode <= 1e-2 (per-op it is ~3e-3);
integer/byte work and the injected-noise update are bit-exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restated as R
from util import bf16_round, nchw, nhwc, pack1, pack3, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from ddpm_image_restoration_b200 import ops as o
    return o


def g(seed):
    return torch.Generator().manual_seed(seed)


# ---------------------------------------------------------------------------------------------------------
def test_philox_known_answer(ops):
    n = 4096 + 3
    z = ops.philox_normal((n,), seed=7, step=3).cpu().numpy()
    ref = R.philox_normal(7, 3, n)
    assert np.abs(z - ref).max() < 2e-5          # fp32 log/sincos vs float64
    big = ops.philox_normal((1 << 22,), seed=(5 << 32) | 9, step=11).cpu()
    assert abs(float(big.mean())) < 3e-3 and abs(float(big.std()) - 1) < 3e-3


def test_ddrm_update_bit_exact(ops, golden):
    d = golden("ops.npz")
    xn, c, y, z, t = (torch.from_numpy(d[k]) for k in ("xn", "webp_q10", "x", "z", "t"))
    for eta_b, key in ((1.0, "update_eta1"), (0.6, "update_eta06")):
        out = ops.ddrm_update(xn.cuda(), c.cuda(), y.cuda(), t.cuda(), 0.2, 0.85, eta_b, z=z.cuda()).cpu()
        assert torch.equal(out, torch.from_numpy(d[key]))
    last = ops.ddrm_update(xn.cuda(), c.cuda(), y.cuda(), t.cuda(), 0.2, last_step=True).cpu()
    assert torch.equal(last, xn - c + y)
    # uint8 HWC decoder pixels path == fp32 path
    u8 = ((c * 0.5 + 0.5) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    cf = u8.permute(0, 3, 1, 2).float().div(255).sub(0.5).mul(2.0).contiguous()
    a = ops.ddrm_update(xn.cuda(), u8.cuda(), y.cuda(), t.cuda(), 0.2, z=z.cuda()).cpu()
    b = ops.ddrm_update(xn.cuda(), cf.cuda(), y.cuda(), t.cuda(), 0.2, z=z.cuda()).cpu()
    assert torch.equal(a, b)
    assert torch.equal(ops.u8_hwc_to_nchw(u8.cuda()).cpu(), cf)


def test_ddrm_update_philox_and_offset(ops):
    B, C, H, W = 4, 3, 16, 16
    x, c, y = (torch.randn(B, C, H, W, generator=g(i)) for i in range(3))
    t = torch.full((B,), 0.5)
    z = torch.from_numpy(R.philox_normal(3, 9, x.numel()).astype(np.float32)).view_as(x)
    ref = R.ddrm_update(x, c, y, z, t, 0.15)
    out = ops.ddrm_update(x.cuda(), c.cuda(), y.cuda(), t.cuda(), 0.15, seed=3, step=9).cpu()
    assert (out - ref).abs().max() < 1e-5
    # micro-batch slices with noise_offset reproduce the full-batch stream
    parts = [ops.ddrm_update(x[s:s + 2].cuda(), c[s:s + 2].cuda(), y[s:s + 2].cuda(), t[s:s + 2].cuda(), 0.15, seed=3,
                             step=9, noise_offset=s * C * H * W).cpu() for s in (0, 2)]
    assert torch.equal(torch.cat(parts), out)


def test_gmm_update_and_lincomb(ops):
    x, p, y, s, z = (torch.randn(2, 3, 32, 32, generator=g(10 + i)) for i in range(5))
    gs = 0.21
    pred = (1 - gs) * p + gs * (y - s)
    x0 = x + pred
    for first in (True, False):
        mean = x0 * 0.9 + x * 0.1 if first else x0 * 1.1 - x * 0.1
        ref = mean + 0.07 * z
        out = ops.gmm_update(x.cuda(), p.cuda(), y.cuda(), s.cuda(), gs, z=z.cuda(), use_first=first, noise_scale=0.07).cpu()
        assert (out - ref).abs().max() < 1e-6
    out = ops.gmm_update(x.cuda(), p.cuda(), last_step=True).cpu()
    assert torch.equal(out, x + p)
    # classical DDPM posterior mean (ddpm.ipynb#c5)
    tt = 57
    betas, alphas, abar = R.ddpm_schedule(100)
    ref = R.ddpm_posterior_mean(x, p, tt)
    wa = float(1 / torch.sqrt(alphas[tt])); wb = float(-(1 - alphas[tt]) / torch.sqrt(1 - abar[tt]) / torch.sqrt(alphas[tt]))
    out = ops.lincomb(x.cuda(), wa, p.cuda(), wb).cpu()
    assert rel(out, ref) < 1e-6


def test_quantize_u8(ops):
    x = torch.randn(3, 3, 24, 40, generator=g(1)) * 0.8
    x[0, 0, 0, :8] = torch.tensor([-1.0, 1.0, -1.2, 1.2, 0.0, 0.999, -0.999, 0.5])
    out = ops.quantize_u8_hwc(x.cuda()).cpu()
    assert torch.equal(out, R.quantize_u8(x).permute(0, 2, 3, 1))


def test_color_l1(ops, golden):
    d = golden("ops.npz")
    xn, x = torch.from_numpy(d["xn"]) * 1.5, torch.from_numpy(d["x"])
    out = float(ops.color_l1(xn.cuda(), x.cuda()))
    assert abs(out - float(d["color_deep"])) < 1e-6
    assert abs(out - float(R.color_l1(xn, x))) < 1e-6


def test_losses_forward_values():
    import ddpm_image_restoration_b200 as P
    gen = g(21)
    target = (torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1)
    pred = (target + 0.2 * torch.randn(2, 3, 64, 64, generator=gen))
    # oracle pieces (restated from webp_training.py:105-132)
    p01, t01 = pred * 0.5 + 0.5, target * 0.5 + 0.5
    freq = 0
    for c in range(3):
        fp, ft = torch.fft.rfft2(p01[:, c]), torch.fft.rfft2(t01[:, c])
        freq = freq + F.mse_loss(fp.abs(), ft.abs()) + 0.5 * F.mse_loss(torch.angle(fp), torch.angle(ft))
    ref = F.mse_loss(pred, target) + 0.5 * freq + 0.3 * (1 - R.ssim(p01, t01, 1.0))
    got = float(P.frequency_aware_loss(pred.cuda(), target.cuda()))
    # the phase term is ill-conditioned where a coefficient is ~0 (angle jumps by 2 pi on fp32 round-off): 1e-3 relative
    assert abs(got - float(ref)) < 1e-3 * abs(float(ref))
    from ddpm_image_restoration_b200 import ops as o
    assert abs(float(o.mse(pred.cuda(), target.cuda())) - float(F.mse_loss(pred, target))) < 1e-6
    assert abs(float(o.ssim(pred.cuda(), target.cuda())) - float(R.ssim(p01, t01, 1.0))) < 2e-6
    ref_c = R.color_preservation_loss(pred, target)
    assert abs(float(P.color_preservation_loss(pred.cuda(), target.cuda())) - float(ref_c)) < 2e-6
    assert abs(float(P.color_preservation_loss(pred.cuda(), target.cuda(), include_ssim=False)) - float(R.color_l1(pred, target))) < 1e-6


def test_time_embed_and_linear_rows(ops):
    sd = {"time_embed.proj.0.weight": torch.randn(1024, 256, generator=g(1)) / 16, "time_embed.proj.0.bias": torch.randn(1024, generator=g(2)) * 0.1,
          "time_embed.proj.2.weight": torch.randn(256, 1024, generator=g(3)) / 32, "time_embed.proj.2.bias": torch.randn(256, generator=g(4)) * 0.1}
    t = torch.tensor([0.0, 0.37, 0.8, 0.9875])
    ref = R.time_embedding(sd, t)
    out = ops.time_embed(t.cuda(), *(sd[k].cuda() for k in sd)).cpu()
    assert rel(out, ref) < 1e-5
    x = torch.randn(5, 96, generator=g(5)); w = torch.randn(40, 96, generator=g(6)); b = torch.randn(40, generator=g(7))
    out = ops.linear_rows(x.cuda(), w.cuda(), b.cuda(), ops.ACT_SIGMOID).cpu()
    assert rel(out, torch.sigmoid(F.linear(x, w, b))) < 1e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (1, 1024, 2, 2), (2, 128, 1, 1), (3, 256, 8, 24)])
def test_groupnorm(ops, dtype, tol, shape):
    B, C, H, W = shape
    x = torch.randn(shape, generator=g(3)) * 2 + 0.7
    if dtype == torch.bfloat16:
        x = bf16_round(x)
    gamma, beta = torch.randn(C, generator=g(4)), torch.randn(C, generator=g(5))
    xd = nhwc(x, dtype)
    st = ops.groupnorm_stats(xd, 8)
    for act, fn in ((ops.ACT_NONE, lambda v: v), (ops.ACT_GELU, F.gelu), (ops.ACT_SILU, F.silu)):
        ref = fn(F.group_norm(x, 8, gamma, beta, 1e-5))
        out = nchw(ops.groupnorm_apply(xd, st, gamma.cuda(), beta.cuda(), act))
        assert rel(out, ref) < tol


def test_groupnorm_nchw_and_conv_input(ops):
    x = torch.randn(2, 3, 32, 32, generator=g(1)) * 0.6 + 0.1
    gamma, beta = torch.randn(3, generator=g(2)), torch.randn(3, generator=g(3))
    w, b = torch.randn(64, 3, 3, 3, generator=g(4)) / 5, torch.randn(64, generator=g(5))
    tb = torch.randn(2, 64, generator=g(6))
    st = ops.groupnorm_stats(x.cuda(), 3, nchw=True)
    ref = F.conv2d(F.group_norm(x, 3, gamma, beta, 1e-5), w, b, padding=1) + tb[:, :, None, None]
    out = nchw(ops.conv_input(x.cuda(), w.cuda(), b.cuda(), torch.float32, st, gamma.cuda(), beta.cuda(), row_bias=tb.cuda()))
    assert rel(out, ref) < 1e-5
    w1 = torch.randn(64, 3, 1, 1, generator=g(7))
    out = nchw(ops.conv_input(x.cuda(), w1.cuda(), b.cuda(), torch.bfloat16))
    assert rel(out, F.conv2d(x, w1, b)) < 5e-3


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("shape", [(2, 64, 128, 16, 16), (1, 128, 32, 8, 8), (3, 32, 64, 5, 7), (1, 1024, 512, 2, 2), (2, 16, 192, 1, 1)])
def test_conv3x3_generic(ops, dtype, tol, shape):
    B, Ci, Co, H, W = shape
    x = torch.randn(B, Ci, H, W, generator=g(1))
    w = torch.randn(Co, Ci, 3, 3, generator=g(2)) / math.sqrt(9 * Ci)
    b = torch.randn(Co, generator=g(3))
    tb = torch.randn(B, Co, generator=g(4))
    res = torch.randn(B, Co, H, W, generator=g(5))
    if dtype == torch.bfloat16:
        x, w, res = bf16_round(x), bf16_round(w), bf16_round(res)
    ref = F.conv2d(x, w, b, padding=1) + tb[:, :, None, None] + res
    out = nchw(ops.conv3x3(nhwc(x, dtype), pack3(w, dtype), Co, ops.IMPL_SIMT, bias=b.cuda(), row_bias=tb.cuda(), res=nhwc(res, dtype)))
    assert rel(out, ref) < tol
    scale = torch.rand(B, generator=g(6)) + 0.5
    ref = torch.sigmoid(F.conv2d(x, w, b, padding=1)) * scale.view(-1, 1, 1, 1)
    out = nchw(ops.conv3x3(nhwc(x, dtype), pack3(w, dtype), Co, ops.IMPL_SIMT, bias=b.cuda(), act=ops.ACT_SIGMOID, img_scale=scale.cuda()))
    assert rel(out, ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("fam,hw", [("webp", (16, 16)), ("jpeg", (16, 24)), ("webp", (2, 2)), ("jpeg", (4, 4)), ("webp", (1, 1))])
def test_dct_freq_block_composite(ops, dtype, tol, fam, hw):
    """block_transform + stacked gate GEMMs (freq_mode 1/2) + conv_out == WebP/JPEGFreqAwareBlock.forward."""
    C, B = 64, 2
    H, W = hw
    f = R.FAMILY[fam]
    gen = g(11)
    rnd = lambda *s: torch.randn(*s, generator=gen)
    sd = {"p.dct.dct_matrix": R.dct_matrix(f["bs"])}
    for gname in ("low_freq_attn", "high_freq_attn"):
        sd[f"p.{gname}.0.weight"] = rnd(C // 2, C, 1, 1) / 8; sd[f"p.{gname}.0.bias"] = rnd(C // 2) * 0.3
        sd[f"p.{gname}.2.weight"] = rnd(C, C // 2, 1, 1) / 5; sd[f"p.{gname}.2.bias"] = rnd(C) * 0.3
    sd["p.conv_out.weight"] = rnd(C, C, 3, 3) / 24; sd["p.conv_out.bias"] = rnd(C) * 0.1
    x = rnd(B, C, H, W)
    level = torch.tensor([0.2, 0.95])
    if dtype == torch.bfloat16:
        x = bf16_round(x)
        sd = {k: (bf16_round(v) if k.endswith("weight") else v) for k, v in sd.items()}
    ref = R.dct_freq_block(sd, "p", x, level, f)
    xd = nhwc(x, dtype)
    d = ops.block_transform(xd, sd["p.dct.dct_matrix"].cuda(), 0.0, 1.0)
    assert rel(nchw(d), R.block_transform(x, sd["p.dct.dct_matrix"])) < (1e-6 if dtype == torch.float32 else 4e-3)
    w1 = torch.cat([sd["p.low_freq_attn.0.weight"], sd["p.high_freq_attn.0.weight"]], 0).reshape(C, C)
    b1 = torch.cat([sd["p.low_freq_attn.0.bias"], sd["p.high_freq_attn.0.bias"]], 0)
    w2 = torch.cat([sd["p.low_freq_attn.2.weight"].reshape(C, C // 2), sd["p.high_freq_attn.2.weight"].reshape(C, C // 2)], 1)
    boost = torch.clamp(1.0 - level, f["clamp"][0], f["clamp"][1])
    g1 = ops.gemm(d, pack1(w1, dtype), C, ops.IMPL_SIMT, bias=b1.cuda(), act=ops.ACT_LRELU02, freq_mode=1, bs=f["bs"], low=f["low"])
    e = ops.gemm(g1, pack1(w2, dtype), C, ops.IMPL_SIMT, bias=sd["p.low_freq_attn.2.bias"].cuda(), bias2=sd["p.high_freq_attn.2.bias"].cuda(),
                 act=ops.ACT_SIGMOID, freq_mode=2, bs=f["bs"], low=f["low"], img_scale=boost.cuda(), mul=d, res=xd)
    out = nchw(ops.conv3x3(e, pack3(sd["p.conv_out.weight"], dtype), C, ops.IMPL_SIMT, bias=sd["p.conv_out.bias"].cuda()))
    assert rel(out, ref) < tol


def test_mixed_dtype_stream_and_operands(ops):
    """bf16 operands with an fp32 result stream: fp32 out + bf16 operand copy + fp32 residual; GroupNorm and the block
    transform reading the fp32 stream and writing bf16 operands."""
    B, Ci, Co, H, W = 2, 64, 128, 8, 8
    x = bf16_round(torch.randn(B, Ci, H, W, generator=g(1)))
    w = bf16_round(torch.randn(Co, Ci, 3, 3, generator=g(2)) / 24)
    b = torch.randn(Co, generator=g(3))
    res = torch.randn(B, Co, H, W, generator=g(4))          # fp32 residual, NOT rounded
    ref = F.conv2d(x, w, b, padding=1) + res
    out, out2 = ops.conv3x3(nhwc(x, torch.bfloat16), pack3(w, torch.bfloat16), Co, ops.IMPL_SIMT, out_dtype=torch.float32,
                            out2_dtype=torch.bfloat16, bias=b.cuda(), res=nhwc(res))
    assert out.dtype == torch.float32 and out2.dtype == torch.bfloat16
    assert rel(nchw(out), ref) < 1e-5
    assert torch.equal(out2, out.to(torch.bfloat16))
    # GroupNorm: fp32 stream in, bf16 operand out + raw operand copy
    xs = torch.randn(B, Ci, H, W, generator=g(5)) * 3 + 1
    gamma, beta = torch.randn(Ci, generator=g(6)), torch.randn(Ci, generator=g(7))
    st = ops.groupnorm_stats(nhwc(xs), 8)
    y, raw = ops.groupnorm_apply(nhwc(xs), st, gamma.cuda(), beta.cuda(), ops.ACT_GELU, out_dtype=torch.bfloat16, raw_copy=True)
    assert y.dtype == torch.bfloat16 and torch.equal(raw, nhwc(xs).to(torch.bfloat16))
    assert rel(nchw(y), F.gelu(F.group_norm(xs, 8, gamma, beta, 1e-5))) < 4e-3
    d = ops.block_transform(nhwc(xs), R.dct_matrix(4).cuda(), 0.0, 1.0, out_dtype=torch.bfloat16)
    assert d.dtype == torch.bfloat16 and rel(nchw(d), R.block_transform(xs, R.dct_matrix(4))) < 4e-3


@pytest.mark.parametrize("per_channel,bs,hw", [(0, 4, (10, 7)), (0, 8, (16, 16)), (1, 8, (8, 24)), (1, 8, (3, 5)), (0, 8, (1, 1))])
def test_block_transform(ops, per_channel, bs, hw):
    C = 96
    x = torch.randn(2, C, *hw, generator=g(1))
    T = torch.randn(C, bs, bs, generator=g(2)) if per_channel else R.dct_matrix(bs)
    ref = 0.3 * x + 1.7 * R.block_transform(x, T)
    out = nchw(ops.block_transform(nhwc(x), T.cuda(), 0.3, 1.7))
    assert rel(out, ref) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_upsample_concat(ops, dtype):
    x = torch.randn(2, 64, 12, 20, generator=g(1))
    skip = torch.randn(2, 32, 24, 40, generator=g(2))
    if dtype == torch.bfloat16:
        x, skip = bf16_round(x), bf16_round(skip)
    assert torch.equal(nchw(ops.maxpool2(nhwc(x, dtype))), F.max_pool2d(x, 2))
    ref = torch.cat([F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), skip], 1)
    out = nchw(ops.upsample2_concat(nhwc(x, dtype), nhwc(skip, dtype)))
    assert rel(out, ref) < (1e-6 if dtype == torch.float32 else 4e-3)
    one = torch.randn(1, 16, 1, 1, generator=g(3)); sk = torch.randn(1, 8, 2, 2, generator=g(4))
    ref = torch.cat([F.interpolate(one, scale_factor=2, mode="bilinear", align_corners=False), sk], 1)
    assert rel(nchw(ops.upsample2_concat(nhwc(one), nhwc(sk))), ref) < 1e-6


@pytest.mark.parametrize("hw", [(16, 16), (8, 8), (4, 4), (2, 2), (1, 1), (24, 40)])
def test_avif_pyramid_and_combine(ops, hw):
    B, C = 2, 64
    H, W = hw
    gen = g(5)
    h = torch.randn(B, C, H, W, generator=gen)
    xt, color, edge = (torch.randn(B, C, H, W, generator=gen) for _ in range(3))
    pooled = ops.avgpool_pyramid(nhwc(h)).cpu()      # [85, B, C]
    off = 0
    gates = torch.rand(85, B, C, generator=gen)
    acc = 0
    for s in (1, 2, 4, 8):
        ref = F.adaptive_avg_pool2d(h, s)            # [B, C, s, s]
        got = pooled[off:off + s * s].view(s, s, B, C).permute(2, 3, 0, 1)
        assert rel(got, ref) < 1e-6
        gmap = gates[off:off + s * s].view(s, s, B, C).permute(2, 3, 0, 1).contiguous()
        if gmap.shape != h.shape:
            gmap = F.interpolate(gmap, size=(H, W), mode="bilinear", align_corners=False)
        acc = acc + gmap
        off += s * s
    ref = h + xt * (acc / 4) * color * edge
    out = nchw(ops.avif_combine(nhwc(h), nhwc(xt), gates.cuda(), nhwc(color), nhwc(edge)))
    assert rel(out, ref) < 1e-6


def test_out_conv_tanh(ops):
    x = torch.randn(2, 64, 16, 24, generator=g(1))
    w, b = torch.randn(3, 64, 3, 3, generator=g(2)) / 24, torch.randn(3, generator=g(3)) * 0.1
    ref = torch.tanh(F.conv2d(x, w, b, padding=1))
    assert rel(ops.out_conv_tanh(nhwc(x), w.cuda(), b.cuda()).cpu(), ref) < 1e-5
    assert rel(ops.out_conv_tanh(nhwc(bf16_round(x), torch.bfloat16), w.cuda(), b.cuda()).cpu(), torch.tanh(F.conv2d(bf16_round(x), w, b, padding=1))) < 1e-5


def _attn_ref(qkv, heads):
    B, L, C3 = qkv.shape
    C = C3 // 3
    q, k, v = qkv.view(B, L, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, C)


@pytest.mark.parametrize("hd,heads,L", [(8, 8, 100), (16, 4, 64), (16, 4, 1), (32, 4, 257), (64, 4, 16), (128, 4, 4), (256, 4, 64)])
def test_attention_simt_fp32(ops, hd, heads, L):
    C = hd * heads
    qkv = torch.randn(2, L, 3 * C, generator=g(hd + L))
    out = ops.attention(qkv.cuda(), heads, ops.IMPL_SIMT).cpu()
    assert rel(out, _attn_ref(qkv, heads)) < 1e-5


@pytest.mark.parametrize("expmode", [-1, 0, 1, 2, 3, 4, 16, 18, 19])
@pytest.mark.parametrize("hd,heads,L", [(8, 8, 1024), (16, 4, 1024), (16, 4, 64), (32, 4, 256), (64, 4, 256), (128, 4, 128), (16, 4, 4096)])
def test_attention_tensor_core_bf16(ops, expmode, hd, heads, L):
    from ddpm_image_restoration_b200 import _lib
    C = hd * heads
    qkv = bf16_round(torch.randn(2, L, 3 * C, generator=g(hd + L)) * 1.5)
    _lib.lib().ddpmir_attention_set_expmode(expmode)
    try:
        out = ops.attention(qkv.to(torch.bfloat16).cuda(), heads, ops.IMPL_TENSOR).float().cpu()
    finally:
        _lib.lib().ddpmir_attention_set_expmode(-1)
    tol = 1.5e-2 if expmode == 1 else 6e-3
    assert rel(out, _attn_ref(qkv, heads)) < tol
    # and the bf16 SIMT kernel agrees too
    out2 = ops.attention(qkv.to(torch.bfloat16).cuda(), heads, ops.IMPL_SIMT).float().cpu()
    assert rel(out2, _attn_ref(qkv, heads)) < 4e-3


def _prescaled(qkv, heads, gain=1.0, dtype=torch.bfloat16):
    """bf16 (or binary16) qkv whose q third carries log2(e)/sqrt(hd) (as the packed in_proj produces it) + the fp32 tensor it
    stands for."""
    C = qkv.shape[-1] // 3
    c = 1.4426950408889634 / math.sqrt(C // heads)
    pre = qkv.clone()
    pre[..., :C] *= c * gain
    pre = pre.to(dtype).float()
    ref = pre.clone()
    ref[..., :C] /= c
    return pre.to(dtype), ref


@pytest.mark.parametrize("sel", [-1, 0, 1, 2, 3, 4, 5, 8, 11, 12, 18, 19, 20, 21, 66, 75])
@pytest.mark.parametrize("hd,heads,L", [(8, 8, 1024), (16, 4, 1024), (16, 8, 64), (8, 8, 4096), (8, 4, 192), (32, 4, 256), (64, 4, 128),
                                        (16, 4, 100), (256, 4, 64), (16, 8, 2048), (8, 4, 1152)])
def test_attention_prescaled_bf16(ops, sel, hd, heads, L):
    """Bounded-softmax kernels (head_dim 8/16: tcgen05/TMEM kernel for L >= 1024 with L % 128 == 0, mma.sync kernel otherwise
    or with +64 in the selector; every exp-pipe split: +8 degree-2 polynomial, +16 packed bf16 pairs) and the exact kernels behind the same entry point."""
    from ddpm_image_restoration_b200 import _lib
    if sel > 0 and hd > 16:
        pytest.skip("the split only exists in the bounded kernel")
    C = hd * heads
    pre, ref = _prescaled(torch.randn(2, L, 3 * C, generator=g(hd + L)) * 1.5, heads)
    _lib.lib().ddpmir_attention_set_expmode(-1 if sel < 0 else ((sel + 1) << 8))
    try:
        out = ops.attention_prescaled(pre.cuda(), heads).float().cpu()
    finally:
        _lib.lib().ddpmir_attention_set_expmode(-1)
    assert rel(out, _attn_ref(ref, heads)) < 6e-3


@pytest.mark.parametrize("L", [512, 2048])      # mma.sync bounded kernel / tcgen05 kernel
@pytest.mark.parametrize("hd,heads", [(8, 8), (16, 4)])
def test_attention_prescaled_large_logits_take_the_exact_kernel(ops, hd, heads, L):
    """Rows whose Cauchy-Schwarz logit bound exceeds the fp32-safe window are declined by the bounded kernel and redone by
    the online-maximum kernel."""
    C = hd * heads
    qkv = torch.randn(2, L, 3 * C, generator=g(77)) * 1.5
    qkv[0, 100:140, :C] *= 40.0      # CTA 0 of image 0: logit bound >> 60 -> exact kernel
    qkv[1, 300:330, :C] *= 4.0       # bound around the limit: some heads bounded, some exact
    qkv[1, :, C:2 * C] *= 1.7
    pre, ref = _prescaled(qkv, heads)
    out = ops.attention_prescaled(pre.cuda(), heads).float().cpu()
    want = _attn_ref(ref, heads)
    assert torch.isfinite(out).all()
    assert rel(out, want) < 6e-3
    assert rel(out[0, 100:140], want[0, 100:140]) < 1e-2
    assert rel(out[1, 300:330], want[1, 300:330]) < 1e-2


@pytest.mark.parametrize("hd,heads", [(8, 8), (16, 4)])
def test_attention_full_resolution_kernels_agree(ops, hd, heads):
    """BASELINE's full size (L = 256*256 tokens): the tcgen05 bounded-softmax kernel, the mma.sync bounded-softmax kernel and the
    exact online-maximum kernel (unscaled q) must agree; softmax rows sum to one, so a constant V must come back unchanged."""
    from ddpm_image_restoration_b200 import _lib
    C, L = hd * heads, 65536
    qkv = torch.randn(1, L, 3 * C, generator=g(3)) * 1.2
    qkv[..., 2 * C:] = 0.75                                   # constant V: out == 0.75 whatever the probabilities are
    pre, ref = _prescaled(qkv, heads)
    out_tc = ops.attention_prescaled(pre.cuda(), heads).float().cpu()
    assert (out_tc - 0.75).abs().max() < 4e-3
    # random V: the three kernels against each other
    qkv = torch.randn(1, L, 3 * C, generator=g(5)) * 1.2
    pre, ref = _prescaled(qkv, heads)
    out_tc = ops.attention_prescaled(pre.cuda(), heads).float()
    _lib.lib().ddpmir_attention_set_expmode((66 + 1) << 8)    # +64: mma.sync bounded kernel
    try:
        out_mma = ops.attention_prescaled(pre.cuda(), heads).float()
    finally:
        _lib.lib().ddpmir_attention_set_expmode(-1)
    out_exact = ops.attention(ref.to(torch.bfloat16).cuda(), heads, ops.IMPL_TENSOR).float()
    assert rel(out_tc.cpu(), out_exact.cpu()) < 6e-3
    assert rel(out_mma.cpu(), out_exact.cpu()) < 6e-3


def _attn_rows_fp64(ref, heads, rows):
    """softmax(q k^T / sqrt(hd)) v in float64 on the CPU for the sampled query rows of image 0: [len(rows), C]."""
    C = ref.shape[-1] // 3
    hd = C // heads
    q, k, v = (ref[0, :, i * C:(i + 1) * C].double().view(-1, heads, hd) for i in range(3))
    out = torch.empty(len(rows), heads, hd, dtype=torch.float64)
    for h in range(heads):
        s = (q[rows, h] @ k[:, h].T) / math.sqrt(hd)
        out[:, h] = torch.softmax(s, dim=-1) @ v[:, h]
    return out.view(len(rows), C)


@pytest.mark.parametrize("split", [0, 2, 4, 5, 7, 3 << 3, 5 << 3])
@pytest.mark.parametrize("scale", [0.3, 0.8])
@pytest.mark.parametrize("hd,heads,L", [(8, 8, 2048), (16, 4, 1024), (8, 4, 1152)])
def test_attention_half_precision_tier(ops, hd, heads, L, scale, split):
    """attn_tc16.cu (S and P in binary16): scale 0.3 -> logit bound < 2, the degree-4 polynomial variant; 0.8 -> bound < 11, the
    range-reduced variant; every MUFU / FMA-pipe split, against the fp32 SDPA reference.  The tolerance is tighter than the
    bf16 tier's: P keeps 11 significand bits."""
    from ddpm_image_restoration_b200 import _lib
    C = hd * heads
    pre, ref = _prescaled(torch.randn(2, L, 3 * C, generator=g(hd + L)) * scale, heads, dtype=torch.float16)
    assert pre.dtype == ops.qkv_dtype_for_attention(L, hd)
    _lib.lib().ddpmir_attention_set_expmode(split << 16)
    _lib.lib().ddpmir_attention_set_lin(-1)                   # the polynomial-kernel tier would take these (image, head) pairs
    try:
        out = ops.attention_prescaled(pre.cuda(), heads).float().cpu()
    finally:
        _lib.lib().ddpmir_attention_set_expmode(-1)
        _lib.lib().ddpmir_attention_set_lin(6)
    want = _attn_ref(ref, heads)
    assert rel(out, want) < 4e-3
    # the bf16 entry point (bf16 tier) on the same values
    out2 = ops.attention_prescaled(pre.to(torch.bfloat16).cuda(), heads).float().cpu()
    assert rel(out2, want) < 8e-3


def test_attention_tiers_mixed_in_one_launch(ops):
    """CTAs of one launch land in different tiers: most in the half-precision tier, one CTA with logits beyond +-11 in the
    bf16 tier, one beyond +-60 in the exact kernel."""
    hd, heads, L = 8, 8, 2048
    C = hd * heads
    qkv = torch.randn(2, L, 3 * C, generator=g(91)) * 0.4
    qkv[0, 256:300, :C] *= 12.0       # CTA 2 of image 0: bound ~ 25 -> bf16 tier
    qkv[1, 1300:1330, :C] *= 80.0     # CTA 10 of image 1: bound > 60 -> exact kernel
    pre, ref = _prescaled(qkv, heads, dtype=torch.float16)
    out = ops.attention_prescaled(pre.cuda(), heads).float().cpu()
    want = _attn_ref(ref, heads)
    # tiles that leave the first tier are recomputed from a bf16 copy of qkv: their reference sees the same rounding
    c = 1.4426950408889634 / math.sqrt(hd)
    ref_bf = pre.to(torch.bfloat16).float()
    ref_bf[..., :C] /= c
    want_bf = _attn_ref(ref_bf, heads)
    assert torch.isfinite(out).all()
    keep = torch.ones(2, L, dtype=torch.bool)
    keep[0, 256:384] = False
    keep[1, 1280:1408] = False
    assert rel(out[keep], want[keep]) < 4e-3
    assert rel(out[0, 256:384], want_bf[0, 256:384]) < 6e-3
    assert rel(out[1, 1280:1408], want_bf[1, 1280:1408]) < 1e-2


def _logit_bound(pre, heads):
    """Per (image, head): the logit bound of the polynomial-kernel tier's pre-pass (attn_lin.cu) -- Cauchy-Schwarz after centring
    q' and k on their means and balancing the two factors per dimension, max_i |D (q'_i - a)| * max_j |(k_j - b) / D|."""
    B, L, C3 = pre.shape
    C = C3 // 3
    q = pre[..., :C].float().view(B, L, heads, -1)
    k = pre[..., C:2 * C].float().view(B, L, heads, -1)
    qc, kc = q - q.mean(1, keepdim=True), k - k.mean(1, keepdim=True)
    D = ((kc * kc).mean(1, keepdim=True) / (qc * qc).mean(1, keepdim=True)).pow(0.25).clamp(1 / 16, 16)
    return (qc * D).norm(dim=-1).amax(1) * (kc / D).norm(dim=-1).amax(1)


@pytest.mark.parametrize("simt", [0, 16])                                      # tcgen05 kernels (attn_lin_tc.cu) / fp32 SIMT kernels (attn_lin.cu)
@pytest.mark.parametrize("target", [0.7, 0.95, 1.2, 1.45, 1.95, 2.45, 3.4])    # logit bound -> polynomial set 0 .. 6 (degree 3, 3, 3, 4, 4, 5, 6)
@pytest.mark.parametrize("hd,heads,L,B", [(8, 8, 2048, 2), (8, 4, 1152, 2), (8, 2, 4096, 2), (8, 1, 16384, 1), (16, 4, 1024, 2), (16, 1, 32768, 1)])
def test_attention_polynomial_kernel_tier(ops, hd, heads, L, B, target, simt):
    """attn_lin*.cu: (image, head) pairs whose logit bound fits a polynomial set are evaluated through the monomial feature map of
    the minimax polynomial of 2^s (two O(L) contractions).  Checked against float64 softmax with a zero-mean V (the output is then
    the small position-dependent part of the attention, nothing hides behind a mean) and against the quadratic half-precision tier."""
    from ddpm_image_restoration_b200 import _lib
    if L >= 16384 and not ((hd == 8 and target > 3.0) or (hd == 16 and 1.3 < target < 2.0)):
        pytest.skip("the long sequences are here for the degrees that need them (head_dim 8: 6, head_dim 16: 4)")
    C = hd * heads
    qkv = torch.randn(B, L, 3 * C, generator=g(hd + L)) * 0.3
    qkv[..., :2 * C] += torch.randn(1, 1, 2 * C, generator=g(7)) * 0.2     # q and k with a common component, as feature maps have
    pre, ref = _prescaled(qkv, heads, dtype=torch.float16)
    gain = target / float(_logit_bound(pre, heads).max())      # scale q so that the largest (image, head) bound hits the target
    pre, ref = _prescaled(qkv, heads, gain=gain, dtype=torch.float16)
    assert float(_logit_bound(pre, heads).max()) < target * 1.02
    _lib.lib().ddpmir_attention_set_lin(6 + simt)
    try:
        out, tiers = ops.attention_prescaled(pre.cuda(), heads, return_tiers=True)
        _lib.lib().ddpmir_attention_set_lin(-1)
        quad = ops.attention_prescaled(pre.cuda(), heads).double().cpu()
    finally:
        _lib.lib().ddpmir_attention_set_lin(6)
    out, tiers = out.double().cpu(), tiers.cpu()
    want = _attn_ref(ref.double(), heads)
    r, rq = rel(out, want), rel(quad, want)
    print(f"polynomial-kernel tier ({'SIMT' if simt else 'tcgen05'}) hd={hd} L={L} bound={target}: sets {sorted(set(tiers.flatten().tolist()))}, "
          f"rel-L2 vs fp64 = {r:.3e} (quadratic tier: {rq:.3e})")
    # the verdicts must be the smallest ALLOWED window that holds each (image, head) bound (attn_lin.cuh::set_mask: head_dim 8 skips
    # set 2; degree 5 from L = 4096, degree 6 from 16384; head_dim 16: degree 4 from L = 32768; the SIMT kernels stop at degree
    # 4 / 3)
    windows = [0.75, 1.0, 1.25, 1.5, 2.0, 2.5, 3.5]
    if hd == 8:
        allowed = [0, 1, 3, 4] + ([5] if (not simt and L >= 4096) else []) + ([6] if (not simt and L >= 16384) else [])
    else:
        allowed = [0, 1, 2] + ([3, 4] if (not simt and L >= 32768) else [])
    bounds = _logit_bound(pre, heads)
    for bnd, t in zip(bounds.flatten().tolist(), tiers.flatten().tolist()):
        ok = [s_ for s_ in allowed if bnd * 0.999 <= windows[s_]]
        near_edge = any(abs(bnd - w) < 0.01 * w for w in windows)
        if not near_edge:
            assert t == (ok[0] if ok else -1), (bnd, t)
    assert r < 6e-3


def test_attention_polynomial_kernel_tier_mixed(ops):
    """One launch, three regimes: most (image, head) pairs through the polynomial kernel, one head with a logit bound of ~6
    (half-precision quadratic tier), one row with a bound > 60 (exact kernel)."""
    hd, heads, L = 8, 8, 2048
    C = hd * heads
    qkv = torch.randn(2, L, 3 * C, generator=g(17)) * 0.3
    qkv[0, :, 3 * hd:4 * hd] *= 4.0        # q of head 3, image 0
    qkv[1, 777, :hd] *= 600.0               # one query row of head 0, image 1
    pre, ref = _prescaled(qkv, heads, dtype=torch.float16)
    b = _logit_bound(pre, heads)
    assert (b <= 2.0).sum() == 2 * heads - 2 and 3.6 < b[0, 3] < 11.0 and b[1, 0] > 60.0
    out, tiers = ops.attention_prescaled(pre.cuda(), heads, return_tiers=True)
    out, tiers = out.float().cpu(), tiers.cpu()
    assert tiers[1, 0] == -1 and tiers[0, 3] == -1 and (tiers >= 0).sum() == 2 * heads - 2   # bound ~5 is beyond the largest window (3.5) too
    want = _attn_ref(ref, heads)
    assert torch.isfinite(out).all()
    assert rel(out, want) < 6e-3
    for (bi, h) in ((0, 3), (1, 0), (0, 0), (1, 7)):
        assert rel(out[bi, :, h * hd:(h + 1) * hd], want[bi, :, h * hd:(h + 1) * hd]) < 1e-2, (bi, h)


@pytest.mark.parametrize("scale", [0.2, 0.35, 0.8, 1.2])     # logit bound ~1 / ~2 (a fresh UNet) / ~9 / ~20: polynomial tier and the three bounded tiers
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("hd,heads", [(8, 8), (16, 4)])
def test_attention_full_resolution_vs_fp64_rows(ops, hd, heads, dtype, scale):
    """L = 65 536 (the bench's full-resolution blocks, head_dim 8 = AVIF, 16 = WebP/JPEG): 384 sampled query rows of the
    production kernels (binary16 qkv = what the UNet feeds at this length: three-tier path; bf16 qkv: bf16 tier) against a
    float64 softmax computed on the CPU -- an independent reference, not a sibling kernel."""
    C, L = hd * heads, 65536
    qkv = torch.randn(1, L, 3 * C, generator=g(11 + hd)) * scale
    qkv[0, :, 2 * C:] += 0.3                                   # V with a mean, as feature maps have
    pre, ref = _prescaled(qkv, heads, dtype=dtype)
    out = ops.attention_prescaled(pre.cuda(), heads).float().cpu()
    rows = torch.randperm(L, generator=g(5))[:384]
    rows[:4] = torch.tensor([0, 127, 128, L - 1])
    want = _attn_rows_fp64(ref, heads, rows)
    r = rel(out[0, rows], want)
    print(f"attention L=65536 hd={hd} {dtype} scale={scale}: rel-L2 of 384 rows vs fp64 softmax = {r:.3e}")
    assert r < 6e-3
    assert (out[0, rows].double() - want).abs().max() < 0.02 * want.abs().max() + 1e-3


def test_cpu_tensor_raises(ops):
    from ddpm_image_restoration_b200._lib import DdpmirError
    with pytest.raises(DdpmirError):
        ops.maxpool2(torch.zeros(1, 2, 2, 8))
