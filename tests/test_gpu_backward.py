"""Backward kernels of the training step against torch.autograd on the CPU (fp32), one test per kernel."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import restated as R
from util import bf16_round, nchw, nhwc, pack1, pack3, rel

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def ops():
    from ddpm_image_restoration_b200 import ops as o
    return o


@pytest.fixture(scope="module")
def T():
    from ddpm_image_restoration_b200 import ops_train as t
    return t


def g(seed):
    return torch.Generator().manual_seed(seed)


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=g(seed)) * scale


@pytest.mark.parametrize("shape", [(2, 64, 128, 8, 8, 9), (1, 128, 64, 4, 6, 9), (2, 64, 192, 16, 16, 1), (3, 32, 48, 5, 5, 9)])
def test_wgrad_colsum_and_dgrad(ops, T, shape):
    B, Ci, Co, H, W, taps = shape
    ks = 3 if taps == 9 else 1
    x = rnd(B, Ci, H, W, seed=1).requires_grad_()
    w = (rnd(Co, Ci, ks, ks, seed=2) / math.sqrt(taps * Ci)).requires_grad_()
    b = rnd(Co, seed=3).requires_grad_()
    dy = rnd(B, Co, H, W, seed=4)
    F.conv2d(x, w, b, padding=ks // 2).backward(dy)
    dw = torch.zeros(Co, Ci, ks, ks).cuda()
    T.wgrad(nhwc(dy), nhwc(x.detach()), dw, taps, oihw=(taps == 9))
    assert rel(dw.cpu(), w.grad) < 1e-5
    # sub-block into a packed buffer (rows 16.., second half of K)
    K = taps * Ci
    sub = torch.zeros(Co - 16, K // 2).cuda()
    T.wgrad(nhwc(dy), nhwc(x.detach()), sub, taps, n_begin=16, n_count=Co - 16, k_begin=K // 2, k_count=K // 2)
    packed = w.grad.permute(0, 2, 3, 1).reshape(Co, K)
    assert rel(sub.cpu(), packed[16:, K // 2:]) < 1e-5
    db = torch.zeros(Co).cuda(); dimg = torch.zeros(B, Co).cuda()
    T.colsum(nhwc(dy), db, dimg)
    assert rel(db.cpu(), b.grad) < 1e-5 and rel(dimg.cpu(), dy.sum((2, 3))) < 1e-5
    # data gradient = the forward kernel on flipped / transposed weights
    wt = w.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, taps * Co).contiguous().cuda()
    fn = ops.conv3x3 if taps == 9 else ops.gemm
    if Co % 16 == 0:
        dx = fn(nhwc(dy), wt, Ci, ops.IMPL_SIMT)
        assert rel(nchw(dx), x.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 64, 128, 16, 16, 9), (1, 128, 64, 8, 8, 9), (2, 64, 192, 16, 16, 1), (3, 256, 256, 4, 4, 9),
                                   (2, 128, 128, 8, 8, 1), (1, 1024, 512, 2, 2, 9)])
def test_wgrad_tensor_core(T, shape):
    """bf16 operands -> mma.sync kernel; against autograd on the same bf16-rounded operands, incl. the gate sub-blocks."""
    B, Ci, Co, H, W, taps = shape
    ks = 3 if taps == 9 else 1
    x = bf16_round(rnd(B, Ci, H, W, seed=1))
    dy = bf16_round(rnd(B, Co, H, W, seed=4))
    w = torch.zeros(Co, Ci, ks, ks, requires_grad=True)
    F.conv2d(x, w, None, padding=ks // 2).backward(dy)
    BF = torch.bfloat16
    dw = torch.zeros(Co, Ci, ks, ks).cuda()
    T.wgrad(nhwc(dy, BF), nhwc(x, BF), dw, taps, oihw=(taps == 9))
    assert rel(dw.cpu(), w.grad) < 2e-5
    if taps == 1:
        lo = torch.zeros(Co, Ci // 2).cuda(); hi = torch.zeros(Co // 2, Ci).cuda()
        T.wgrad(nhwc(dy, BF), nhwc(x, BF), lo, 1, k_begin=Ci // 2, k_count=Ci // 2, out_ld=Ci // 2)
        T.wgrad(nhwc(dy, BF), nhwc(x, BF), hi, 1, n_begin=Co // 2, n_count=Co // 2)
        g2 = w.grad.reshape(Co, Ci)
        assert rel(lo.cpu(), g2[:, Ci // 2:]) < 2e-5 and rel(hi.cpu(), g2[Co // 2:]) < 2e-5


def test_colsum_class_split(T):
    B, N, H, W = 2, 32, 8, 12
    dy = rnd(B, N, H, W, seed=1)
    m = R.low_mask(H, W, 4, 3).float()
    lo = torch.zeros(N).cuda(); hi = torch.zeros(N).cuda()
    T.colsum(nhwc(dy), lo, None, cls=1, bs=4, low=3)
    T.colsum(nhwc(dy), hi, None, cls=0, bs=4, low=3)
    assert rel(lo.cpu(), (dy * m).sum((0, 2, 3))) < 1e-5 and rel(hi.cpu(), (dy * (1 - m)).sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("act,fn", [(0, lambda v: v), (5, F.gelu), (4, F.silu)])
def test_groupnorm_backward(ops, T, act, fn):
    B, C, H, W = 2, 64, 6, 10
    x = (rnd(B, C, H, W, seed=1) * 2 + 0.5).requires_grad_()
    gamma, beta = rnd(C, seed=2).requires_grad_(), rnd(C, seed=3).requires_grad_()
    dy = rnd(B, C, H, W, seed=4)
    fn(F.group_norm(x, 8, gamma, beta, 1e-5)).backward(dy)
    st = ops.groupnorm_stats(nhwc(x.detach()), 8)
    dgamma, dbeta = torch.zeros(C).cuda(), torch.zeros(C).cuda()
    dx = T.groupnorm_backward(nhwc(x.detach()), nhwc(dy), st, gamma.detach().cuda(), beta.detach().cuda(), act, dgamma, dbeta)
    assert rel(nchw(dx), x.grad) < 2e-5
    assert rel(dgamma.cpu(), gamma.grad) < 2e-5 and rel(dbeta.cpu(), beta.grad) < 2e-5


def test_gate_and_lrelu_mask_backward(T):
    B, C, H, W, bs, low = 2, 64, 8, 8, 4, 3
    m = R.low_mask(H, W, bs, low).float()
    h3 = rnd(B, C, H, W, seed=1)
    d = rnd(B, C, H, W, seed=2).requires_grad_()
    z = rnd(B, C, H, W, seed=3).requires_grad_()
    boost = torch.tensor([0.7, 0.2])
    s = torch.where(m.bool(), torch.ones(1), boost.view(B, 1, 1, 1))
    e = h3 + torch.sigmoid(z) * s * d
    de = rnd(B, C, H, W, seed=4)
    e.backward(de)
    dz, dd = T.gate_backward(nhwc(de), nhwc(torch.sigmoid(z.detach())), nhwc(d.detach()), boost.cuda(), bs, low)
    assert rel(nchw(dz), z.grad) < 1e-5 and rel(nchw(dd), d.grad) < 1e-5
    pre = rnd(B, C, H, W, seed=5).requires_grad_()
    hid = F.leaky_relu(pre, 0.2)
    g1 = torch.cat([hid[:, :C // 2] * m, hid[:, C // 2:] * (1 - m)], 1)
    dg1 = rnd(B, C, H, W, seed=6)
    g1.backward(dg1)
    dpre = T.lrelu_mask_backward(nhwc(dg1), nhwc(g1.detach()), bs, low)
    assert rel(nchw(dpre), pre.grad) < 1e-6


def test_dropout_is_its_own_backward(T):
    x = rnd(4, 8, 8, 64, seed=1).cuda()
    y = T.dropout(x, 0.1, seed=5)
    keep = (y != 0)
    assert abs(float(keep.float().mean()) - 0.9) < 0.02
    assert torch.allclose(y[keep], x[keep] / 0.9)
    assert torch.equal(T.dropout(x, 0.1, seed=5), y) and not torch.equal(T.dropout(x, 0.1, seed=6), y)
    assert torch.equal(T.dropout(x, 0.0, seed=5), x)


def test_pool_and_upsample_backward(T):
    x = rnd(2, 16, 8, 12, seed=1).requires_grad_()
    dy = rnd(2, 16, 4, 6, seed=2)
    F.max_pool2d(x, 2).backward(dy)
    assert torch.equal(nchw(T.maxpool2_backward(nhwc(x.detach()), nhwc(dy))), x.grad)
    for hw in ((4, 6), (1, 1), (2, 1)):
        lo = rnd(2, 16, *hw, seed=3).requires_grad_()
        skip = rnd(2, 8, 2 * hw[0], 2 * hw[1], seed=4).requires_grad_()
        out = torch.cat([F.interpolate(lo, scale_factor=2, mode="bilinear", align_corners=False), skip], 1)
        dy = rnd(*out.shape, seed=5)
        out.backward(dy)
        dlo, dskip = T.upsample2_concat_backward(nhwc(dy), 16)
        assert rel(nchw(dlo), lo.grad) < 1e-6 and torch.equal(nchw(dskip), skip.grad)


@pytest.mark.parametrize("hd,heads,L", [(16, 4, 64), (8, 8, 100), (32, 4, 33), (64, 2, 16), (128, 2, 8), (256, 2, 4)])
def test_attention_backward(T, hd, heads, L):
    C = hd * heads
    qkv = rnd(2, L, 3 * C, seed=hd + L).requires_grad_()
    q, k, v = qkv.view(2, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, L, C)
    do = rnd(2, L, C, seed=7)
    o.backward(do)
    out, lse = T.attention_train_forward(qkv.detach().cuda(), heads)
    assert rel(out.cpu(), o.detach()) < 1e-5
    dqkv = T.attention_backward(qkv.detach().cuda(), out, do.cuda(), lse, heads)
    assert rel(dqkv.cpu(), qkv.grad) < 2e-5


@pytest.mark.parametrize("hd,heads,L", [(16, 4, 256), (8, 8, 128), (32, 4, 64), (64, 4, 192), (16, 4, 1024), (128, 2, 64), (16, 2, 96)])
def test_attention_train_forward_bf16_lse(T, hd, heads, L):
    """bf16 training forward (tensor-core kernel) returns the log-sum-exp the backward needs; bf16 backward is consistent."""
    C = hd * heads
    qkv = bf16_round(rnd(2, L, 3 * C, seed=hd + L)).requires_grad_()
    q, k, v = qkv.view(2, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(2, L, C)
    do = rnd(2, L, C, seed=7)
    o.backward(do)
    out, lse = T.attention_train_forward(qkv.detach().to(torch.bfloat16).cuda(), heads)
    assert rel(lse.cpu(), torch.logsumexp(s.detach(), -1)) < 2e-3
    assert rel(out.float().cpu(), o.detach()) < 6e-3
    dqkv = T.attention_backward(qkv.detach().to(torch.bfloat16).cuda(), out, do.cuda(), lse, heads)
    assert rel(dqkv.cpu(), qkv.grad) < 1e-2


def test_small_layers_backward(ops, T):
    # row-wise linear + SiLU (time embedding MLP)
    x = rnd(4, 256, seed=1).requires_grad_()
    w = (rnd(1024, 256, seed=2) / 16).requires_grad_(); b = rnd(1024, seed=3).requires_grad_()
    u = F.linear(x, w, b)
    y = F.silu(u)
    dy = rnd(4, 1024, seed=4)
    y.backward(dy)
    du = T.act_backward(dy.cuda(), u.detach().cuda(), ops.ACT_SILU)
    dw, db = torch.zeros(1024, 256).cuda(), torch.zeros(1024).cuda()
    dx = T.linear_rows_backward(du, x.detach().cuda(), w.detach().cuda(), dw, db)
    assert rel(dx.cpu(), x.grad) < 1e-5 and rel(dw.cpu(), w.grad) < 1e-5 and rel(db.cpu(), b.grad) < 1e-5
    assert rel(T.act_forward(u.detach().cuda(), ops.ACT_SILU).cpu(), y.detach()) < 1e-6
    # input conv with folded GroupNorm
    img = (rnd(2, 3, 16, 16, seed=5) * 0.6).requires_grad_(False)
    gamma, beta = rnd(3, seed=6).requires_grad_(), rnd(3, seed=7).requires_grad_()
    wc = (rnd(64, 3, 3, 3, seed=8) / 5).requires_grad_()
    h = F.conv2d(F.group_norm(img, 3, gamma, beta, 1e-5), wc, None, padding=1)
    dh = rnd(2, 64, 16, 16, seed=9)
    h.backward(dh)
    st = ops.groupnorm_stats(img.cuda(), 3, nchw=True)
    dwc, dga, dbe = torch.zeros(64, 3, 3, 3).cuda(), torch.zeros(3).cuda(), torch.zeros(3).cuda()
    T.conv_input_backward(img.cuda(), nhwc(dh), wc.detach().cuda(), dwc, st, gamma.detach().cuda(), beta.detach().cuda(), dga, dbe)
    assert rel(dwc.cpu(), wc.grad) < 1e-5 and rel(dga.cpu(), gamma.grad) < 1e-5 and rel(dbe.cpu(), beta.grad) < 1e-5
    w1 = rnd(64, 3, 1, 1, seed=10).requires_grad_()
    F.conv2d(img, w1).backward(dh)
    dw1 = torch.zeros(64, 3, 1, 1).cuda()
    T.conv_input_backward(img.cuda(), nhwc(dh), w1.detach().cuda(), dw1)
    assert rel(dw1.cpu(), w1.grad) < 1e-5
    # out_conv + tanh
    a = rnd(2, 64, 12, 12, seed=11).requires_grad_()
    wo = (rnd(3, 64, 3, 3, seed=12) / 24).requires_grad_(); bo = rnd(3, seed=13).requires_grad_()
    yv = torch.tanh(F.conv2d(a, wo, bo, padding=1))
    dyo = rnd(2, 3, 12, 12, seed=14)
    yv.backward(dyo)
    dwo, dbo = torch.zeros(3, 64, 3, 3).cuda(), torch.zeros(3).cuda()
    da = T.out_conv_tanh_backward(nhwc(a.detach()), yv.detach().cuda(), dyo.cuda(), wo.detach().cuda(), dwo, dbo)
    assert rel(nchw(da), a.grad) < 1e-5 and rel(dwo.cpu(), wo.grad) < 1e-5 and rel(dbo.cpu(), bo.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 16, 32, 32), (1, 64, 64, 64), (2, 8, 4, 4), (1, 16, 2, 2), (2, 8, 1, 1), (1, 8, 12, 20), (1, 8, 8, 8)])
def test_avif_gate_pyramid_backward(ops, T, shape):
    """e = h + xt * mean_s(up(gate_s)) * color * edge (avif.py:293-321): the product rule, the transposed bilinear up-sampling
    (incl. s > H, s == H and ragged ratios) and the transposed adaptive pooling, each against torch.autograd."""
    B, C, H, W = shape
    h = rnd(B, C, H, W, seed=1)
    xt = rnd(B, C, H, W, seed=2).requires_grad_()
    gate_maps = [torch.sigmoid(rnd(B, C, s, s, seed=10 + s)).requires_grad_() for s in (1, 2, 4, 8)]
    bc, be = torch.tensor([0.9, 1.3][:B]), torch.tensor([0.8, 1.1][:B])
    zc, ze = rnd(B, C, H, W, seed=3).requires_grad_(), rnd(B, C, H, W, seed=4).requires_grad_()
    color = torch.sigmoid(zc) * bc.view(-1, 1, 1, 1)
    edge = torch.sigmoid(ze) * be.view(-1, 1, 1, 1)
    acc = 0
    for q in gate_maps:
        acc = acc + (q if q.shape[-2:] == (H, W) else F.interpolate(q, size=(H, W), mode="bilinear", align_corners=False))
    e = h + xt * (acc / 4) * color * edge
    de = rnd(B, C, H, W, seed=5)
    e.backward(de)
    gates = torch.cat([q.detach().permute(2, 3, 0, 1).reshape(-1, B, C) for q in gate_maps], 0).contiguous().cuda()
    fwd = ops.avif_combine(nhwc(h), nhwc(xt.detach()), gates, nhwc(color.detach()), nhwc(edge.detach()))
    assert rel(nchw(fwd), e.detach()) < 1e-6
    dxt, dzc, dze, dattn = T.avif_combine_backward(nhwc(de), nhwc(xt.detach()), gates, nhwc(color.detach()), nhwc(edge.detach()),
                                                   bc.cuda(), be.cuda())
    assert rel(nchw(dxt), xt.grad) < 1e-5 and rel(nchw(dzc), zc.grad) < 1e-5 and rel(nchw(dze), ze.grad) < 1e-5
    dg = T.avif_gates_backward(dattn).cpu()
    want = torch.cat([q.grad.permute(2, 3, 0, 1).reshape(-1, B, C) for q in gate_maps], 0)
    assert rel(dg, want) < 1e-5
    # adaptive pooling pyramid, forward and transposed
    x = rnd(B, C, H, W, seed=6).requires_grad_()
    pooled = torch.cat([F.adaptive_avg_pool2d(x, s).permute(2, 3, 0, 1).reshape(-1, B, C) for s in (1, 2, 4, 8)], 0)
    dp = rnd(85, B, C, seed=7)
    pooled.backward(dp)
    assert rel(ops.avgpool_pyramid(nhwc(x.detach())).cpu(), pooled.detach()) < 1e-5
    base = rnd(B, C, H, W, seed=8)
    dx = T.avgpool_pyramid_backward(dp.cuda(), nhwc(base))
    assert rel(nchw(dx) - base, x.grad) < 1e-5


@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (1, 8, 8, 8), (3, 72, 12, 20), (1, 128, 4, 4)])
def test_learned_transform_weight_gradient(ops, T, shape):
    """Z = T_c X T_c^T per channel and 8x8 block (AVIFAdaptiveTransform, avif.py:205-232): dT and dX against autograd over the
    oracle's block_transform (ragged sizes pad with zeros and crop)."""
    B, C, H, W = shape
    x = rnd(B, C, H, W, seed=1).requires_grad_()
    Tw = (rnd(C, 8, 8, seed=2) * 0.4).requires_grad_()
    z = R.block_transform(x, Tw)
    dz = rnd(B, C, H, W, seed=3)
    z.backward(dz)
    dT = torch.zeros(C, 8, 8).cuda()
    T.block_transform_wgrad(nhwc(x.detach()), nhwc(dz), Tw.detach().cuda(), dT)
    assert rel(dT.cpu(), Tw.grad) < 1e-5
    dx = ops.block_transform(nhwc(dz), Tw.detach().transpose(1, 2).contiguous().cuda(), 0.0, 1.0)
    assert rel(nchw(dx), x.grad) < 1e-5
    y = torch.relu(rnd(2, 8, 8, 16, seed=4))
    dy = rnd(2, 8, 8, 16, seed=5)
    for dt in (torch.float32, torch.bfloat16):
        assert torch.equal(T.relu_mask_backward(dy.cuda(), y.to(dt).cuda()).cpu(), dy * (y > 0))


@pytest.mark.parametrize("shape", [(64, 64, 9), (128, 64, 9), (48, 40, 9), (192, 64, 1), (7, 33, 1), (1024, 512, 9)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
def test_pack_weight(T, shape, dt):
    """ddpmir_pack_weight against the torch expressions it replaces: [N,(kh,kw,cin)] and the transposed, tap-flipped
    data-gradient operand; bit-exact (a cast and a permutation), incl. placement inside a stacked operand."""
    N, Cin, taps = shape
    w = rnd(N, Cin, 3, 3, seed=1) if taps == 9 else rnd(N, Cin, 1, 1, seed=1)
    want_f = w.permute(0, 2, 3, 1).reshape(N, -1).to(dt)
    want_b = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, -1).to(dt)
    fwd = torch.full((N, taps * Cin), 7.0, dtype=dt).cuda()
    bwd = torch.full((Cin, taps * N), 7.0, dtype=dt).cuda()
    T.pack_weight(w.cuda(), fwd, bwd)
    assert torch.equal(fwd.cpu(), want_f) and torch.equal(bwd.cpu(), want_b)
    if taps == 1:       # two tensors stacked along rows (forward) / columns (transposed), as the low/high gate MLP does
        w2 = rnd(N, Cin, 1, 1, seed=2)
        st, stt = torch.zeros(2 * N, Cin, dtype=dt).cuda(), torch.zeros(Cin, 2 * N, dtype=dt).cuda()
        T.pack_weight(w.cuda(), st[:N], stt, bwd_ld=2 * N, bwd_off=0)
        T.pack_weight(w2.cuda(), st[N:], stt, bwd_ld=2 * N, bwd_off=N)
        cat = torch.cat([w, w2], 0).reshape(2 * N, Cin)
        assert torch.equal(st.cpu(), cat.to(dt)) and torch.equal(stt.cpu(), cat.t().contiguous().to(dt))
        sk, skt = torch.zeros(N, 2 * Cin, dtype=dt).cuda(), torch.zeros(2 * Cin, N, dtype=dt).cuda()
        T.pack_weight(w.cuda(), sk, skt[:Cin], fwd_ld=2 * Cin, fwd_off=0)
        T.pack_weight(w2.cuda(), sk, skt[Cin:], fwd_ld=2 * Cin, fwd_off=Cin)
        catk = torch.cat([w.reshape(N, Cin), w2.reshape(N, Cin)], 1)
        assert torch.equal(sk.cpu(), catk.to(dt)) and torch.equal(skt.cpu(), catk.t().contiguous().to(dt))


def test_frequency_aware_loss_backward(T):
    gen = g(3)
    target = torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1
    pred = (target + 0.2 * torch.randn(2, 3, 32, 32, generator=gen)).requires_grad_()
    p01, t01 = pred * 0.5 + 0.5, target * 0.5 + 0.5
    freq = 0
    for c in range(3):
        fp, ft = torch.fft.rfft2(p01[:, c]), torch.fft.rfft2(t01[:, c])
        freq = freq + F.mse_loss(fp.abs(), ft.abs()) + 0.5 * F.mse_loss(torch.angle(fp), torch.angle(ft))
    loss = F.mse_loss(pred, target) + 0.5 * freq + 0.3 * (1 - R.ssim(p01, t01, 1.0))
    loss.backward()
    got = T.frequency_aware_loss_backward(pred.detach().cuda(), target.cuda()).cpu()
    assert rel(got, pred.grad) < 2e-3      # the phase term's gradient ~ 1/|P| is ill-conditioned at small coefficients


def test_avif_frequency_aware_loss_forward_and_backward(T):
    """avif.py:126-164 (full fft2 spectrum, gradient-edge term, weights 0.3 / 0.4 / 0.2): value and gradient against torch.autograd
    over the oracle's restatement (pinned against the verbatim reference in test_oracle_vs_reference.py)."""
    import ddpm_image_restoration_b200 as P
    gen = g(5)
    for hw in ((32, 32), (64, 32)):
        target = torch.rand(2, 3, *hw, generator=gen) * 2 - 1
        pred = (target + 0.2 * torch.randn(2, 3, *hw, generator=gen)).requires_grad_()
        loss = R.avif_frequency_aware_loss(pred, target)
        loss.backward()
        got = float(P.avif_frequency_aware_loss(pred.detach().cuda(), target.cuda()))
        assert abs(got - float(loss)) < 1e-3 * abs(float(loss)), (got, float(loss))
        grad = T.avif_frequency_aware_loss_backward(pred.detach().cuda(), target.cuda()).cpu()
        assert rel(grad, pred.grad) < 2e-3      # the phase term's gradient ~ 1/|P| is ill-conditioned at small coefficients
        grad2 = T.avif_frequency_aware_loss_backward(pred.detach().cuda(), target.cuda(), upstream=0.5).cpu()
        assert rel(grad2, 0.5 * pred.grad) < 2e-3


def test_adamw_and_clip(T):
    torch.manual_seed(0)
    ps = [torch.randn(1000), torch.randn(37, 5)]
    gs = [torch.randn(1000) * 3, torch.randn(37, 5)]
    ref = [p.clone().requires_grad_() for p in ps]
    opt = torch.optim.AdamW(ref, lr=2e-4, weight_decay=1e-5, betas=(0.9, 0.99))
    dev = [p.clone().cuda() for p in ps]
    ms = [torch.zeros_like(p) for p in dev]; vs = [torch.zeros_like(p) for p in dev]
    for step in (1, 2, 3):
        for r, gr in zip(ref, gs):
            r.grad = gr.clone() * step
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        acc = torch.zeros(1, dtype=torch.float64).cuda()
        gd = [(gr * step).cuda() for gr in gs]
        for gg in gd:
            T.sumsq(gg, acc)
        for p, gg, m, v in zip(dev, gd, ms, vs):
            T.adamw_step(p, gg, m, v, 2e-4, 0.9, 0.99, 1e-8, 1e-5, step, acc, 1.0)
    for p, r in zip(dev, ref):
        assert rel(p.cpu(), r.detach()) < 1e-6
