"""Tensor-level wrappers of the backward / optimiser kernels (training step, webp_training.py:476-537).
Same conventions as ops.py: CUDA tensors only, outputs allocated with torch, kernels on torch's current stream."""
import torch

from . import _lib
from .ops import LAUNCHES, _code, _f32, _p, _stream, ACT_NONE  # noqa: F401

F32 = torch.float32


def wgrad(dy, x, out, taps, n_begin=0, n_count=None, k_begin=0, k_count=None, out_ld=None, oihw=False):
    """out += dY^T im2col(X) (sub-block).  dy [B,H,W,N], x [B,H,W,Cin]; out fp32 (packed [n_count, k_count] or OIHW)."""
    B, H, W, N = dy.shape
    Cin = x.shape[-1]
    K = taps * Cin
    n_count = N if n_count is None else n_count
    k_count = K if k_count is None else k_count
    out_ld = k_count if out_ld is None else out_ld
    _lib.check(_lib.lib().ddpmir_wgrad(_p(dy), _code(dy.dtype), _p(x), _code(x.dtype), _p(_f32(out, "out")), B, H, W, Cin, N, taps,
                                       n_begin, n_count, k_begin, k_count, out_ld, int(oihw), _stream()), "wgrad")
    LAUNCHES[0] += 1


def colsum(dy, out_total=None, out_img=None, cls=-1, bs=0, low=0, n_begin=0, n_count=None):
    B, H, W, N = dy.shape
    n_count = N - n_begin if n_count is None else n_count
    _lib.check(_lib.lib().ddpmir_colsum(_p(dy), _code(dy.dtype), B, H, W, N, cls, bs, low, n_begin, n_count, _p(out_total),
                                        _p(out_img), _stream()), "colsum")
    LAUNCHES[0] += 1


def groupnorm_backward(x, dy, mean_rstd, gamma, beta, act, dgamma, dbeta, dx=None, accumulate=False):
    B, H, W, C = x.shape
    G = mean_rstd.shape[1]
    if dx is None:
        dx = torch.empty_like(x)
        accumulate = False
    ws = torch.empty((B * G * 2,), dtype=torch.float64, device=x.device)
    _lib.check(_lib.lib().ddpmir_groupnorm_backward(_p(_f32(x, "x")), _p(_f32(dy, "dy")), B, H * W, C, G, act, _p(mean_rstd), _p(gamma),
                                                    _p(beta), _p(dx), int(accumulate), _p(dgamma), _p(dbeta), _p(ws), _stream()),
               "groupnorm_backward")
    LAUNCHES[0] += 2
    return dx


def gate_backward(de, g, d, boost, bs, low):
    B, H, W, C = de.shape
    dz = torch.empty_like(de)
    dd = torch.empty_like(de)
    _lib.check(_lib.lib().ddpmir_gate_backward(_p(_f32(de, "de")), _p(g), _p(d), _code(g.dtype), _p(boost), _p(dz), _p(dd), B, H, W, C,
                                               bs, low, _stream()), "gate_backward")
    LAUNCHES[0] += 1
    return dz, dd


def lrelu_mask_backward(dg1, g1, bs, low):
    B, H, W, N = dg1.shape
    dpre = torch.empty_like(dg1)
    _lib.check(_lib.lib().ddpmir_lrelu_mask_backward(_p(_f32(dg1, "dg1")), _p(g1), _code(g1.dtype), _p(dpre), B, H, W, N, bs, low,
                                                     _stream()), "lrelu_mask_backward")
    LAUNCHES[0] += 1
    return dpre


def avif_combine_backward(de, xt, gates, color, edge, boost_color, boost_edge):
    """Product rule of e = h + xt * A * color * edge (ops.avif_combine): -> dxt, dz_color, dz_edge (pre-sigmoid), dattn."""
    B, H, W, C = de.shape
    outs = [torch.empty_like(de) for _ in range(4)]
    _lib.check(_lib.lib().ddpmir_avif_combine_backward(_p(_f32(de, "de")), _p(xt), _p(_f32(gates, "gates")), _p(color), _p(edge),
                                                       _code(xt.dtype), _p(boost_color), _p(boost_edge), B, H, W, C,
                                                       *[_p(o) for o in outs], _stream()), "avif_combine_backward")
    LAUNCHES[0] += 1
    return outs


def avif_gates_backward(dattn):
    """Transposed bilinear up-sampling of the gate pyramid: [B,H,W,C] -> [85, B, C]."""
    B, H, W, C = dattn.shape
    dg = torch.empty((85, B, C), dtype=F32, device=dattn.device)
    _lib.check(_lib.lib().ddpmir_avif_gates_backward(_p(_f32(dattn, "dattn")), B, H, W, C, _p(dg), _stream()), "avif_gates_backward")
    LAUNCHES[0] += 1
    return dg


def avgpool_pyramid_backward(dpooled, dx):
    """dx += transposed adaptive average pooling of dpooled [85, B, C]."""
    B, H, W, C = dx.shape
    _lib.check(_lib.lib().ddpmir_avgpool_pyramid_backward(_p(_f32(dpooled, "dpooled")), B, H, W, C, _p(_f32(dx, "dx")), 1, _stream()),
               "avgpool_pyramid_backward")
    LAUNCHES[0] += 1
    return dx


def block_transform_wgrad(x, dz, Tm, dT):
    """dT += weight gradient of the learned per-channel 8x8 transform Z = T X T^T."""
    B, H, W, C = x.shape
    _lib.check(_lib.lib().ddpmir_block_transform_wgrad(_p(_f32(x, "x")), _p(_f32(dz, "dz")), _p(_f32(Tm, "T")), Tm.shape[-1], B, H, W, C,
                                                       _p(_f32(dT, "dT")), _stream()), "block_transform_wgrad")
    LAUNCHES[0] += 1


def relu_mask_backward(dy, y):
    dpre = torch.empty_like(dy)
    _lib.check(_lib.lib().ddpmir_relu_mask_backward(_p(_f32(dy, "dy")), _p(y), _code(y.dtype), _p(dpre), dy.numel(), _stream()),
               "relu_mask_backward")
    LAUNCHES[0] += 1
    return dpre


def dropout(x, p, seed, out_dtype=None):
    out = torch.empty(x.shape, dtype=x.dtype if out_dtype is None else out_dtype, device=x.device)
    _lib.check(_lib.lib().ddpmir_dropout(_p(x), _code(x.dtype), _p(out), _code(out.dtype), x.numel(), float(p), int(seed), _stream()),
               "dropout")
    LAUNCHES[0] += 1
    return out


def maxpool2_backward(x, dy):
    B, H, W, C = x.shape
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().ddpmir_maxpool2_backward(_p(_f32(x, "x")), _p(_f32(dy, "dy")), _p(dx), B, H, W, C, _stream()), "maxpool2_backward")
    LAUNCHES[0] += 1
    return dx


def upsample2_concat_backward(dy, C1):
    B, Ho, Wo, Ct = dy.shape
    H, W, C2 = Ho // 2, Wo // 2, Ct - C1
    dlo = torch.empty((B, H, W, C1), dtype=F32, device=dy.device)
    dskip = torch.empty((B, Ho, Wo, C2), dtype=F32, device=dy.device)
    _lib.check(_lib.lib().ddpmir_upsample2_concat_backward(_p(_f32(dy, "dy")), _p(dlo), _p(dskip), B, H, W, C1, C2, _stream()),
               "upsample2_concat_backward")
    LAUNCHES[0] += 1
    return dlo, dskip


def attention_train_forward(qkv, heads):
    B, L, C3 = qkv.shape
    C = C3 // 3
    out = torch.empty((B, L, C), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B, heads, L), dtype=F32, device=qkv.device)
    _lib.check(_lib.lib().ddpmir_attention_train_forward(_p(qkv), _code(qkv.dtype), B, L, C, heads, _p(out), _p(lse), _stream()),
               "attention_train_forward")
    LAUNCHES[0] += 1
    return out, lse


def attention_backward(qkv, o, dout, lse, heads, dout_op=None):
    """dout_op: optional bf16 copy of dout (the tensor-core kernel's operand); made here when the caller has none."""
    B, L, C3 = qkv.shape
    C = C3 // 3
    dqkv = torch.empty((B, L, C3), dtype=F32, device=qkv.device)
    delta = torch.empty((B, heads, L), dtype=F32, device=qkv.device)
    if not (qkv.dtype == torch.bfloat16 and L % 64 == 0 and (C // heads) <= 64):
        dout_op = None
    elif dout_op is None:
        from .ops import cast_bf16
        dout_op = cast_bf16(dout.contiguous())
    elif dout_op.dtype != torch.bfloat16 or dout_op.numel() != dout.numel():
        raise _lib.DdpmirError("attention_backward: dout_op must be the bf16 copy of dout")
    _lib.check(_lib.lib().ddpmir_attention_backward(_p(qkv), _p(o), _code(qkv.dtype), _p(_f32(dout, "dout")), _p(dout_op), _p(lse),
                                                    _p(delta), _p(dqkv), B, L, C, heads, _stream()), "attention_backward")
    LAUNCHES[0] += 3
    return dqkv


def time_features(t, dim=256):
    out = torch.empty((t.shape[0], dim), dtype=F32, device=t.device)
    _lib.check(_lib.lib().ddpmir_time_features(_p(_f32(t, "t")), t.shape[0], dim, _p(out), _stream()), "time_features")
    LAUNCHES[0] += 1
    return out


def act_forward(x, act, out=None):
    if out is None:
        out = torch.empty_like(x)
    elif out.numel() != x.numel() or out.dtype != F32 or not out.is_contiguous():
        raise _lib.DdpmirError("act_forward: bad out tensor")
    _lib.check(_lib.lib().ddpmir_act_forward(_p(_f32(x, "x")), act, _p(out), x.numel(), _stream()), "act_forward")
    LAUNCHES[0] += 1
    return out


def act_backward(dy, u, act):
    dx = torch.empty_like(dy)
    _lib.check(_lib.lib().ddpmir_act_backward(_p(_f32(dy, "dy")), _p(_f32(u, "u")), act, _p(dx), dy.numel(), _stream()), "act_backward")
    LAUNCHES[0] += 1
    return dx


def linear_rows_backward(dy, x, w, dw=None, db=None, need_dx=True, dx_accum=None):
    """dx = dy W (or dx_accum += dy W), dw += dy^T x, db += colsum(dy)."""
    rows, N = dy.shape
    K = x.shape[1]
    dx = dx_accum if dx_accum is not None else (torch.empty((rows, K), dtype=F32, device=dy.device) if need_dx else None)
    _lib.check(_lib.lib().ddpmir_linear_rows_backward(_p(_f32(dy, "dy")), _p(_f32(x, "x")), _p(_f32(w, "w")), rows, K, N, _p(dx),
                                                      int(dx_accum is not None), _p(dw), _p(db), _stream()), "linear_rows_backward")
    LAUNCHES[0] += 2
    return dx


def conv_input_backward(x, dh, w, dw, mean_rstd=None, gamma=None, beta=None, dgamma=None, dbeta=None):
    B, Cin, H, W = x.shape
    N, ks = w.shape[0], w.shape[-1]
    _lib.check(_lib.lib().ddpmir_conv_input_backward(_p(_f32(x, "x")), _p(_f32(dh, "dh")), B, Cin, H, W, N, ks, _p(_f32(w, "w")),
                                                     _p(mean_rstd), _p(gamma), _p(beta), _p(dw), _p(dgamma), _p(dbeta), _stream()),
               "conv_input_backward")
    LAUNCHES[0] += 2 if mean_rstd is not None else 1


def out_conv_tanh_backward(a, y, dy, w, dw, dbias):
    B, H, W, Cin = a.shape
    N = w.shape[0]
    da = torch.empty((B, H, W, Cin), dtype=F32, device=a.device)
    _lib.check(_lib.lib().ddpmir_out_conv_tanh_backward(_p(a), _code(a.dtype), _p(_f32(y, "y")), _p(_f32(dy, "dy")), B, H, W, Cin, N,
                                                        _p(_f32(w, "w")), _p(da), _p(dw), _p(dbias), _stream()), "out_conv_tanh_backward")
    LAUNCHES[0] += 2
    return da


def frequency_aware_loss_backward(pred, target, upstream=1.0):
    """d frequency_aware_loss(pred, target) / d pred  (webp_training.py:105-132), scaled by `upstream`."""
    B, C, H, W = pred.shape
    dpred = torch.empty_like(pred)
    lib = _lib.lib()
    _lib.check(lib.ddpmir_mse_backward(_p(_f32(pred, "pred")), _p(_f32(target, "target")), pred.numel(), float(upstream), _p(dpred), 0,
                                       _stream()), "mse_backward")
    count = float(B * H * (W // 2 + 1))
    wp = torch.empty((B * C, H, W, 2), dtype=F32, device=pred.device)
    wt = torch.empty_like(wp)
    wg = torch.empty_like(wp)
    # 0.5 * sum_c [mse_mag + 0.5 * mse_phase], each mse a mean over `count` coefficients
    _lib.check(lib.ddpmir_freq_loss_backward(_p(pred), _p(target), B * C, H, W, upstream * 0.5 / count, upstream * 0.25 / count, _p(wp),
                                             _p(wt), _p(wg), _p(dpred), _stream()), "freq_loss_backward")
    ws = torch.empty((B * C * 3 * (H - 10) * (W - 10),), dtype=F32, device=pred.device)
    _lib.check(lib.ddpmir_ssim_backward(_p(pred), _p(target), B * C, H, W, 0, -0.3 * upstream, _p(dpred), _p(ws), _stream()),
               "ssim_backward")
    LAUNCHES[0] += 7
    return dpred


def avif_frequency_aware_loss_backward(pred, target, upstream=1.0):
    """d avif_frequency_aware_loss(pred, target) / d pred  (avif.py:126-164), scaled by `upstream`."""
    B, C, H, W = pred.shape
    dpred = torch.empty_like(pred)
    lib = _lib.lib()
    _lib.check(lib.ddpmir_mse_backward(_p(_f32(pred, "pred")), _p(_f32(target, "target")), pred.numel(), float(upstream), _p(dpred), 0,
                                       _stream()), "mse_backward")
    count = float(B * H * W)
    wp = torch.empty((B * C, H, W, 2), dtype=F32, device=pred.device)
    wt = torch.empty_like(wp)
    wg = torch.empty_like(wp)
    # 0.3 * sum_c [mse_mag + 0.3 * mse_phase], each mse a mean over `count` coefficients of the full spectrum
    _lib.check(lib.ddpmir_fft2_loss_backward(_p(pred), _p(target), B * C, H, W, upstream * 0.3 / count, upstream * 0.09 / count, _p(wp),
                                             _p(wt), _p(wg), _p(dpred), _stream()), "fft2_loss_backward")
    ws = torch.empty((B * C * 3 * (H - 10) * (W - 10),), dtype=F32, device=pred.device)
    _lib.check(lib.ddpmir_ssim_backward(_p(pred), _p(target), B * C, H, W, 0, -0.4 * upstream, _p(dpred), _p(ws), _stream()),
               "ssim_backward")
    _lib.check(lib.ddpmir_edge_loss_backward(_p(pred), _p(target), B * C, H, W, 0.2 * upstream / float(B * C * (H - 1) * W),
                                             0.2 * upstream / float(B * C * H * (W - 1)), _p(dpred), 1, _stream()), "edge_loss_backward")
    LAUNCHES[0] += 8
    return dpred


def color_preservation_loss_backward(pred, target, upstream=1.0, include_ssim=True, out=None):
    """d color_preservation_loss(pred, target) / d pred (0409_method.ipynb#c0:L64-82: colour L1 + 0.5 (1 - SSIM) on images
    clamped to [0,1]), scaled by `upstream`; accumulated into `out` when given."""
    B, C, H, W = pred.shape
    if C != 3:
        raise _lib.DdpmirError("color_preservation_loss needs 3-channel images")
    lib = _lib.lib()
    dpred = out if out is not None else torch.empty_like(pred)
    _lib.check(lib.ddpmir_color_l1_backward(_p(_f32(pred, "pred")), _p(_f32(target, "target")), B, H, W, float(upstream), _p(dpred),
                                            int(out is not None), _stream()), "color_l1_backward")
    LAUNCHES[0] += 1
    if include_ssim:
        ws = torch.empty((B * C * 3 * (H - 10) * (W - 10),), dtype=F32, device=pred.device)
        _lib.check(lib.ddpmir_ssim_backward(_p(pred), _p(target), B * C, H, W, 1, -0.5 * upstream, _p(dpred), _p(ws), _stream()),
                   "ssim_backward")
        LAUNCHES[0] += 2
    return dpred


def huber_backward(a, b, delta=1.0, upstream=1.0, out=None):
    """d HuberLoss(mean, delta)(a, b) / d a, scaled by `upstream`; accumulated into `out` when given."""
    da = out if out is not None else torch.empty_like(a)
    _lib.check(_lib.lib().ddpmir_huber_backward(_p(_f32(a, "a")), _p(_f32(b, "b")), a.numel(), float(delta), float(upstream), _p(da),
                                                int(out is not None), _stream()), "huber_backward")
    LAUNCHES[0] += 1
    return da


def huber_color_loss_backward(pred_noise, noise, xt, x0, color_weight, delta=1.0):
    """Gradient wrt pred_noise of the 0409 notebook's training loss (#c0:L567-574):
    huber(pred_noise, noise) + color_weight * color_preservation_loss(xt + pred_noise, x0)."""
    from .ops import lincomb
    d = huber_backward(pred_noise, noise, delta)
    recon = lincomb(xt, 1.0, pred_noise, 1.0)
    return color_preservation_loss_backward(recon, x0, upstream=color_weight, out=d)


def pack_weight(w, fwd=None, bwd=None, fwd_ld=None, fwd_off=0, bwd_ld=None, bwd_off=0):
    """One pass over a checkpoint-layout weight [N, Cin, kh, kw] / [N, K] (fp32) -> the forward operand [N, (kh,kw,cin)] and the
    data-gradient operand [Cin, (kh,kw,cout)] (transposed, taps flipped) in the dtype of the outputs; leading dimensions /
    offsets place it inside a stacked operand."""
    N, Cin = w.shape[0], w.shape[1]
    taps = w.numel() // (N * Cin)
    ref = fwd if fwd is not None else bwd
    if bwd is not None and fwd is not None and bwd.dtype != fwd.dtype:
        raise _lib.DdpmirError("pack_weight: fwd and bwd must share a dtype")
    _lib.check(_lib.lib().ddpmir_pack_weight(_p(_f32(w, "w")), N, Cin, taps, _p(fwd), taps * Cin if fwd_ld is None else fwd_ld, fwd_off,
                                             _p(bwd), taps * N if bwd_ld is None else bwd_ld, bwd_off, _code(ref.dtype), _stream()),
               "pack_weight")
    LAUNCHES[0] += 1


def sumsq(x, acc):
    _lib.check(_lib.lib().ddpmir_sumsq(_p(_f32(x, "x")), x.numel(), _p(acc), _stream()), "sumsq")
    LAUNCHES[0] += 1


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_sumsq=None, max_norm=1.0):
    _lib.check(_lib.lib().ddpmir_adamw_step(_p(p), _p(_f32(g, "g")), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2),
                                            float(eps), float(weight_decay), int(step), _p(grad_sumsq), float(max_norm), _stream()),
               "adamw_step")
    LAUNCHES[0] += 1


def adamw_multi(chunk_start, chunk_len, chunk_param, g_flat, m_flat, v_flat, lr, beta1, beta2, eps, weight_decay, step,
                grad_sumsq=None, max_norm=1.0):
    """clip + AdamW over every tensor of the chunk table in one launch (tables: device int64 / int32 / int64-as-pointer)."""
    _lib.check(_lib.lib().ddpmir_adamw_multi(_p(chunk_start), _p(chunk_len), _p(chunk_param), chunk_start.numel(), _p(_f32(g_flat, "g")),
                                             _p(m_flat), _p(v_flat), float(lr), float(beta1), float(beta2), float(eps),
                                             float(weight_decay), int(step), _p(grad_sumsq), float(max_norm), _stream()), "adamw_multi")
    LAUNCHES[0] += 1
