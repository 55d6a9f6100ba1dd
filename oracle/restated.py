"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the *oracle*) of the reference's restoration hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker or the timed CPU baseline.  The product package never imports it.

Everything here is plain PyTorch fp32 on the CPU, written functionally over a flat ``state_dict`` (the
reference's own checkpoint keys) instead of the reference's nn.Module tree.  Two deliberate restatements
(SURVEY.md section 8(c)); both are pinned against the unmodified reference in tests/test_oracle_vs_reference.py
and through the committed fixtures in tests/golden/ (made by oracle/make_golden.py from the reference itself):

  * self-attention: the reference's nn.MultiheadAttention materialises the [B*heads, L, L] score tensor
    (webp_inference.py:295,317-321) and cannot run at 256x256.  We apply in_proj -> scaled-dot-product
    attention (query-chunked) -> out_proj on the same parameters.
  * low/high frequency split: the reference's Python double loop over blocks (webp_inference.py:236-252,
    svd.ipynb#c1:L148-164) equals a static mask  (i % bs < low) & (j % bs < low)  -- including ragged edges,
    where low_size = max(1, min(low, rows_left, cols_left)).

Third-party arithmetic with no pinned version in the reference (parity unpinned):
  * pytorch_msssim.ssim  (call sites webp_training.py:129, 0409_method.ipynb#c0:L79) -- restated in `ssim`.
Pillow's JPEG/WebP/AVIF codecs are called by both the oracle and the product (same installed library).
"""
from __future__ import annotations

import io
import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------------------------------------
# family constants
# ----------------------------------------------------------------------------------------------------------
FAMILY = {
    # heads: nn.MultiheadAttention(out_c, heads)  webp_inference.py:295 / svd.ipynb#c1:L207 / avif_inference.py:281
    # bs/low/clamp: WebPFreqAwareBlock webp_inference.py:208,247,263 ; JPEGFreqAwareBlock svd.ipynb#c1:L120,159,175
    # tail: `u5 + 0.1*dct_layer(u5)` webp_inference.py:396-397 ; `u5 + 0.15*avif_layer(u5)` avif_inference.py:382-383
    "webp": dict(heads=4, bs=4, low=3, clamp=(0.15, 1.9), tail=0.1),
    "jpeg": dict(heads=4, bs=8, low=4, clamp=(0.2, 2.0), tail=0.1),
    "avif": dict(heads=8, bs=8, tail=0.15),
}

BLOCKS = [  # (prefix, in_c, out_c)  webp_inference.py:337-356
    ("down1", 3, 64), ("down2", 64, 128), ("down3", 128, 256), ("down4", 256, 512), ("down5", 512, 512),
    ("bottleneck.0", 512, 1024), ("bottleneck.1", 1024, 1024), ("bottleneck.2", 1024, 512),
    ("up1", 1024, 512), ("up2", 1024, 256), ("up3", 512, 128), ("up4", 256, 64), ("up5", 128, 64),
]


def gn_groups(c: int) -> int:
    """num_groups rule of webp_inference.py:277-279."""
    g = min(8, c)
    while c % g != 0 and g > 1:
        g -= 1
    return g


# ----------------------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------------------
def dct_matrix(n: int) -> torch.Tensor:
    """Orthonormal DCT-II matrix, webp_inference.py:194-203 (computed in fp32 like the reference)."""
    d = torch.zeros(n, n)
    for i in range(n):
        for j in range(n):
            if i == 0:
                d[i, j] = 1.0 / torch.sqrt(torch.tensor(float(n)))
            else:
                d[i, j] = torch.sqrt(torch.tensor(2.0 / n)) * torch.cos(torch.tensor(math.pi * (2 * j + 1) * i / (2 * n)))
    return d


def time_embedding(sd: SD, t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """webp_inference.py:145-151."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half) * -k)
    e = t[:, None].float() * freqs[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = F.linear(e, sd["time_embed.proj.0.weight"], sd["time_embed.proj.0.bias"])
    e = F.silu(e)
    return F.linear(e, sd["time_embed.proj.2.weight"], sd["time_embed.proj.2.bias"])


def block_transform(x: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """Per-channel blockwise  T_c X T_c^T  with zero padding to a block multiple and crop back.

    DCTLayer.forward webp_inference.py:161-192 (T shared, [bs,bs]) and AVIFAdaptiveTransform.forward
    avif_inference.py:140-177 (T per channel, [C,bs,bs]).
    """
    b, c, h, w = x.shape
    bs = T.shape[-1]
    hp, wp = (-h) % bs, (-w) % bs
    xp = F.pad(x, (0, wp, 0, hp))
    H, W = h + hp, w + wp
    blk = xp.view(b, c, H // bs, bs, W // bs, bs)  # [b,c,I,r,J,s]
    if T.dim() == 2:
        out = torch.einsum("ur,bcirjs,vs->bciujv", T, blk, T)
    else:
        out = torch.einsum("cur,bcirjs,cvs->bciujv", T, blk, T)
    return out.reshape(b, c, H, W)[:, :, :h, :w]


def low_mask(h: int, w: int, bs: int, low: int) -> torch.Tensor:
    """Static restatement of the block loop webp_inference.py:241-252 (ragged edges included)."""
    ii = torch.arange(h)
    jj = torch.arange(w)
    # rows/cols left in the (possibly ragged) block that holds i / j
    rows_left = torch.clamp(h - (ii // bs) * bs, max=bs)
    cols_left = torch.clamp(w - (jj // bs) * bs, max=bs)
    ls = torch.clamp(torch.minimum(rows_left[:, None], cols_left[None, :]), min=1, max=low)
    return ((ii % bs)[:, None] < ls) & ((jj % bs)[None, :] < ls)


def _gate(sd: SD, p: str, x: torch.Tensor, act) -> torch.Tensor:
    h = F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=sd[p + ".0.weight"].shape[-1] // 2)
    h = act(h)
    h = F.conv2d(h, sd[p + ".2.weight"], sd[p + ".2.bias"], padding=sd[p + ".2.weight"].shape[-1] // 2)
    return torch.sigmoid(h)


def dct_freq_block(sd: SD, p: str, x: torch.Tensor, level: Optional[torch.Tensor], fam: dict) -> torch.Tensor:
    """WebPFreqAwareBlock.forward webp_inference.py:231-270 / JPEGFreqAwareBlock svd.ipynb#c1:L143-182."""
    d = block_transform(x, sd[p + ".dct.dct_matrix"])
    m = low_mask(x.shape[2], x.shape[3], fam["bs"], fam["low"]).to(d.dtype)
    lowf, highf = d * m, d * (1 - m)
    lrelu = lambda z: F.leaky_relu(z, 0.2)
    la = _gate(sd, p + ".low_freq_attn", lowf, lrelu)
    ha = _gate(sd, p + ".high_freq_attn", highf, lrelu)
    if level is not None:
        ha = ha * torch.clamp(1.0 - level.view(-1, 1, 1, 1), fam["clamp"][0], fam["clamp"][1])
    return F.conv2d(x + la * lowf + ha * highf, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], padding=1)


def avif_transform(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AVIFAdaptiveTransform.forward avif_inference.py:140-182 (`inverse_weights` is never used)."""
    tr = block_transform(x, sd[p + ".transform_weights"])
    return tr * _gate(sd, p + ".quantization", tr, F.relu)


def avif_freq_block(sd: SD, p: str, x: torch.Tensor, level: Optional[torch.Tensor]) -> torch.Tensor:
    """AVIFFreqAwareBlock.forward avif_inference.py:222-256."""
    xt = avif_transform(sd, p + ".adaptive_transform", x)
    acc = 0
    for idx, s in enumerate((1, 2, 4, 8)):
        q = F.adaptive_avg_pool2d(x, s)
        q = F.conv2d(q, sd[f"{p}.multi_scale_attn.{idx}.1.weight"], sd[f"{p}.multi_scale_attn.{idx}.1.bias"])
        q = F.relu(q)
        q = F.conv2d(q, sd[f"{p}.multi_scale_attn.{idx}.3.weight"], sd[f"{p}.multi_scale_attn.{idx}.3.bias"])
        q = torch.sigmoid(q)
        if q.shape != x.shape:
            q = F.interpolate(q, size=x.shape[-2:], mode="bilinear", align_corners=False)
        acc = acc + q
    attn = acc / 4
    color = _gate(sd, p + ".color_consistency", x, F.relu)
    edge = _gate(sd, p + ".edge_preserve", x, F.relu)
    if level is not None:
        lv = level.view(-1, 1, 1, 1)
        color = color * torch.clamp(0.5 + 0.5 * (1.0 - lv), 0.3, 1.5)
        edge = edge * torch.clamp(0.7 + 0.3 * (1.0 - lv), 0.5, 1.3)
    return F.conv2d(x + xt * attn * color * edge, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], padding=1)


def self_attention(sd: SD, p: str, h: torch.Tensor, heads: int, q_chunk: int = 4096) -> torch.Tensor:
    """`h_attn, _ = self.attn(h_flat, h_flat, h_flat)` webp_inference.py:317-320, memory-tiled."""
    b, c, H, W = h.shape
    L = H * W
    seq = h.flatten(2).transpose(1, 2)  # [B, L, C]
    qkv = F.linear(seq, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = qkv.view(b, L, 3, heads, c // heads).permute(2, 0, 3, 1, 4)  # each [B, heads, L, hd]
    outs = []
    for s in range(0, L, q_chunk):
        outs.append(F.scaled_dot_product_attention(q[:, :, s:s + q_chunk], k, v))
    o = torch.cat(outs, dim=2).transpose(1, 2).reshape(b, L, c)
    o = F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])
    return o.transpose(1, 2).reshape(b, c, H, W)


def res_attn_block(sd: SD, p: str, x: torch.Tensor, t_emb: torch.Tensor, level, family: str) -> torch.Tensor:
    """{WebP,JPEG,AVIF}ResAttnBlock.forward webp_inference.py:303-327 (eval mode: dropout is a no-op)."""
    fam = FAMILY[family]
    in_c = x.shape[1]
    out_c = sd[p + ".conv1.weight"].shape[0]
    h = F.group_norm(x, gn_groups(in_c), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps=1e-5)
    h = F.conv2d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + F.linear(t_emb, sd[p + ".time_proj.weight"], sd[p + ".time_proj.bias"])[..., None, None]
    h = F.group_norm(h, gn_groups(out_c), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps=1e-5)
    h = F.gelu(h)
    h = F.conv2d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    h = h + self_attention(sd, p + ".attn", h, fam["heads"])
    if family == "avif":
        h = avif_freq_block(sd, p + ".freq_guide", h, level)
    else:
        h = dct_freq_block(sd, p + ".freq_guide", h, level, fam)
    if p + ".shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"])
    return x + h


def _up_cat(a: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
    return torch.cat([F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=False), skip], dim=1)


def unet_forward(sd: SD, x: torch.Tensor, t: torch.Tensor, level: Optional[torch.Tensor] = None,
                 family: str = "webp", taps: Optional[dict] = None, enable_grad: bool = False) -> torch.Tensor:
    """{WebP,JPEG,AVIF}DiffusionModel.forward webp_inference.py:369-399 / avif_inference.py:355-385.

    `taps` (optional dict) receives the output of every block, for per-layer parity checks.
    """
    fam = FAMILY[family]
    with (torch.enable_grad() if enable_grad else torch.no_grad()):
        t_emb = time_embedding(sd, t)
        if level is None:
            level = t.clone()
        blk = lambda p, z: res_attn_block(sd, p, z, t_emb, level, family)
        pool = lambda z: F.max_pool2d(z, 2)
        d1 = blk("down1", x)
        d2 = blk("down2", pool(d1))
        d3 = blk("down3", pool(d2))
        d4 = blk("down4", pool(d3))
        d5 = blk("down5", pool(d4))
        bn = blk("bottleneck.0", pool(d5))
        bn = blk("bottleneck.1", bn)
        bn = blk("bottleneck.2", bn)
        u1 = blk("up1", _up_cat(bn, d5))
        u2 = blk("up2", _up_cat(u1, d4))
        u3 = blk("up3", _up_cat(u2, d3))
        u4 = blk("up4", _up_cat(u3, d2))
        u5 = blk("up5", _up_cat(u4, d1))
        if taps is not None:
            taps.update(t_emb=t_emb, d1=d1, d2=d2, d3=d3, d4=d4, d5=d5, bn=bn, u1=u1, u2=u2, u3=u3, u4=u4, u5=u5)
        if family == "avif":
            comb = u5 + fam["tail"] * avif_transform(sd, "avif_layer", u5)
        else:
            comb = u5 + fam["tail"] * block_transform(u5, sd["dct_layer.dct_matrix"])
        o = F.group_norm(comb, 8, sd["out_conv.0.weight"], sd["out_conv.0.bias"], eps=1e-5)
        o = F.silu(o)
        o = F.conv2d(o, sd["out_conv.2.weight"], sd["out_conv.2.bias"], padding=1)
        return torch.tanh(o)


# ----------------------------------------------------------------------------------------------------------
# 0409 UNet variant: experiments/code/0409_method.ipynb cell 0, the model the SVD/GMM solver was
# written for.  Same 13-block skeleton; blocks flagged below carry FrequencyAwareBlock + HFCM, the others none.
# ----------------------------------------------------------------------------------------------------------
FREQ_BLOCKS_0409 = ("down2", "down3", "bottleneck.0", "bottleneck.2", "up2", "up3")     # use_freq_guide=True, L378-397


def freq_aware_block_0409(sd: SD, p: str, x: torch.Tensor, level: torch.Tensor) -> torch.Tensor:
    """FrequencyAwareBlock.forward 0409_method.ipynb#c0:L234-263: x + conv3x3(DCT8(x)) * (SE-gate * (1 - level) + 0.5)."""
    xf = F.conv2d(block_transform(x, dct_matrix(8)), sd[p + ".freq_conv.weight"], sd[p + ".freq_conv.bias"], padding=1)
    a = xf.mean(dim=(2, 3), keepdim=True)
    a = F.relu(F.conv2d(a, sd[p + ".freq_attn.1.weight"], sd[p + ".freq_attn.1.bias"]))
    a = torch.sigmoid(F.conv2d(a, sd[p + ".freq_attn.3.weight"], sd[p + ".freq_attn.3.bias"]))
    a = a * (1.0 - level.view(-1, 1, 1, 1)) + 0.5
    return x + xf * a


def hfcm_0409(sd: SD, p: str, x: torch.Tensor, level: torch.Tensor) -> torch.Tensor:
    """HFCM.forward 0409_method.ipynb#c0:L197-218: conv1x1(x + sigmoid(conv3(relu(conv3(x)))) * DCT8(x) * (1 - level))."""
    mask = _gate(sd, p + ".high_freq_attn", x, F.relu)
    enhanced = x + mask * block_transform(x, dct_matrix(8)) * (1.0 - level.view(-1, 1, 1, 1))
    return F.conv2d(enhanced, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"])


def res_attn_block_0409(sd: SD, p: str, x: torch.Tensor, t_emb: torch.Tensor, level) -> torch.Tensor:
    """ResAttnBlock.forward 0409_method.ipynb#c0:L296-318 (eval mode).  Unlike the shipped families the attention output
    REPLACES h (no residual around it), the activation is SiLU and there is no trailing 3x3 conv."""
    in_c, out_c = x.shape[1], sd[p + ".conv1.weight"].shape[0]
    h = F.group_norm(x, gn_groups(in_c), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps=1e-5)
    h = F.conv2d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + F.linear(t_emb, sd[p + ".time_proj.weight"], sd[p + ".time_proj.bias"])[..., None, None]
    h = F.group_norm(h, gn_groups(out_c), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps=1e-5)
    h = F.conv2d(F.silu(h), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    h = self_attention(sd, p + ".attn", h, 4)
    if level is not None and p in FREQ_BLOCKS_0409:
        h = freq_aware_block_0409(sd, p + ".freq_guide", h, level)
        h = hfcm_0409(sd, p + ".hfcm", h, level)
    if p + ".shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".shortcut.weight"], sd[p + ".shortcut.bias"])
    return x + h


def unet0409_forward(sd: SD, x: torch.Tensor, t: torch.Tensor, level: Optional[torch.Tensor] = None) -> torch.Tensor:
    """JPEGDiffusionModel.forward of the 0409 notebook (0409_method.ipynb#c0:L402-428): 1x1 output conv, no tanh."""
    with torch.no_grad():
        t_emb = time_embedding(sd, t)
        if level is None:
            level = t.clone()
        blk = lambda p, z: res_attn_block_0409(sd, p, z, t_emb, level)
        pool = lambda z: F.max_pool2d(z, 2)
        d1 = blk("down1", x)
        d2 = blk("down2", pool(d1))
        d3 = blk("down3", pool(d2))
        d4 = blk("down4", pool(d3))
        d5 = blk("down5", pool(d4))
        bn = blk("bottleneck.2", blk("bottleneck.1", blk("bottleneck.0", pool(d5))))
        u1 = blk("up1", _up_cat(bn, d5))
        u2 = blk("up2", _up_cat(u1, d4))
        u3 = blk("up3", _up_cat(u2, d3))
        u4 = blk("up4", _up_cat(u3, d2))
        u5 = blk("up5", _up_cat(u4, d1))
        return F.conv2d(u5, sd["out_conv.weight"], sd["out_conv.bias"])


# ----------------------------------------------------------------------------------------------------------
# host codec round trip (the data-consistency operator)
# ----------------------------------------------------------------------------------------------------------
def quantize_u8(x: torch.Tensor) -> torch.Tensor:
    """`(x*127.5+127.5).clamp(0,255).to(torch.uint8)` -- truncation, webp_inference.py:509."""
    return (x * 127.5 + 127.5).clamp(0, 255).to(torch.uint8)


def codec_roundtrip(x: torch.Tensor, quality, codec: str) -> torch.Tensor:
    """webp_compress webp_inference.py:506-528 / avif_compress avif_inference.py:64-98 (without the silent JPEG
    fallback -- an AVIF failure raises) / jpeg_compress svd.ipynb#c1:L20-44."""
    from PIL import Image
    u8 = quantize_u8(x).cpu().numpy()
    out = np.empty(u8.shape, dtype=np.float32)
    for n in range(u8.shape[0]):
        img = Image.fromarray(np.ascontiguousarray(u8[n].transpose(1, 2, 0)), mode="RGB")
        buf = io.BytesIO()
        if codec == "webp":
            q = max(0, min(100, int(quality)))
            img.save(buf, format="WEBP", quality=q)
        elif codec == "avif":
            q = max(1, min(100, int(quality)))
            img.save(buf, format="AVIF", quality=q)
        elif codec == "jpeg":
            q = max(1, min(100, int(quality)))
            img.save(buf, format="JPEG", quality=q, subsampling="4:4:4" if q > 30 else "4:2:0")
        else:
            raise ValueError(codec)
        buf.seek(0)
        dec = np.asarray(Image.open(buf).convert("RGB"), dtype=np.uint8)
        out[n] = dec.transpose(2, 0, 1).astype(np.float32) / 255.0
    return torch.from_numpy(out).sub(0.5).mul(2.0)


# ----------------------------------------------------------------------------------------------------------
# sampler pieces
# ----------------------------------------------------------------------------------------------------------
def phase_consistency(x: torch.Tensor, ref: torch.Tensor, alpha: float = 0.7) -> torch.Tensor:
    """webp_inference.py:531-550."""
    xf = torch.fft.fft2(x)
    ph = torch.angle(torch.fft.fft2(ref))
    mag = torch.abs(xf)
    adj = torch.fft.ifft2(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))).real
    return alpha * x + (1 - alpha) * adj


def ddrm_update(x_theta, codec_x_theta, y, z, t, sigma_scale: float, eta: float = 0.85, eta_b: float = 1.0):
    """webp_inference.py:584-592: x' = x_theta - codec + y ; x_t = eta_b x' + (1-eta_b) x_theta + eta*(sigma_scale*t)*z."""
    x_prime = x_theta - codec_x_theta + y
    noise = z * (t.float() * sigma_scale).view(-1, 1, 1, 1)
    return eta_b * x_prime + (1 - eta_b) * x_theta + eta * noise


DDRM = {  # codec, noise scale, phase rule (quality threshold, period, alpha)
    "webp": dict(codec="webp", sigma=0.2, q_thr=15, period=5, alpha=0.7),   # webp_inference.py:588,595-597
    "jpeg": dict(codec="jpeg", sigma=0.2, q_thr=20, period=5, alpha=0.7),   # svd.ipynb#c1:L371,378-380
    "avif": dict(codec="avif", sigma=0.15, q_thr=30, period=3, alpha=0.8),  # avif_inference.py:445,452-454
}


def ddrm_sample(model_fn: Callable, y0: torch.Tensor, quality: int, steps: int, family: str,
                noise_fn: Callable, eta: float = 0.85, eta_b: float = 1.0,
                codec_fn: Optional[Callable] = None, trace: Optional[list] = None) -> torch.Tensor:
    """DDRM{WebP,AVIF,JPEG}Sampler.sample (webp_inference.py:557-602).  `noise_fn(i, like)` supplies z."""
    cfg = DDRM[family]
    codec_fn = codec_fn or (lambda z, q: codec_roundtrip(z, q, cfg["codec"]))
    x_t = y0.clone()
    y = y0.clone()
    B = y0.shape[0]
    for i in range(steps - 1, -1, -1):
        t = torch.full((B,), i).float() / steps
        x_theta = model_fn(x_t, t, t.clone())
        c = codec_fn(x_theta, quality)
        if i > 0:
            x_t = ddrm_update(x_theta, c, y, noise_fn(i, x_t), t, cfg["sigma"], eta, eta_b)
            if quality < cfg["q_thr"] and i % cfg["period"] == 0:
                x_t = phase_consistency(x_t, y, cfg["alpha"])
        else:
            x_t = x_theta - c + y
        if trace is not None:
            trace.append(x_t.clone())
    return x_t


def svd_structure_preservation(x: torch.Tensor, k_ratio: float = 0.5) -> torch.Tensor:
    """0409_method.ipynb#c0:L321-346 -- rank-k truncation of every (b,c) plane, k = max(1, int(min(h,w)*k_ratio))."""
    b, c, h, w = x.shape
    U, S, Vh = torch.linalg.svd(x.reshape(b * c, h, w), full_matrices=False)
    k = max(1, int(min(h, w) * k_ratio))
    S = S.clone()
    S[:, k:] = 0
    return (U * S[:, None, :] @ Vh).reshape(b, c, h, w)


def gmm_sample(model_fn: Callable, x0: torch.Tensor, steps: int, noise_fn: Callable, coin_fn: Callable,
               num_timesteps: int = 100, use_phase_consistency: bool = True, use_svd_guide: bool = True,
               guidance_scale: float = 1.0, trace: Optional[list] = None) -> torch.Tensor:
    """GaussianMixtureSampler.sample 0409_method.ipynb#c1:L395-449.  `coin_fn(i)` supplies torch.rand(1).item()."""
    x_t = x0.clone()
    y = x0.clone()
    B = x0.shape[0]
    for i in range(steps - 1, -1, -1):
        t = torch.full((B,), i).float() / num_timesteps
        pred = model_fn(x_t, t, t.clone())
        if use_svd_guide and i > steps // 2:
            kr = i / steps
            g = kr * 0.3
            pred = (1 - g) * pred + g * (y - svd_structure_preservation(x_t, kr))
        if i > 0:
            x0p = x_t + pred
            mu1 = x0p * 0.9 + x_t * 0.1
            mu2 = x0p * 1.1 - x_t * 0.1
            p_cons = max(0.2, min(0.8, i / steps))
            mean = mu1 if coin_fn(i) < p_cons else mu2
            x_n = mean + (0.1 * i / steps * guidance_scale) * noise_fn(i, x_t)
            if use_phase_consistency and i % 5 == 0:
                x_n = phase_consistency(x_n, y, 0.6 + 0.3 * (1 - i / steps))
            x_t = x_n
        else:
            x_t = x_t + pred
        if trace is not None:
            trace.append(x_t.clone())
    return x_t


def ddpm_schedule(T: int = 100):
    """experiments/code/ddpm.ipynb#c5: beta = linspace(1e-4, 0.02, T)."""
    betas = torch.linspace(1e-4, 0.02, T)
    alphas = 1.0 - betas
    return betas, alphas, torch.cumprod(alphas, dim=0)


def ddpm_posterior_mean(x_t, eps, t: int, T: int = 100):
    """ddpm.ipynb#c5:L63-76: x = (x - (1-alpha_t)/sqrt(1-alphabar_t) * eps) / sqrt(alpha_t)  (no noise term)."""
    betas, alphas, abar = ddpm_schedule(T)
    return (x_t - (1 - alphas[t]) / torch.sqrt(1 - abar[t]) * eps) / torch.sqrt(alphas[t])


# ----------------------------------------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------------------------------------
def _gauss_win(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim(X: torch.Tensor, Y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """pytorch_msssim.ssim(X, Y, data_range, size_average=True) restated from the library's published algorithm
    (gaussian window 11, sigma 1.5, separable 'valid' filtering, K=(0.01,0.03)).  PARITY UNPINNED: the package
    is not installed here and the reference pins no version."""
    C = X.shape[1]
    g = _gauss_win().to(X.dtype)
    def filt(z):
        z = F.conv2d(z, g.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
        return F.conv2d(z, g.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = filt(X), filt(Y)
    s11 = filt(X * X) - mu1 * mu1
    s22 = filt(Y * Y) - mu2 * mu2
    s12 = filt(X * Y) - mu1 * mu2
    cs = (2 * s12 + C2) / (s11 + s22 + C2)
    sm = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs
    return sm.flatten(2).mean(-1).mean()


def frequency_aware_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """webp_training.py:105-132 (SSIM restated, see `ssim`)."""
    spatial = F.mse_loss(pred, target)
    p01, t01 = pred * 0.5 + 0.5, target * 0.5 + 0.5
    freq = 0
    for c in range(3):
        fp, ft = torch.fft.rfft2(p01[:, c]), torch.fft.rfft2(t01[:, c])
        freq = freq + F.mse_loss(torch.abs(fp), torch.abs(ft)) + 0.5 * F.mse_loss(torch.angle(fp), torch.angle(ft))
    return spatial + 0.5 * freq + 0.3 * (1.0 - ssim(p01, t01, 1.0))


def avif_frequency_aware_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """avif.py:126-164 (SSIM restated, see `ssim`): full fft2 spectrum, gradient (edge) term, weights 0.3 / 0.4 / 0.2."""
    spatial = F.mse_loss(pred, target)
    p01, t01 = pred * 0.5 + 0.5, target * 0.5 + 0.5

    def gradient_loss(x, y):                                           # avif.py:136-144
        gxx = torch.abs(x[:, :, :-1, :] - x[:, :, 1:, :]); gxy = torch.abs(x[:, :, :, :-1] - x[:, :, :, 1:])
        gyx = torch.abs(y[:, :, :-1, :] - y[:, :, 1:, :]); gyy = torch.abs(y[:, :, :, :-1] - y[:, :, :, 1:])
        return F.mse_loss(gxx, gyx) + F.mse_loss(gxy, gyy)
    edge = gradient_loss(p01, t01)
    freq = 0
    for c in range(3):
        fp, ft = torch.fft.fft2(p01[:, c]), torch.fft.fft2(t01[:, c])
        freq = freq + F.mse_loss(torch.abs(fp), torch.abs(ft)) + 0.3 * F.mse_loss(torch.angle(fp), torch.angle(ft))
    return spatial + 0.3 * freq + 0.4 * (1.0 - ssim(p01, t01, 1.0)) + 0.2 * edge


def training_step_reference(sd: SD, xt, t, x0, family: str = "webp"):
    """Loss and gradients of one training step (webp_training.py:511-521 / avif.py:562-573, dropout disabled) by torch.autograd.
    Parameters the forward never touches (AVIFAdaptiveTransform.inverse_weights) have no gradient and are left out, as
    torch leaves their .grad None."""
    params = {k: (v.clone().requires_grad_() if v.dtype.is_floating_point and not k.endswith("dct_matrix") else v) for k, v in sd.items()}
    pred = unet_forward(params, xt, t, t.clone(), family, enable_grad=True)
    loss = (avif_frequency_aware_loss if family == "avif" else frequency_aware_loss)(xt + pred, x0)     # avif.py:570
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}


def color_l1(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """The channel-weighted L1 part of color_preservation_loss, 0409_method.ipynb#c0:L66-76."""
    p = (pred * 0.5 + 0.5).clamp(0, 1)
    q = (target * 0.5 + 0.5).clamp(0, 1)
    l = [F.l1_loss(p[:, c], q[:, c]) for c in range(3)]
    return 0.25 * l[0] + 0.5 * l[1] + 0.25 * l[2]


def color_preservation_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """0409_method.ipynb#c0:L64-82 = colour L1 + 0.5*(1-SSIM)."""
    p = (pred * 0.5 + 0.5).clamp(0, 1)
    q = (target * 0.5 + 0.5).clamp(0, 1)
    return color_l1(pred, target) + 0.5 * (1 - ssim(p, q, 1.0))


def color_loss_conv_deep(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """conv_deep.ipynb#c0:L60-73: same weighted L1 on clamped [0,1] images, no SSIM term."""
    return color_l1(pred, target)


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """PSNR on clamped [0,1] images, webp_inference.py:691-704: -10 log10(mse + 1e-8)."""
    a01 = (a * 0.5 + 0.5).clamp(0, 1)
    b01 = (b * 0.5 + 0.5).clamp(0, 1)
    return float(-10 * torch.log10(F.mse_loss(a01, b01) + 1e-8))


# ----------------------------------------------------------------------------------------------------------
# DCT-domain JPEG projection: experiments/code/dct.ipynb cell 2, DCTProcessor
# ----------------------------------------------------------------------------------------------------------
JPEG_QUANT_Y = torch.tensor([      # dct.ipynb#c2:L47-56
    [16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55], [14, 13, 16, 24, 40, 57, 69, 56],
    [14, 17, 22, 29, 51, 87, 80, 62], [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
    [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], dtype=torch.float32)
JPEG_QUANT_C = torch.tensor([      # dct.ipynb#c2:L58-67
    [17, 18, 24, 47, 99, 99, 99, 99], [18, 21, 26, 66, 99, 99, 99, 99], [24, 26, 56, 99, 99, 99, 99, 99],
    [47, 66, 99, 99, 99, 99, 99, 99], [99] * 8, [99] * 8, [99] * 8, [99] * 8], dtype=torch.float32)


def jpeg_quant_tables(quality) -> Tuple[torch.Tensor, torch.Tensor]:
    """dct.ipynb#c2:L105-112: scale = 50/q below 50, 2 - q/50 above; tables rounded and clamped to >= 1."""
    scale = 50 / quality if quality < 50 else 2 - quality / 50
    return (torch.clamp((JPEG_QUANT_Y * scale).round(), min=1), torch.clamp((JPEG_QUANT_C * scale).round(), min=1))


def dct_jpeg_project(images: torch.Tensor, quality) -> torch.Tensor:
    """DCTProcessor.jpeg_compress (dct.ipynb#c2:L100-139), vectorised: per channel and 8x8 block, orthonormal DCT-II of
    (block - 128), round(dct / Q) * Q with the luma table for channel 0 and the chroma table for the others, inverse DCT,
    + 128.  Images are on the 0..255 scale; no colour conversion, no clamping (as in the reference)."""
    B, C, H, W = images.shape
    assert H % 8 == 0 and W % 8 == 0
    k = torch.arange(8, dtype=torch.float64)
    D = torch.cos((2 * k[None, :] + 1) * k[:, None] * math.pi / 16) * 0.5          # D[u, x], 0.25 = 0.5 * 0.5 over both axes
    D[0] *= 1 / math.sqrt(2)
    D = D.to(images.dtype)
    qy, qc = jpeg_quant_tables(quality)
    blocks = (images - 128).reshape(B, C, H // 8, 8, W // 8, 8).permute(0, 1, 2, 4, 3, 5)      # [B,C,by,bx,8,8]
    coef = D @ blocks @ D.T
    q = torch.stack([qy] + [qc] * (C - 1)).to(images.dtype).view(1, C, 1, 1, 8, 8)
    deq = torch.round(coef / q) * q
    rec = D.T @ deq @ D + 128
    return rec.permute(0, 1, 2, 4, 3, 5).reshape(B, C, H, W)


# ----------------------------------------------------------------------------------------------------------
# Philox4x32-10 known-answer reference for the in-kernel noise generator
# ----------------------------------------------------------------------------------------------------------
def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11).  counter [...,4] uint32, key [...,2] uint32 -> [...,4] uint32."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
    c = [counter[..., i].astype(np.uint32) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32)
    k1 = key[..., 1].astype(np.uint32)
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = ((k0.astype(np.uint64) + np.uint64(W0)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        k1 = ((k1.astype(np.uint64) + np.uint64(W1)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_normal(seed: int, step: int, n: int) -> np.ndarray:
    """The product's noise definition: element e (flat NCHW index) of step `step` takes word e%4 of
    Philox(counter=(e//4, step, 0, 0), key=(seed_lo, seed_hi)); words (0,1) and (2,3) feed Box-Muller:
        u = ((w >> 9) + 0.5) * 2^-23  (23-bit uniforms: exactly representable in fp32) ; r = sqrt(-2 ln u1)
        z_a = r cos(2 pi u2) ; z_b = r sin(2 pi u2).
    Evaluated in float64 here; the kernel uses fp32 intrinsics (tolerance stated in the tests)."""
    groups = (n + 3) // 4
    ctr = np.zeros((groups, 4), dtype=np.uint32)
    ctr[:, 0] = np.arange(groups, dtype=np.uint64).astype(np.uint32)
    ctr[:, 1] = np.uint32(step)
    key = np.zeros((groups, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w = (philox4x32_10(ctr, key) >> np.uint32(9)).astype(np.float64)
    u = (w + 0.5) * (2.0 ** -23)
    r0 = np.sqrt(-2.0 * np.log(u[:, 0]))
    r1 = np.sqrt(-2.0 * np.log(u[:, 2]))
    z = np.stack([r0 * np.cos(2 * np.pi * u[:, 1]), r0 * np.sin(2 * np.pi * u[:, 1]),
                  r1 * np.cos(2 * np.pi * u[:, 3]), r1 * np.sin(2 * np.pi * u[:, 3])], axis=-1)
    return z.reshape(-1)[:n]
