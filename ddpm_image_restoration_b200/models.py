"""Drop-in replacements for the reference's UNets: WebPDiffusionModel (webp_inference.py:330-399),
JPEGDiffusionModel (svd.ipynb#c1:L242-311) and AVIFDiffusionModel (avif_inference.py:316-385).

Same constructor (no arguments), same `forward(x, t, compression_level=None)`, same checkpoint keys/shapes
(356 / 356 / 634 state_dict entries), same default initialisation (the parameter containers are the stock
torch.nn modules).  The forward pass does not call those modules: it runs the hand-written sm_100a kernels of
libddpmir.so on NHWC activations (bf16 by default, fp32 in check mode).  CPU tensors raise -- no fallback.
"""
import math

import torch
from torch import nn

from . import ops

_FAMILY = {
    "webp": dict(heads=4, bs=4, low=3, clamp=(0.15, 1.9), tail=0.1),   # webp_inference.py:208,247,263,295,397
    "jpeg": dict(heads=4, bs=8, low=4, clamp=(0.2, 2.0), tail=0.1),    # svd.ipynb#c1:L120,159,175,207,309
    "avif": dict(heads=8, bs=8, tail=0.15),                            # avif_inference.py:123,281,383
}


def _groups(c):
    g = min(8, c)
    while c % g != 0 and g > 1:
        g -= 1
    return g


def _dct_matrix(n):
    d = torch.zeros(n, n)
    for i in range(n):
        for j in range(n):
            if i == 0:
                d[i, j] = 1.0 / torch.sqrt(torch.tensor(float(n)))
            else:
                d[i, j] = torch.sqrt(torch.tensor(2.0 / n)) * torch.cos(torch.tensor(math.pi * (2 * j + 1) * i / (2 * n)))
    return d


# ---------------------------------------------------------------------------------------------------------
# parameter containers (names = the reference's checkpoint keys)
# ---------------------------------------------------------------------------------------------------------
class TimeEmbedding(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.proj = nn.Sequential(nn.Linear(dim, dim * 4), nn.SiLU(), nn.Linear(dim * 4, dim))


class DCTLayer(nn.Module):
    def __init__(self, block_size):
        super().__init__()
        self.block_size = block_size
        self.register_buffer("dct_matrix", _dct_matrix(block_size))


def _gate_mlp(c_in, c_mid, c_out, k, act):
    return nn.Sequential(nn.Conv2d(c_in, c_mid, k, padding=k // 2), act, nn.Conv2d(c_mid, c_out, k, padding=k // 2),
                         nn.Sigmoid())


class DCTFreqAwareBlock(nn.Module):
    """WebPFreqAwareBlock / JPEGFreqAwareBlock parameters."""

    def __init__(self, channels, block_size):
        super().__init__()
        self.block_size = block_size
        self.dct = DCTLayer(block_size)
        self.low_freq_attn = _gate_mlp(channels, channels // 2, channels, 1, nn.LeakyReLU(0.2))
        self.high_freq_attn = _gate_mlp(channels, channels // 2, channels, 1, nn.LeakyReLU(0.2))
        self.conv_out = nn.Conv2d(channels, channels, 3, padding=1)


class AVIFAdaptiveTransform(nn.Module):
    def __init__(self, channels, block_size=8):
        super().__init__()
        self.block_size = block_size
        self.channels = channels
        self.transform_weights = nn.Parameter(torch.randn(channels, block_size, block_size))
        self.inverse_weights = nn.Parameter(torch.randn(channels, block_size, block_size))  # unused by forward, as in the reference
        self.quantization = _gate_mlp(channels, channels, channels, 1, nn.ReLU())


class AVIFFreqAwareBlock(nn.Module):
    def __init__(self, channels, block_size=8):
        super().__init__()
        self.block_size = block_size
        self.adaptive_transform = AVIFAdaptiveTransform(channels, block_size)
        self.multi_scale_attn = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d(s), nn.Conv2d(channels, channels // 4, 1), nn.ReLU(),
                          nn.Conv2d(channels // 4, channels, 1), nn.Sigmoid()) for s in (1, 2, 4, 8)])
        self.color_consistency = _gate_mlp(channels, channels, channels, 1, nn.ReLU())
        self.edge_preserve = _gate_mlp(channels, channels // 2, channels, 3, nn.ReLU())
        self.conv_out = nn.Conv2d(channels, channels, 3, padding=1)


class ResAttnBlock(nn.Module):
    def __init__(self, family, in_c, out_c, time_dim, dropout=0.1):
        super().__init__()
        fam = _FAMILY[family]
        self.in_c, self.out_c = in_c, out_c
        self.norm1 = nn.GroupNorm(_groups(in_c), in_c)
        self.conv1 = nn.Conv2d(in_c, out_c, 3, padding=1)
        self.time_proj = nn.Linear(time_dim, out_c)
        self.norm2 = nn.GroupNorm(_groups(out_c), out_c)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_c, out_c, 3, padding=1)
        self.attn = nn.MultiheadAttention(out_c, fam["heads"], batch_first=True)
        self.freq_guide = AVIFFreqAwareBlock(out_c) if family == "avif" else DCTFreqAwareBlock(out_c, fam["bs"])
        self.shortcut = nn.Conv2d(in_c, out_c, 1) if in_c != out_c else nn.Identity()


_BLOCKS = [("down1", 3, 64), ("down2", 64, 128), ("down3", 128, 256), ("down4", 256, 512), ("down5", 512, 512),
           ("bottleneck.0", 512, 1024), ("bottleneck.1", 1024, 1024), ("bottleneck.2", 1024, 512),
           ("up1", 1024, 512), ("up2", 1024, 256), ("up3", 512, 128), ("up4", 256, 64), ("up5", 128, 64)]


class _RestorationUNet(nn.Module):
    family = None

    def __init__(self):
        super().__init__()
        fam = self.family
        time_dim = 256
        mk = lambda i, o: ResAttnBlock(fam, i, o, time_dim)
        self.time_embed = TimeEmbedding(time_dim)
        self.down1, self.down2, self.down3 = mk(3, 64), mk(64, 128), mk(128, 256)
        self.down4, self.down5 = mk(256, 512), mk(512, 512)
        self.pool = nn.MaxPool2d(2)
        self.bottleneck = nn.Sequential(mk(512, 1024), mk(1024, 1024), mk(1024, 512))
        self.up1, self.up2, self.up3 = mk(1024, 512), mk(1024, 256), mk(512, 128)
        self.up4, self.up5 = mk(256, 64), mk(128, 64)
        if fam == "avif":
            self.avif_layer = AVIFAdaptiveTransform(64, block_size=8)
        else:
            self.dct_layer = DCTLayer(block_size=_FAMILY[fam]["bs"])
        self.out_conv = nn.Sequential(nn.GroupNorm(8, 64), nn.SiLU(), nn.Conv2d(64, 3, 3, padding=1), nn.Tanh())
        self._packed = None
        self._packed_key = None
        self.precision = "bf16"   # "bf16" (production) or "fp32" (check mode)
        self.impl = ops.IMPL_AUTO  # kernel family for the GEMM-shaped ops and attention

    # -- configuration -------------------------------------------------------------------------------------
    def set_precision(self, precision):
        if precision not in ("bf16", "fp32"):
            raise ValueError(precision)
        self.precision = precision
        return self

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def invalidate_packed(self):
        """Drop the pre-packed weight copies.  prepack() notices load_state_dict, .to()/.cuda(), set_precision and in-place
        edits that bump a parameter's version counter (nn.init.*_, p.copy_(), optimizer steps of torch.optim); writes through
        `p.data` or raw kernels bypass that counter -- call this after them (training.Trainer does)."""
        self._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # -- weight pre-packing (once per precision; OIHW fp32 -> [N, (kh,kw,cin)] in the activation dtype) ----------
    def prepack(self):
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        key = (self.precision, next(self.parameters()).device, sum(p._version for p in self.parameters()))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        dev = key[1]
        if dev.type != "cuda":
            raise RuntimeError("the B200 UNet runs on CUDA only (no CPU fallback); call .cuda() first")

        def cast(w):
            w = w.contiguous().float()
            return ops.cast_bf16(w) if dt == torch.bfloat16 else w

        def conv3(w):   # [N,Cin,3,3] -> [N, 9*Cin]
            return cast(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))

        def lin(w):     # [N,K,1,1] or [N,K]
            return cast(w.reshape(w.shape[0], -1))

        f32 = lambda v: v.contiguous().float()
        P = {}
        for p, ci, co in _BLOCKS:
            q = {}
            if ci == 3:
                q["conv1_w"] = f32(sd[f"{p}.conv1.weight"])
                q["sc_w"] = f32(sd[f"{p}.shortcut.weight"])
            else:
                q["conv1_w"] = conv3(sd[f"{p}.conv1.weight"])
                if ci != co:
                    q["sc_w"] = lin(sd[f"{p}.shortcut.weight"])
            q["conv2_w"] = conv3(sd[f"{p}.conv2.weight"])
            in_w, in_b = f32(sd[f"{p}.attn.in_proj_weight"]), f32(sd[f"{p}.attn.in_proj_bias"])
            if dt == torch.bfloat16:
                # production path: softmax scale and log2(e) folded into the q rows (ops.attention_prescaled)
                fold = torch.ones(3 * co, 1, device=dev)
                fold[:co] = ops.Q_PRESCALE_LOG2E / math.sqrt(co // _FAMILY[self.family]["heads"])
                in_w, in_b = in_w * fold, in_b * fold[:, 0]
            q["in_w"], q["in_b"] = lin(in_w), in_b.contiguous()
            q["out_w"] = lin(sd[f"{p}.attn.out_proj.weight"])
            f = f"{p}.freq_guide"
            if self.family == "avif":
                a = f"{f}.adaptive_transform"
                q["q0_w"] = lin(sd[f"{a}.quantization.0.weight"])
                q["q2_w"] = lin(sd[f"{a}.quantization.2.weight"])
                q["c0_w"] = lin(sd[f"{f}.color_consistency.0.weight"])
                q["c2_w"] = lin(sd[f"{f}.color_consistency.2.weight"])
                # edge_preserve: C -> C/2 -> C.  At C = 64 the hidden width (32) is below the tcgen05 kernel's 64-wide K block, so
                # the hidden layer is zero-padded to 64 channels: relu(0 * x + 0) = 0 feeds zero weights, the result is identical
                w0, b0, w2 = sd[f"{f}.edge_preserve.0.weight"], f32(sd[f"{f}.edge_preserve.0.bias"]), sd[f"{f}.edge_preserve.2.weight"]
                if dt == torch.bfloat16 and w0.shape[0] < 64:
                    padn = 64 - w0.shape[0]
                    w0 = torch.cat([w0, w0.new_zeros(padn, *w0.shape[1:])], 0)
                    b0 = torch.cat([b0, b0.new_zeros(padn)], 0)
                    w2 = torch.cat([w2, w2.new_zeros(w2.shape[0], padn, 3, 3)], 1)
                q["e0_w"], q["e0_b"], q["e2_w"] = conv3(w0), b0.contiguous(), conv3(w2)
                for i in range(4):
                    q[f"ms{i}_w1"] = f32(sd[f"{f}.multi_scale_attn.{i}.1.weight"].reshape(co // 4, co))
                    q[f"ms{i}_w3"] = f32(sd[f"{f}.multi_scale_attn.{i}.3.weight"].reshape(co, co // 4))
            else:
                # stacked low/high gate MLP: hidden = [low | high] halves, second layer concatenated along K
                q["g1_w"] = cast(torch.cat([sd[f"{f}.low_freq_attn.0.weight"], sd[f"{f}.high_freq_attn.0.weight"]], 0).reshape(co, co))
                q["g1_b"] = f32(torch.cat([sd[f"{f}.low_freq_attn.0.bias"], sd[f"{f}.high_freq_attn.0.bias"]], 0))
                q["g2_w"] = cast(torch.cat([sd[f"{f}.low_freq_attn.2.weight"].reshape(co, co // 2),
                                            sd[f"{f}.high_freq_attn.2.weight"].reshape(co, co // 2)], 1))
            q["fo_w"] = conv3(sd[f"{f}.conv_out.weight"])
            P[p] = q
        if self.family == "avif":
            P["tail_q0_w"] = lin(sd["avif_layer.quantization.0.weight"])
            P["tail_q2_w"] = lin(sd["avif_layer.quantization.2.weight"])
        self._packed, self._packed_key = P, key
        return P

    # -- forward -------------------------------------------------------------------------------------------
    def forward(self, x, t, compression_level=None):
        if self.training:
            raise NotImplementedError(
                "forward() is the inference path (no autograd history); call .eval() for inference, or use "
                "ddpm_image_restoration_b200.training.Trainer(model).train_step(xt, t, x0) for the training step "
                "(webp_training.py:476-537: explicit forward tape + hand-written backward kernels)")
        if not x.is_cuda:
            raise RuntimeError("the B200 UNet runs on CUDA only (no CPU fallback)")
        with torch.no_grad():
            return self._forward(x.contiguous().float(), t.contiguous().float(), compression_level)

    def _forward(self, x, t, level, taps=None):
        fam = _FAMILY[self.family]
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32   # GEMM operand dtype
        P = self.prepack()
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        level = t.clone() if level is None else level.contiguous().float().view(-1)
        B = x.shape[0]
        t_emb = ops.time_embed(t, sd["time_embed.proj.0.weight"], sd["time_embed.proj.0.bias"],
                               sd["time_embed.proj.2.weight"], sd["time_embed.proj.2.bias"])
        if self.family == "avif":
            boosts = (torch.clamp(0.5 + 0.5 * (1.0 - level), 0.3, 1.5).contiguous(),
                      torch.clamp(0.7 + 0.3 * (1.0 - level), 0.5, 1.3).contiguous())
        else:
            boosts = (torch.clamp(1.0 - level, fam["clamp"][0], fam["clamp"][1]).contiguous(),)

        blk = lambda p, z: self._block(p, z, t_emb, boosts, sd, P[p], dt)
        d1 = blk("down1", x)
        d2 = blk("down2", ops.maxpool2(d1))
        d3 = blk("down3", ops.maxpool2(d2))
        d4 = blk("down4", ops.maxpool2(d3))
        d5 = blk("down5", ops.maxpool2(d4))
        bn = blk("bottleneck.0", ops.maxpool2(d5))
        bn = blk("bottleneck.1", bn)
        bn = blk("bottleneck.2", bn)
        u1 = blk("up1", ops.upsample2_concat(bn, d5))
        u2 = blk("up2", ops.upsample2_concat(u1, d4))
        u3 = blk("up3", ops.upsample2_concat(u2, d3))
        u4 = blk("up4", ops.upsample2_concat(u3, d2))
        u5 = blk("up5", ops.upsample2_concat(u4, d1))
        if taps is not None:
            taps.update(t_emb=t_emb, d1=d1, d2=d2, d3=d3, d4=d4, d5=d5, bn=bn, u1=u1, u2=u2, u3=u3, u4=u4, u5=u5)
        if self.family == "avif":
            tr = ops.block_transform(u5, sd["avif_layer.transform_weights"], 0.0, 1.0, out_dtype=dt)
            q1 = ops.gemm(tr, P["tail_q0_w"], 64, self.impl, bias=sd["avif_layer.quantization.0.bias"], act=ops.ACT_RELU)
            tail = torch.full((B,), fam["tail"], dtype=torch.float32, device=x.device)
            comb = ops.gemm(q1, P["tail_q2_w"], 64, self.impl, out_dtype=torch.float32,
                            bias=sd["avif_layer.quantization.2.bias"], act=ops.ACT_SIGMOID, img_scale=tail, mul=tr, res=u5)
        else:
            comb = ops.block_transform(u5, sd["dct_layer.dct_matrix"], 1.0, fam["tail"])
        st = ops.groupnorm_stats(comb, 8)
        a = ops.groupnorm_apply(comb, st, sd["out_conv.0.weight"], sd["out_conv.0.bias"], ops.ACT_SILU, out_dtype=dt)
        return ops.out_conv_tanh(a, sd["out_conv.2.weight"], sd["out_conv.2.bias"])

    def _block(self, p, x, t_emb, boosts, sd, W, dt):
        """{WebP,JPEG,AVIF}ResAttnBlock.forward (webp_inference.py:303-327) on NHWC activations.

        Two tensor classes: the *stream* (block inputs/outputs, conv outputs feeding GroupNorm, the attention and
        shortcut residuals) stays fp32 -- only element-wise / normalisation kernels read it -- while every GEMM
        operand is written in `dt` (bf16 in production) by the kernel that produces it.  Keeping the stream in fp32
        removes the largest avoidable share of the bf16 error budget ("error budget")."""
        fam = _FAMILY[self.family]
        impl = self.impl
        f32 = torch.float32
        first = p == "down1"
        co = sd[f"{p}.conv1.bias"].shape[0]
        tb = ops.linear_rows(t_emb, sd[f"{p}.time_proj.weight"], sd[f"{p}.time_proj.bias"])
        if first:
            st = ops.groupnorm_stats(x, 3, nchw=True)
            h1 = ops.conv_input(x, W["conv1_w"], sd[f"{p}.conv1.bias"], f32, st, sd[f"{p}.norm1.weight"],
                                sd[f"{p}.norm1.bias"], row_bias=tb)
            sc = ops.conv_input(x, W["sc_w"], sd[f"{p}.shortcut.bias"], f32)
        else:
            ci = x.shape[-1]
            st = ops.groupnorm_stats(x, _groups(ci))
            if "sc_w" in W:
                a, x_op = ops.groupnorm_apply(x, st, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE,
                                              out_dtype=dt, raw_copy=True)
                sc = ops.gemm(x_op, W["sc_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.shortcut.bias"])
            else:
                a = ops.groupnorm_apply(x, st, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE, out_dtype=dt)
                sc = x
            h1 = ops.conv3x3(a, W["conv1_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.conv1.bias"], row_bias=tb)
        st = ops.groupnorm_stats(h1, _groups(co))
        a = ops.groupnorm_apply(h1, st, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], ops.ACT_GELU, out_dtype=dt)
        h2, h2_op = ops.conv3x3(a, W["conv2_w"], co, impl, out_dtype=f32, out2_dtype=dt, bias=sd[f"{p}.conv2.bias"])
        Bn, H, Wd, _ = h2.shape
        qdt = ops.qkv_dtype_for_attention(H * Wd, co // fam["heads"]) if dt == torch.bfloat16 else None
        qkv = ops.gemm(h2_op, W["in_w"], 3 * co, impl, out_dtype=qdt, bias=W["in_b"])
        if dt == torch.bfloat16:
            ao = ops.attention_prescaled(qkv.view(Bn, H * Wd, 3 * co), fam["heads"]).view(Bn, H, Wd, co)
        else:
            ao = ops.attention(qkv.view(Bn, H * Wd, 3 * co), fam["heads"], impl).view(Bn, H, Wd, co)
        f = f"{p}.freq_guide"
        if self.family == "avif":
            h3, h3_op = ops.gemm(ao, W["out_w"], co, impl, out_dtype=f32, out2_dtype=dt,
                                 bias=sd[f"{p}.attn.out_proj.bias"], res=h2)
            a_ = f"{f}.adaptive_transform"
            tr = ops.block_transform(h3, sd[f"{a_}.transform_weights"], 0.0, 1.0, out_dtype=dt)
            q1 = ops.gemm(tr, W["q0_w"], co, impl, bias=sd[f"{a_}.quantization.0.bias"], act=ops.ACT_RELU)
            xt = ops.gemm(q1, W["q2_w"], co, impl, bias=sd[f"{a_}.quantization.2.bias"], act=ops.ACT_SIGMOID, mul=tr)
            pooled = ops.avgpool_pyramid(h3)                       # [85, B, C] fp32
            gates = torch.empty_like(pooled)
            off = 0
            for i, s in enumerate((1, 2, 4, 8)):
                rows = pooled[off:off + s * s].reshape(s * s * Bn, co)
                hid = ops.linear_rows(rows, W[f"ms{i}_w1"], sd[f"{f}.multi_scale_attn.{i}.1.bias"], ops.ACT_RELU)
                ops.linear_rows(hid, W[f"ms{i}_w3"], sd[f"{f}.multi_scale_attn.{i}.3.bias"], ops.ACT_SIGMOID,
                                out=gates[off:off + s * s])
                off += s * s
            c1 = ops.gemm(h3_op, W["c0_w"], co, impl, bias=sd[f"{f}.color_consistency.0.bias"], act=ops.ACT_RELU)
            color = ops.gemm(c1, W["c2_w"], co, impl, bias=sd[f"{f}.color_consistency.2.bias"], act=ops.ACT_SIGMOID,
                             img_scale=boosts[0])
            e1 = ops.conv3x3(h3_op, W["e0_w"], W["e0_b"].shape[0], impl, bias=W["e0_b"], act=ops.ACT_RELU)
            edge = ops.conv3x3(e1, W["e2_w"], co, impl, bias=sd[f"{f}.edge_preserve.2.bias"], act=ops.ACT_SIGMOID,
                               img_scale=boosts[1])
            e = ops.avif_combine(h3, xt, gates, color, edge)
        else:
            h3 = ops.gemm(ao, W["out_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.attn.out_proj.bias"], res=h2)
            d = ops.block_transform(h3, sd[f"{f}.dct.dct_matrix"], 0.0, 1.0, out_dtype=dt)
            g1 = ops.gemm(d, W["g1_w"], co, impl, bias=W["g1_b"], act=ops.ACT_LRELU02, freq_mode=1, bs=fam["bs"],
                          low=fam["low"])
            e = ops.gemm(g1, W["g2_w"], co, impl, bias=sd[f"{f}.low_freq_attn.2.bias"],
                         bias2=sd[f"{f}.high_freq_attn.2.bias"], act=ops.ACT_SIGMOID, freq_mode=2, bs=fam["bs"],
                         low=fam["low"], img_scale=boosts[0], mul=d, res=h3)
        return ops.conv3x3(e, W["fo_w"], co, impl, out_dtype=f32, bias=sd[f"{f}.conv_out.bias"], res=sc)


class WebPDiffusionModel(_RestorationUNet):
    family = "webp"


class JPEGDiffusionModel(_RestorationUNet):
    family = "jpeg"


class AVIFDiffusionModel(_RestorationUNet):
    family = "avif"
