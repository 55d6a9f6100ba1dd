"""The 0409 notebook's own UNet: experiments/code/0409_method.ipynb cell 0, L184-428 -- the model
the SVD-guided GaussianMixtureSampler was written for.  Same 13-block skeleton as the shipped families, but

* a block is GN -> conv3x3 -> +time -> GN -> SiLU -> conv3x3 -> MHA(4 heads), and the attention output REPLACES h
  (no residual around it, L307-311); six blocks (use_freq_guide=True, L378-397) then apply FrequencyAwareBlock
  (L221-263: x + conv3x3(DCT8(x)) * (squeeze-excite gate * (1 - level) + 0.5)) and HFCM (L184-218:
  conv1x1(x + sigmoid(conv3(relu(conv3(x)))) * DCT8(x) * (1 - level)));
* the output layer is a plain 1x1 convolution, no tanh (L400, L428).

Same constructor / forward / checkpoint keys (282 entries, 119 873 161 parameters) as the reference class, which is also
called JPEGDiffusionModel in that notebook; the forward runs on the libddpmir.so kernels (NHWC, bf16 operands / fp32 stream,
fp32 check mode), no CPU fallback.
"""
import math

import torch
from torch import nn

from . import ops
from .models import _BLOCKS, TimeEmbedding, _dct_matrix, _groups

FREQ_BLOCKS = ("down2", "down3", "bottleneck.0", "bottleneck.2", "up2", "up3")


class DCTLayer(nn.Module):      # 0409_method.ipynb#c0:L104-181: no buffer, the matrix is rebuilt per call in the reference
    def __init__(self, block_size=8):
        super().__init__()
        self.block_size = block_size


class HFCM(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.dct = DCTLayer(8)
        self.high_freq_attn = nn.Sequential(nn.Conv2d(channels, channels, 3, 1, 1), nn.ReLU(),
                                            nn.Conv2d(channels, channels, 3, 1, 1), nn.Sigmoid())
        self.conv_out = nn.Conv2d(channels, channels, 1)


class FrequencyAwareBlock(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.dct_layer = DCTLayer(8)
        self.freq_conv = nn.Conv2d(channels, channels, 3, padding=1)
        self.freq_attn = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(channels, channels // 4, 1), nn.ReLU(),
                                       nn.Conv2d(channels // 4, channels, 1), nn.Sigmoid())


class ResAttnBlock(nn.Module):
    def __init__(self, in_c, out_c, time_dim, dropout=0.1, use_freq_guide=False):
        super().__init__()
        self.norm1 = nn.GroupNorm(_groups(in_c), in_c)
        self.conv1 = nn.Conv2d(in_c, out_c, 3, padding=1)
        self.time_proj = nn.Linear(time_dim, out_c)
        self.norm2 = nn.GroupNorm(_groups(out_c), out_c)
        self.conv2 = nn.Conv2d(out_c, out_c, 3, padding=1)
        self.attn = nn.MultiheadAttention(out_c, 4, batch_first=True)
        self.dropout = nn.Dropout(dropout)
        self.shortcut = nn.Conv2d(in_c, out_c, 1) if in_c != out_c else nn.Identity()
        self.use_freq_guide = use_freq_guide
        if use_freq_guide:
            self.freq_guide = FrequencyAwareBlock(out_c)
            self.hfcm = HFCM(out_c)


class JPEGDiffusionModel(nn.Module):
    """Drop-in for the 0409 notebook's JPEGDiffusionModel (0409_method.ipynb#c0:L371-428)."""
    HEADS = 4

    def __init__(self):
        super().__init__()
        time_dim = 256
        mk = lambda i, o, f=False: ResAttnBlock(i, o, time_dim, use_freq_guide=f)
        self.time_embed = TimeEmbedding(time_dim)
        self.down1, self.down2, self.down3 = mk(3, 64), mk(64, 128, True), mk(128, 256, True)
        self.down4, self.down5 = mk(256, 512), mk(512, 512)
        self.pool = nn.MaxPool2d(2)
        self.bottleneck = nn.Sequential(mk(512, 1024, True), mk(1024, 1024), mk(1024, 512, True))
        self.up1, self.up2, self.up3 = mk(1024, 512), mk(1024, 256, True), mk(512, 128, True)
        self.up4, self.up5 = mk(256, 64), mk(128, 64)
        self.out_conv = nn.Conv2d(64, 3, 1)
        self._packed, self._packed_key = None, None
        self.precision = "bf16"
        self.impl = ops.IMPL_AUTO

    def set_precision(self, precision):
        if precision not in ("bf16", "fp32"):
            raise ValueError(precision)
        self.precision = precision
        return self

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # -- weight pre-packing ---------------------------------------------------------------------------------------
    def prepack(self):
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        key = (self.precision, next(self.parameters()).device)
        if self._packed is not None and self._packed_key == key:
            return self._packed
        dev = key[1]
        if dev.type != "cuda":
            raise RuntimeError("the B200 UNet runs on CUDA only (no CPU fallback); call .cuda() first")
        sd = {k: v.detach() for k, v in self.state_dict().items()}
        f32 = lambda v: v.contiguous().float()

        def cast(w):
            w = w.contiguous().float()
            return ops.cast_bf16(w) if dt == torch.bfloat16 else w

        conv3 = lambda w: cast(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
        lin = lambda w: cast(w.reshape(w.shape[0], -1))
        P = {"dct8": _dct_matrix(8).to(dev)}
        for p, ci, co in _BLOCKS:
            q = {}
            if ci == 3:
                q["conv1_w"], q["sc_w"] = f32(sd[f"{p}.conv1.weight"]), f32(sd[f"{p}.shortcut.weight"])
            else:
                q["conv1_w"] = conv3(sd[f"{p}.conv1.weight"])
                if ci != co:
                    q["sc_w"] = lin(sd[f"{p}.shortcut.weight"])
            q["conv2_w"] = conv3(sd[f"{p}.conv2.weight"])
            in_w, in_b = f32(sd[f"{p}.attn.in_proj_weight"]), f32(sd[f"{p}.attn.in_proj_bias"])
            if dt == torch.bfloat16:        # softmax scale and log2(e) folded into the q rows (ops.attention_prescaled)
                fold = torch.ones(3 * co, 1, device=dev)
                fold[:co] = ops.Q_PRESCALE_LOG2E / math.sqrt(co // self.HEADS)
                in_w, in_b = in_w * fold, in_b * fold[:, 0]
            q["in_w"], q["in_b"] = lin(in_w), in_b.contiguous()
            q["out_w"] = lin(sd[f"{p}.attn.out_proj.weight"])
            if p in FREQ_BLOCKS:
                q["fc_w"] = conv3(sd[f"{p}.freq_guide.freq_conv.weight"])
                q["fa1_w"] = f32(sd[f"{p}.freq_guide.freq_attn.1.weight"].reshape(co // 4, co))
                q["fa3_w"] = f32(sd[f"{p}.freq_guide.freq_attn.3.weight"].reshape(co, co // 4))
                q["hf0_w"] = conv3(sd[f"{p}.hfcm.high_freq_attn.0.weight"])
                q["hf2_w"] = conv3(sd[f"{p}.hfcm.high_freq_attn.2.weight"])
                q["hfc_w"] = lin(sd[f"{p}.hfcm.conv_out.weight"])
            P[p] = q
        # 1x1 output conv 64 -> 3, padded to 8 output channels for the GEMM kernel
        w = torch.zeros(8, 64, device=dev)
        w[:3] = sd["out_conv.weight"].reshape(3, 64).float()
        b = torch.zeros(8, device=dev)
        b[:3] = sd["out_conv.bias"].float()
        P["out_w"], P["out_b"] = cast(w), b
        self._packed, self._packed_key = P, key
        return P

    # -- forward ---------------------------------------------------------------------------------------------------
    def forward(self, x, t, compression_level=None):
        if self.training:
            raise NotImplementedError("inference only: call .eval()")
        if not x.is_cuda:
            raise RuntimeError("the B200 UNet runs on CUDA only (no CPU fallback)")
        with torch.no_grad():
            return self._forward(x.contiguous().float(), t.contiguous().float(), compression_level)

    def _forward(self, x, t, level):
        dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        P = self.prepack()
        sd = dict(self.named_parameters())
        level = t.clone() if level is None else level.contiguous().float().view(-1)
        inv_level = (1.0 - level).contiguous()
        t_emb = ops.time_embed(t, sd["time_embed.proj.0.weight"], sd["time_embed.proj.0.bias"],
                               sd["time_embed.proj.2.weight"], sd["time_embed.proj.2.bias"])
        blk = lambda p, z: self._block(p, z, t_emb, inv_level, sd, P, dt)
        d1 = blk("down1", x)
        d2 = blk("down2", ops.maxpool2(d1))
        d3 = blk("down3", ops.maxpool2(d2))
        d4 = blk("down4", ops.maxpool2(d3))
        d5 = blk("down5", ops.maxpool2(d4))
        bn = blk("bottleneck.2", blk("bottleneck.1", blk("bottleneck.0", ops.maxpool2(d5))))
        u1 = blk("up1", ops.upsample2_concat(bn, d5))
        u2 = blk("up2", ops.upsample2_concat(u1, d4))
        u3 = blk("up3", ops.upsample2_concat(u2, d3))
        u4 = blk("up4", ops.upsample2_concat(u3, d2))
        u5 = blk("up5", ops.upsample2_concat(u4, d1))
        u5_op = u5 if dt == torch.float32 else ops.cast_bf16(u5)
        o = ops.gemm(u5_op, P["out_w"], 8, self.impl, out_dtype=torch.float32, bias=P["out_b"])
        return o[..., :3].permute(0, 3, 1, 2).contiguous()

    def _block(self, p, x, t_emb, inv_level, sd, P, dt):
        """ResAttnBlock.forward 0409_method.ipynb#c0:L296-318 on NHWC activations (fp32 stream, `dt` GEMM operands)."""
        W, impl, f32 = P[p], self.impl, torch.float32
        co = sd[f"{p}.conv1.bias"].shape[0]
        tb = ops.linear_rows(t_emb, sd[f"{p}.time_proj.weight"], sd[f"{p}.time_proj.bias"])
        if p == "down1":
            st = ops.groupnorm_stats(x, 3, nchw=True)
            h1 = ops.conv_input(x, W["conv1_w"], sd[f"{p}.conv1.bias"], f32, st, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"],
                                row_bias=tb)
            sc = ops.conv_input(x, W["sc_w"], sd[f"{p}.shortcut.bias"], f32)
        else:
            st = ops.groupnorm_stats(x, _groups(x.shape[-1]))
            if "sc_w" in W:
                a, x_op = ops.groupnorm_apply(x, st, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE, out_dtype=dt,
                                              raw_copy=True)
                sc = ops.gemm(x_op, W["sc_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.shortcut.bias"])
            else:
                a = ops.groupnorm_apply(x, st, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE, out_dtype=dt)
                sc = x
            h1 = ops.conv3x3(a, W["conv1_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.conv1.bias"], row_bias=tb)
        st = ops.groupnorm_stats(h1, _groups(co))
        a = ops.groupnorm_apply(h1, st, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], ops.ACT_SILU, out_dtype=dt)
        h2 = ops.conv3x3(a, W["conv2_w"], co, impl, out_dtype=dt, bias=sd[f"{p}.conv2.bias"])
        Bn, H, Wd, _ = h2.shape
        qdt = ops.qkv_dtype_for_attention(H * Wd, co // self.HEADS) if dt == torch.bfloat16 else None
        qkv = ops.gemm(h2, W["in_w"], 3 * co, impl, out_dtype=qdt, bias=W["in_b"])
        if dt == torch.bfloat16:
            ao = ops.attention_prescaled(qkv.view(Bn, H * Wd, 3 * co), self.HEADS).view(Bn, H, Wd, co)
        else:
            ao = ops.attention(qkv.view(Bn, H * Wd, 3 * co), self.HEADS, impl).view(Bn, H, Wd, co)
        if p not in FREQ_BLOCKS:
            return ops.gemm(ao, W["out_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.attn.out_proj.bias"], res=sc)
        h3 = ops.gemm(ao, W["out_w"], co, impl, out_dtype=f32, bias=sd[f"{p}.attn.out_proj.bias"])
        # FrequencyAwareBlock (L234-263)
        f = f"{p}.freq_guide"
        xd = ops.block_transform(h3, P["dct8"], 0.0, 1.0, out_dtype=dt)
        xf = ops.conv3x3(xd, W["fc_w"], co, impl, out_dtype=f32, bias=sd[f"{f}.freq_conv.bias"])
        pooled = ops.avgpool_pyramid(xf)[0].contiguous()                    # global average [B, C]
        hid = ops.linear_rows(pooled, W["fa1_w"], sd[f"{f}.freq_attn.1.bias"], ops.ACT_RELU)
        gate = ops.linear_rows(hid, W["fa3_w"], sd[f"{f}.freq_attn.3.bias"], ops.ACT_SIGMOID)
        scale = (gate * inv_level[:, None] + 0.5).contiguous()
        g1, g1_op = ops.channel_scale_add(h3, xf, scale, out2_dtype=dt)
        # HFCM (L197-218)
        hf = f"{p}.hfcm"
        m1 = ops.conv3x3(g1_op, W["hf0_w"], co, impl, bias=sd[f"{hf}.high_freq_attn.0.bias"], act=ops.ACT_RELU)
        xd2 = ops.block_transform(g1, P["dct8"], 0.0, 1.0, out_dtype=dt)
        enh = ops.conv3x3(m1, W["hf2_w"], co, impl, bias=sd[f"{hf}.high_freq_attn.2.bias"], act=ops.ACT_SIGMOID,
                          img_scale=inv_level, mul=xd2, res=g1)
        return ops.gemm(enh, W["hfc_w"], co, impl, out_dtype=f32, bias=sd[f"{hf}.conv_out.bias"], res=sc)
