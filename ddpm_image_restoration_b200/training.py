"""Training step of the reference (`train_epoch_ddrm_webp`, webp_training.py:476-537) on the B200 kernels:

    pred = model(xt, t/100, t/100)                  (train mode: Dropout(0.1) active, webp_inference.py:291,313)
    loss = frequency_aware_loss(xt + pred, x0)      webp_training.py:515-518
    zero_grad; backward; clip_grad_norm_(1.0); AdamW(lr 2e-4, wd 1e-5, betas (0.9, 0.99)).step()   :521-524, 775

There is no autograd here: `Trainer` runs an explicit forward that records what the backward needs, then the
hand-written backward kernels (csrc/backward*.cu, attn_simt.cu, loss.cu, fft.cu), then a fused clip+AdamW kernel per
tensor.  All gradients live in ONE flat fp32 buffer (parameters' gradients are views into it), so data-parallel training
is a single NCCL all-reduce of that buffer over NVLink (the only collective of the whole design).

Supported model families: WebP and JPEG (the DCT frequency block); the AVIF family's extra operators have no backward yet.
"""
import math

import torch

from . import ops
from . import ops_train as T
from . import parallel
from .models import _BLOCKS, _FAMILY, _groups

F32 = torch.float32


def grad_span_starts(model):
    """Offset, in the flat gradient buffer (parameters in registration order), at which each top-level block's parameters
    start: {"time_embed": 0, "down1": ..., "bottleneck.0": ..., "out_conv": ...}.  The backward finishes blocks in reverse
    registration order, so "everything from span_start[block] on is final" holds after each block."""
    spans, off = {}, 0
    for k, p in model.named_parameters():
        owner = ".".join(k.split(".")[:2]) if k.startswith("bottleneck.") else k.split(".")[0]
        spans.setdefault(owner, off)
        off += p.numel()
    return spans


class Trainer:
    def __init__(self, model, lr=2e-4, weight_decay=1e-5, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=1.0, dropout=0.1,
                 seed=0, overlap_allreduce=True, bucket_bytes=64 << 20, scheduler="cosine_warm_restarts"):
        if model.family not in ("webp", "jpeg"):
            raise NotImplementedError("training kernels cover the WebP/JPEG families (DCT frequency block) only")
        self.model = model
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_grad_norm, self.dropout_p, self.seed = max_grad_norm, dropout, seed
        self.step_count = 0
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("training runs on CUDA only (no CPU fallback)")
        self.params = {k: p for k, p in model.named_parameters()}
        n = sum(p.numel() for p in self.params.values())
        self.flat_grad = torch.zeros((n,), dtype=F32, device=dev)
        self.grads, off = {}, 0
        for k, p in self.params.items():
            self.grads[k] = self.flat_grad[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.m = {k: torch.zeros_like(p, dtype=F32) for k, p in self.params.items()}
        self.v = {k: torch.zeros_like(p, dtype=F32) for k, p in self.params.items()}
        self._norm_acc = torch.zeros((1,), dtype=torch.float64, device=dev)
        # gradient buckets for the data-parallel all-reduce: the backward finishes the blocks in reverse registration order,
        # so every finished block closes a contiguous tail [start(block), previous start) of the flat buffer
        self.overlap_allreduce = overlap_allreduce
        self._span_start = grad_span_starts(model)
        self.buckets = parallel.GradBuckets(self.flat_grad, bucket_bytes)
        self._overlap_now = False
        # per-epoch lr schedule of the reference (webp_training.py:776): CosineAnnealingWarmRestarts(T_0=100, T_mult=2);
        # `scheduler` may also be None (constant lr) or any object with step() -> lr
        self.scheduler = parallel.CosineWarmRestarts(lr) if scheduler == "cosine_warm_restarts" else scheduler
        import torch.distributed as dist
        self.rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0

    # ------------------------------------------------------------------------------------------------------------
    def _pack(self):
        """Per-step weight packs (weights change every step): forward [N,(kh,kw,cin)] and backward (transposed,
        tap-flipped) operands in the compute dtype."""
        m = self.model
        dt = torch.bfloat16 if m.precision == "bf16" else F32
        sd = {k: v.detach() for k, v in m.state_dict().items()}

        def cast(w):
            w = w.contiguous().float()
            return ops.cast_bf16(w) if dt == torch.bfloat16 else w
        c3 = lambda w: cast(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
        c3t = lambda w: cast(w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1))   # dgrad: [Cin, 9*Cout]
        l1 = lambda w: cast(w.reshape(w.shape[0], -1))
        l1t = lambda w: cast(w.reshape(w.shape[0], -1).t())
        P = {}
        for p, ci, co in _BLOCKS:
            q = {}
            if ci != 3:
                q["conv1"], q["conv1_t"] = c3(sd[f"{p}.conv1.weight"]), c3t(sd[f"{p}.conv1.weight"])
                if ci != co:
                    q["sc"], q["sc_t"] = l1(sd[f"{p}.shortcut.weight"]), l1t(sd[f"{p}.shortcut.weight"])
            q["conv2"], q["conv2_t"] = c3(sd[f"{p}.conv2.weight"]), c3t(sd[f"{p}.conv2.weight"])
            q["in"], q["in_t"] = l1(sd[f"{p}.attn.in_proj_weight"]), l1t(sd[f"{p}.attn.in_proj_weight"])
            q["out"], q["out_t"] = l1(sd[f"{p}.attn.out_proj.weight"]), l1t(sd[f"{p}.attn.out_proj.weight"])
            f = f"{p}.freq_guide"
            w1 = torch.cat([sd[f"{f}.low_freq_attn.0.weight"], sd[f"{f}.high_freq_attn.0.weight"]], 0).reshape(co, co)
            w2 = torch.cat([sd[f"{f}.low_freq_attn.2.weight"].reshape(co, co // 2), sd[f"{f}.high_freq_attn.2.weight"].reshape(co, co // 2)], 1)
            q["g1"], q["g1_t"] = cast(w1), cast(w1.t())
            q["g2"], q["g2_t"] = cast(w2), cast(w2.t())
            q["g1_b"] = torch.cat([sd[f"{f}.low_freq_attn.0.bias"], sd[f"{f}.high_freq_attn.0.bias"]], 0).contiguous().float()
            q["fo"], q["fo_t"] = c3(sd[f"{f}.conv_out.weight"]), c3t(sd[f"{f}.conv_out.weight"])
            P[p] = q
        return P, dt

    def _op(self, x, dt):
        """fp32 gradient/stream tensor -> GEMM operand dtype."""
        return ops.cast_bf16(x) if dt == torch.bfloat16 else x

    # ------------------------------------------------------------------------------------------------------------
    def scheduler_step(self):
        """End of an epoch: advance the lr schedule (webp_training.py:530) and return the new lr."""
        if self.scheduler is not None:
            self.lr = float(self.scheduler.step())
        return self.lr

    def forward_backward(self, xt, t, x0, dropout_seed=None, overlap=False):
        """One forward + backward of  frequency_aware_loss(xt + model(xt, t, t), x0).  Fills self.grads; returns loss.
        overlap=True (what train_step passes) lets the backward launch the bucketed gradient all-reduces itself; the caller
        then MUST call allreduce_grads() before touching the gradients.  Standalone calls leave the gradients local."""
        m = self.model
        fam = _FAMILY[m.family]
        sd = dict(m.named_parameters()); sd.update(dict(m.named_buffers()))
        sd = {k: v.detach() for k, v in sd.items()}
        P, dt = self._pack()
        impl = m.impl
        G = self.grads
        self.buckets.reset()          # raises if a previous overlapped backward was never finished
        self._overlap_now = bool(overlap and self.overlap_allreduce)
        self.flat_grad.zero_()
        xt = xt.contiguous().float(); x0 = x0.contiguous().float(); t = t.contiguous().float()
        B = xt.shape[0]
        p_drop = self.dropout_p
        # every rank draws its own dropout masks (the ranks hold different images)
        dseed = ((self.seed * 1000003 + self.step_count) * 4099 + self.rank) if dropout_seed is None else dropout_seed
        with torch.no_grad():
            # ---------------- forward ----------------
            # TimeEmbedding with the pre-activation kept: features -> Linear -> SiLU -> Linear
            feat = T.time_features(t)
            u1 = ops.linear_rows(feat, sd["time_embed.proj.0.weight"], sd["time_embed.proj.0.bias"])
            hmid = T.act_forward(u1, ops.ACT_SILU)
            t_emb = ops.linear_rows(hmid, sd["time_embed.proj.2.weight"], sd["time_embed.proj.2.bias"])
            boost = torch.clamp(1.0 - t, fam["clamp"][0], fam["clamp"][1]).contiguous()
            tapes = {}

            def blk(p, z, idx):
                out, tape = self._block_fwd(p, z, t_emb, boost, sd, P[p], dt, impl, p_drop, dseed * 64 + idx)
                tapes[p] = tape
                return out
            d1 = blk("down1", xt, 0)
            d2 = blk("down2", ops.maxpool2(d1), 1)
            d3 = blk("down3", ops.maxpool2(d2), 2)
            d4 = blk("down4", ops.maxpool2(d3), 3)
            d5 = blk("down5", ops.maxpool2(d4), 4)
            b0 = blk("bottleneck.0", ops.maxpool2(d5), 5)
            b1 = blk("bottleneck.1", b0, 6)
            b2 = blk("bottleneck.2", b1, 7)
            u1_ = blk("up1", ops.upsample2_concat(b2, d5), 8)
            u2 = blk("up2", ops.upsample2_concat(u1_, d4), 9)
            u3 = blk("up3", ops.upsample2_concat(u2, d3), 10)
            u4 = blk("up4", ops.upsample2_concat(u3, d2), 11)
            u5 = blk("up5", ops.upsample2_concat(u4, d1), 12)
            Dm = sd["dct_layer.dct_matrix"]
            comb = ops.block_transform(u5, Dm, 1.0, fam["tail"])
            st_t = ops.groupnorm_stats(comb, 8)
            a_t = ops.groupnorm_apply(comb, st_t, sd["out_conv.0.weight"], sd["out_conv.0.bias"], ops.ACT_SILU, out_dtype=dt)
            pred = ops.out_conv_tanh(a_t, sd["out_conv.2.weight"], sd["out_conv.2.bias"])
            recon = ops.lincomb(xt, 1.0, pred, 1.0)                       # xt + pred, webp_training.py:515
            from .losses import frequency_aware_loss
            loss = frequency_aware_loss(recon, x0)

            # ---------------- backward ----------------
            dpred = T.frequency_aware_loss_backward(recon, x0)
            da = T.out_conv_tanh_backward(a_t, pred, dpred, sd["out_conv.2.weight"], G["out_conv.2.weight"], G["out_conv.2.bias"])
            dcomb = T.groupnorm_backward(comb, da, st_t, sd["out_conv.0.weight"], sd["out_conv.0.bias"], ops.ACT_SILU,
                                         G["out_conv.0.weight"], G["out_conv.0.bias"])
            du5 = ops.block_transform(dcomb, Dm.t().contiguous(), 1.0, fam["tail"])
            dt_emb = torch.zeros_like(t_emb)

            def bwd(p, dout):
                r = self._block_bwd(p, dout, tapes.pop(p), t_emb, dt_emb, boost, sd, P[p], dt, impl, p_drop)
                self._grads_final_from(self._span_start[p])
                return r
            dcat = bwd("up5", du5)
            du4, dd1 = T.upsample2_concat_backward(dcat, 64)
            dcat = bwd("up4", du4)
            du3, dd2 = T.upsample2_concat_backward(dcat, 128)
            dcat = bwd("up3", du3)
            du2, dd3 = T.upsample2_concat_backward(dcat, 256)
            dcat = bwd("up2", du2)
            du1, dd4 = T.upsample2_concat_backward(dcat, 512)
            dcat = bwd("up1", du1)
            db2, dd5 = T.upsample2_concat_backward(dcat, 512)
            db1 = bwd("bottleneck.2", db2)
            db0 = bwd("bottleneck.1", db1)
            dp5 = bwd("bottleneck.0", db0)
            add = lambda a, b_: ops.lincomb(a, 1.0, b_, 1.0)
            dd5 = add(dd5, T.maxpool2_backward(d5, dp5))
            dp4 = bwd("down5", dd5)
            dd4 = add(dd4, T.maxpool2_backward(d4, dp4))
            dp3 = bwd("down4", dd4)
            dd3 = add(dd3, T.maxpool2_backward(d3, dp3))
            dp2 = bwd("down3", dd3)
            dd2 = add(dd2, T.maxpool2_backward(d2, dp2))
            dp1 = bwd("down2", dd2)
            dd1 = add(dd1, T.maxpool2_backward(d1, dp1))
            bwd("down1", dd1)
            # time-embedding MLP
            dh = T.linear_rows_backward(dt_emb, hmid, sd["time_embed.proj.2.weight"], G["time_embed.proj.2.weight"],
                                        G["time_embed.proj.2.bias"])
            du = T.act_backward(dh, u1, ops.ACT_SILU)
            T.linear_rows_backward(du, feat, sd["time_embed.proj.0.weight"], G["time_embed.proj.0.weight"],
                                   G["time_embed.proj.0.bias"], need_dx=False)
        return loss

    # ------------------------------------------------------------------------------------------------------------
    def _block_fwd(self, p, x, t_emb, boost, sd, W, dt, impl, p_drop, dseed):
        fam = _FAMILY[self.model.family]
        first = p == "down1"
        co = sd[f"{p}.conv1.bias"].shape[0]
        tp = dict(first=first)
        tb = ops.linear_rows(t_emb, sd[f"{p}.time_proj.weight"], sd[f"{p}.time_proj.bias"])
        if first:
            st1 = ops.groupnorm_stats(x, 3, nchw=True)
            h1 = ops.conv_input(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], F32, st1, sd[f"{p}.norm1.weight"],
                                sd[f"{p}.norm1.bias"], row_bias=tb)
            sc = ops.conv_input(x, sd[f"{p}.shortcut.weight"], sd[f"{p}.shortcut.bias"], F32)
            tp.update(x=x, st1=st1)
        else:
            ci = x.shape[-1]
            st1 = ops.groupnorm_stats(x, _groups(ci))
            if "sc" in W:
                a, x_op = ops.groupnorm_apply(x, st1, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE, out_dtype=dt,
                                              raw_copy=True)
                sc = ops.gemm(x_op, W["sc"], co, impl, out_dtype=F32, bias=sd[f"{p}.shortcut.bias"])
                tp.update(x_op=x_op)
            else:
                a = ops.groupnorm_apply(x, st1, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE, out_dtype=dt)
                sc = x
            h1 = ops.conv3x3(a, W["conv1"], co, impl, out_dtype=F32, bias=sd[f"{p}.conv1.bias"], row_bias=tb)
            tp.update(x=x, st1=st1, a=a)
        st2 = ops.groupnorm_stats(h1, _groups(co))
        a2 = ops.groupnorm_apply(h1, st2, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], ops.ACT_GELU, out_dtype=dt)
        a2d = T.dropout(a2, p_drop, dseed) if p_drop > 0 else a2
        h2, h2_op = ops.conv3x3(a2d, W["conv2"], co, impl, out_dtype=F32, out2_dtype=dt, bias=sd[f"{p}.conv2.bias"])
        Bn, H, Wd, _ = h2.shape
        qkv = ops.gemm(h2_op, W["in"], 3 * co, impl, bias=sd[f"{p}.attn.in_proj_bias"])
        ao, lse = T.attention_train_forward(qkv.view(Bn, H * Wd, 3 * co), fam["heads"])
        ao = ao.view(Bn, H, Wd, co)
        h3 = ops.gemm(ao, W["out"], co, impl, out_dtype=F32, bias=sd[f"{p}.attn.out_proj.bias"], res=h2)
        f = f"{p}.freq_guide"
        d = ops.block_transform(h3, sd[f"{f}.dct.dct_matrix"], 0.0, 1.0, out_dtype=dt)
        g1 = ops.gemm(d, W["g1"], co, impl, bias=W["g1_b"], act=ops.ACT_LRELU02, freq_mode=1, bs=fam["bs"], low=fam["low"])
        # gate values g = sigmoid(z) kept for the backward (the inference path fuses g*s*d + h3 into this epilogue)
        g = ops.gemm(g1, W["g2"], co, impl, bias=sd[f"{f}.low_freq_attn.2.bias"], bias2=sd[f"{f}.high_freq_attn.2.bias"],
                     act=ops.ACT_SIGMOID, freq_mode=2, bs=fam["bs"], low=fam["low"])
        e = ops.gemm(g1, W["g2"], co, impl, bias=sd[f"{f}.low_freq_attn.2.bias"], bias2=sd[f"{f}.high_freq_attn.2.bias"],
                     act=ops.ACT_SIGMOID, freq_mode=2, bs=fam["bs"], low=fam["low"], img_scale=boost, mul=d, res=h3)
        out = ops.conv3x3(e, W["fo"], co, impl, out_dtype=F32, bias=sd[f"{f}.conv_out.bias"], res=sc)
        tp.update(h1=h1, st2=st2, a2d=a2d, dseed=dseed, h2_op=h2_op, qkv=qkv, ao=ao, lse=lse, d=d, g1=g1, g=g, e=e, co=co)
        return out, tp

    def _block_bwd(self, p, dout, tp, t_emb, dt_emb, boost, sd, W, dt, impl, p_drop):
        fam = _FAMILY[self.model.family]
        G = self.grads
        co = tp["co"]
        bs, low = fam["bs"], fam["low"]
        f = f"{p}.freq_guide"
        op = lambda z: self._op(z, dt)
        Bn, H, Wd, _ = dout.shape
        # out = conv_out(e) + sc.  Every gradient that feeds a GEMM is cast ONCE to the operand dtype (bf16 in production)
        # and that copy serves both the data-gradient GEMM and the tensor-core weight-gradient kernel.
        dout_op = op(dout)
        T.wgrad(dout_op, tp["e"], G[f"{f}.conv_out.weight"], 9, oihw=True)
        T.colsum(dout, G[f"{f}.conv_out.bias"])
        de = ops.conv3x3(dout_op, W["fo_t"], co, impl, out_dtype=F32)
        # e = h3 + g*s*d
        dz, dd = T.gate_backward(de, tp["g"], tp["d"], boost, bs, low)
        dz_op = op(dz)
        T.wgrad(dz_op, tp["g1"], G[f"{f}.low_freq_attn.2.weight"], 1, k_begin=0, k_count=co // 2, out_ld=co // 2)
        T.wgrad(dz_op, tp["g1"], G[f"{f}.high_freq_attn.2.weight"], 1, k_begin=co // 2, k_count=co // 2, out_ld=co // 2)
        T.colsum(dz, G[f"{f}.low_freq_attn.2.bias"], cls=1, bs=bs, low=low)
        T.colsum(dz, G[f"{f}.high_freq_attn.2.bias"], cls=0, bs=bs, low=low)
        dg1 = ops.gemm(dz_op, W["g2_t"], co, impl, out_dtype=F32)
        dpre = T.lrelu_mask_backward(dg1, tp["g1"], bs, low)
        dpre_op = op(dpre)
        T.wgrad(dpre_op, tp["d"], G[f"{f}.low_freq_attn.0.weight"], 1, n_begin=0, n_count=co // 2)
        T.wgrad(dpre_op, tp["d"], G[f"{f}.high_freq_attn.0.weight"], 1, n_begin=co // 2, n_count=co // 2)
        T.colsum(dpre, G[f"{f}.low_freq_attn.0.bias"], n_begin=0, n_count=co // 2)
        T.colsum(dpre, G[f"{f}.high_freq_attn.0.bias"], n_begin=co // 2, n_count=co // 2)
        dd = ops.gemm(dpre_op, W["g1_t"], co, impl, out_dtype=F32, res=dd)
        # h3 receives de directly and through d = DCT(h3)
        Dt = sd[f"{f}.dct.dct_matrix"].t().contiguous()
        dh3 = ops.lincomb(de, 1.0, ops.block_transform(dd, Dt, 0.0, 1.0), 1.0)
        # h3 = out_proj(ao) + h2
        dh3_op = op(dh3)
        T.wgrad(dh3_op, tp["ao"], G[f"{p}.attn.out_proj.weight"], 1)
        T.colsum(dh3, G[f"{p}.attn.out_proj.bias"])
        dao = ops.gemm(dh3_op, W["out_t"], co, impl, out_dtype=F32)
        dqkv = T.attention_backward(tp["qkv"].view(Bn, H * Wd, 3 * co), tp["ao"].view(Bn, H * Wd, co), dao.view(Bn, H * Wd, co),
                                    tp["lse"], fam["heads"]).view(Bn, H, Wd, 3 * co)
        dqkv_op = op(dqkv)
        T.wgrad(dqkv_op, tp["h2_op"], G[f"{p}.attn.in_proj_weight"], 1)
        T.colsum(dqkv, G[f"{p}.attn.in_proj_bias"])
        dh2 = ops.gemm(dqkv_op, W["in_t"], co, impl, out_dtype=F32, res=dh3)
        # h2 = conv2(dropout(gelu(gn2(h1))))
        dh2_op = op(dh2)
        T.wgrad(dh2_op, tp["a2d"], G[f"{p}.conv2.weight"], 9, oihw=True)
        T.colsum(dh2, G[f"{p}.conv2.bias"])
        da2 = ops.conv3x3(dh2_op, W["conv2_t"], co, impl, out_dtype=F32)
        if p_drop > 0:
            da2 = T.dropout(da2, p_drop, tp["dseed"])
        dh1 = T.groupnorm_backward(tp["h1"], da2, tp["st2"], sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], ops.ACT_GELU,
                                   G[f"{p}.norm2.weight"], G[f"{p}.norm2.bias"])
        # h1 = conv1(gn1(x)) + b + time_proj(t_emb)
        dtb = torch.zeros((Bn, co), dtype=F32, device=dout.device)
        T.colsum(dh1, G[f"{p}.conv1.bias"], dtb)
        T.linear_rows_backward(dtb, t_emb, sd[f"{p}.time_proj.weight"], G[f"{p}.time_proj.weight"], G[f"{p}.time_proj.bias"],
                               dx_accum=dt_emb)
        if tp["first"]:
            x = tp["x"]
            T.conv_input_backward(x, dh1, sd[f"{p}.conv1.weight"], G[f"{p}.conv1.weight"], tp["st1"], sd[f"{p}.norm1.weight"],
                                  sd[f"{p}.norm1.bias"], G[f"{p}.norm1.weight"], G[f"{p}.norm1.bias"])
            T.conv_input_backward(x, dout, sd[f"{p}.shortcut.weight"], G[f"{p}.shortcut.weight"])
            T.colsum(dout, G[f"{p}.shortcut.bias"])
            return None
        ci = tp["x"].shape[-1]
        dh1_op = op(dh1)
        T.wgrad(dh1_op, tp["a"], G[f"{p}.conv1.weight"], 9, oihw=True)
        da = ops.conv3x3(dh1_op, W["conv1_t"], ci, impl, out_dtype=F32)
        if "sc" in W:
            T.wgrad(dout_op, tp["x_op"], G[f"{p}.shortcut.weight"], 1)
            T.colsum(dout, G[f"{p}.shortcut.bias"])
            dx = ops.gemm(dout_op, W["sc_t"], ci, impl, out_dtype=F32)
        else:
            dx = dout.clone()
        return T.groupnorm_backward(tp["x"], da, tp["st1"], sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], ops.ACT_NONE,
                                    G[f"{p}.norm1.weight"], G[f"{p}.norm1.bias"], dx=dx, accumulate=True)

    # ------------------------------------------------------------------------------------------------------------
    def _grads_final_from(self, lo):
        """Called by the backward when every gradient at offsets >= lo is final.  In an overlapped step with several ranks
        the finished tail is all-reduced right away in buckets (parallel.GradBuckets), under the rest of the backward
        (SURVEY section 8e)."""
        if self._overlap_now:
            self.buckets.final_from(lo)

    def allreduce_grads(self):
        """Data-parallel gradient averaging over NCCL: whatever the backward has not already sent in buckets goes out as
        one more all-reduce of the head of the flat buffer; then one scale by 1 / world."""
        self.buckets.finish()
        self._overlap_now = False

    def optimizer_step(self):
        """clip_grad_norm_(max_grad_norm) + AdamW, webp_training.py:522-524."""
        self.step_count += 1
        self._norm_acc.zero_()
        T.sumsq(self.flat_grad, self._norm_acc)
        b1, b2 = self.betas
        for k, p in self.params.items():
            T.adamw_step(p.data, self.grads[k], self.m[k], self.v[k], self.lr, b1, b2, self.eps, self.wd, self.step_count,
                         self._norm_acc, self.max_grad_norm)
        self.model._packed = None      # inference weight packs are stale now

    def train_step(self, xt, t, x0, dropout_seed=None):
        loss = self.forward_backward(xt, t, x0, dropout_seed, overlap=True)
        self.allreduce_grads()
        self.optimizer_step()
        return loss

    def grad_norm(self):
        return float(torch.sqrt(self._norm_acc)[0])


def train_epoch_ddrm_webp(trainer, batches, quality_for_t=None):
    """Loop form of webp_training.py:476-537 over an iterable of (x0, xt, t) device batches (the host-side WebP
    compression of x0 at the sampled quality is the caller's data pipeline, as in the reference's loop body)."""
    total, n = 0.0, 0
    for x0, xt, t in batches:
        loss = trainer.train_step(xt, t.float() / 100.0, x0)
        total += float(loss); n += 1
    trainer.scheduler_step()          # scheduler.step() once per epoch, webp_training.py:530
    return total / max(n, 1)
