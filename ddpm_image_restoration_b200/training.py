"""Training step of the reference (`train_epoch_ddrm_webp`, webp_training.py:476-537) on the B200 kernels:

    pred = model(xt, t/100, t/100)                  (train mode: Dropout(0.1) active, webp_inference.py:291,313)
    loss = frequency_aware_loss(xt + pred, x0)      webp_training.py:515-518
    zero_grad; backward; clip_grad_norm_(1.0); AdamW(lr 2e-4, wd 1e-5, betas (0.9, 0.99)).step()   :521-524, 775

There is no autograd here: `Trainer` runs an explicit forward that records what the backward needs, then the
hand-written backward kernels (csrc/backward*.cu, attn_simt.cu, loss.cu, fft.cu), then a fused clip+AdamW kernel per
tensor.  All gradients live in ONE flat fp32 buffer (parameters' gradients are views into it), so data-parallel training
is a single NCCL all-reduce of that buffer over NVLink (the only collective of the whole design).

Model families: WebP / JPEG (DCT frequency block) and AVIF (`train_epoch_ddrm_avif`, avif.py:528-590: learned per-channel
transform, multi-scale / colour / edge gates -- csrc/backward_avif.cu -- 8 heads, avif_frequency_aware_loss, AdamW lr 1.5e-4).
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
    def __init__(self, model, lr=None, weight_decay=1e-5, betas=(0.9, 0.99), eps=1e-8, max_grad_norm=1.0, dropout=0.1,
                 seed=0, overlap_allreduce=True, bucket_bytes=64 << 20, scheduler="cosine_warm_restarts"):
        if model.family not in ("webp", "jpeg", "avif"):
            raise NotImplementedError(f"no training kernels for the {model.family!r} family")
        if lr is None:          # webp_training.py:775 (2e-4); avif.py:796 (1.5e-4)
            lr = 1.5e-4 if model.family == "avif" else 2e-4
        self.model = model
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.max_grad_norm, self.dropout_p, self.seed = max_grad_norm, dropout, seed
        self.step_count = 0
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("training runs on CUDA only (no CPU fallback)")
        self.params = {k: p for k, p in model.named_parameters()}
        # AVIFAdaptiveTransform.inverse_weights is registered but never used in forward (avif.py:192): torch leaves its .grad
        # None, so clip_grad_norm_ and AdamW (weight decay included) skip it -- so does optimizer_step below
        self.unused = {k for k in self.params if k.endswith(".inverse_weights")}
        n = sum(p.numel() for p in self.params.values())
        self.flat_grad = torch.zeros((n,), dtype=F32, device=dev)
        self.grads, off = {}, 0
        for k, p in self.params.items():
            self.grads[k] = self.flat_grad[off:off + p.numel()].view(p.shape)
            off += p.numel()
        # AdamW moments: flat like the gradient (one multi-tensor launch updates everything, ops_train.adamw_multi)
        self.flat_m, self.flat_v = torch.zeros_like(self.flat_grad), torch.zeros_like(self.flat_grad)
        self.m, self.v, off = {}, {}, 0
        for k, p in self.params.items():
            self.m[k], self.v[k] = self.flat_m[off:off + p.numel()].view(p.shape), self.flat_v[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self._chunks = None
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
        tap-flipped) operands in the compute dtype, written by ddpmir_pack_weight (one pass per checkpoint tensor) into buffers
        that live as long as the Trainer."""
        m = self.model
        dt = torch.bfloat16 if m.precision == "bf16" else F32
        sd = {k: v.detach() for k, v in m.named_parameters()}
        bufs = self._pack_bufs if getattr(self, "_pack_dt", None) == dt else None
        fresh = bufs is None
        if fresh:
            bufs, self._pack_dt = {}, dt
        dev = next(m.parameters()).device

        def buf(key, *shape):
            if key not in bufs:
                bufs[key] = torch.empty(shape, dtype=dt, device=dev)
            return bufs[key]

        def both(q, p, key, name):
            """conv or linear weight -> q[key] (forward operand) and q[key + "_t"] (data-gradient operand)."""
            w = sd[name]
            N, Cin = w.shape[0], w.shape[1]
            taps = w.numel() // (N * Cin)
            q[key], q[key + "_t"] = buf((p, key), N, taps * Cin), buf((p, key, "t"), Cin, taps * N)
            T.pack_weight(w, q[key], q[key + "_t"])

        P = {}
        avif = m.family == "avif"
        for p, ci, co in _BLOCKS:
            q = {}
            if ci != 3:
                both(q, p, "conv1", f"{p}.conv1.weight")
                if ci != co:
                    both(q, p, "sc", f"{p}.shortcut.weight")
            both(q, p, "conv2", f"{p}.conv2.weight")
            both(q, p, "in", f"{p}.attn.in_proj_weight")
            both(q, p, "out", f"{p}.attn.out_proj.weight")
            f = f"{p}.freq_guide"
            if avif:
                a = f"{f}.adaptive_transform"
                for key, name in (("q0", f"{a}.quantization.0"), ("q2", f"{a}.quantization.2"), ("c0", f"{f}.color_consistency.0"),
                                  ("c2", f"{f}.color_consistency.2"), ("e0", f"{f}.edge_preserve.0"), ("e2", f"{f}.edge_preserve.2")):
                    both(q, p, key, name + ".weight")
                q["Tt"] = sd[f"{a}.transform_weights"].transpose(1, 2).contiguous().float()
            else:
                # stacked low/high gate MLP: g1 = [W1_low; W1_high] (rows), g2 = [W2_low | W2_high] (columns), and their transposes
                h = co // 2
                g1, g1t, g2, g2t = buf((p, "g1"), co, co), buf((p, "g1t"), co, co), buf((p, "g2"), co, co), buf((p, "g2t"), co, co)
                T.pack_weight(sd[f"{f}.low_freq_attn.0.weight"], g1[:h], g1t, bwd_ld=co, bwd_off=0)
                T.pack_weight(sd[f"{f}.high_freq_attn.0.weight"], g1[h:], g1t, bwd_ld=co, bwd_off=h)
                T.pack_weight(sd[f"{f}.low_freq_attn.2.weight"], g2, g2t[:h], fwd_ld=co, fwd_off=0)
                T.pack_weight(sd[f"{f}.high_freq_attn.2.weight"], g2, g2t[h:], fwd_ld=co, fwd_off=h)
                q["g1"], q["g1_t"], q["g2"], q["g2_t"] = g1, g1t, g2, g2t
                q["g1_b"] = torch.cat([sd[f"{f}.low_freq_attn.0.bias"], sd[f"{f}.high_freq_attn.0.bias"]], 0).contiguous().float()
            both(q, p, "fo", f"{f}.conv_out.weight")
            P[p] = q
        if avif:
            q = {}
            both(q, "tail", "q0", "avif_layer.quantization.0.weight")
            both(q, "tail", "q2", "avif_layer.quantization.2.weight")
            q["Tt"] = sd["avif_layer.transform_weights"].transpose(1, 2).contiguous().float()
            P["tail"] = q
        self._pack_bufs = bufs
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
            avif = m.family == "avif"
            if avif:      # colour / edge boosts of AVIFFreqAwareBlock (avif.py:309-310); compression_level = t (avif.py:566)
                boost = (torch.clamp(0.5 + 0.5 * (1.0 - t), 0.3, 1.5).contiguous(), torch.clamp(0.7 + 0.3 * (1.0 - t), 0.5, 1.3).contiguous())
            else:
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
            if avif:      # u5 + 0.15 * avif_layer(u5), avif_inference.py:383
                tailv = torch.full((B,), fam["tail"], dtype=F32, device=xt.device)
                comb, tail_tape = self._tgate_fwd("avif_layer", u5, sd, P["tail"], dt, impl, scale=tailv, res=u5)
            else:
                Dm = sd["dct_layer.dct_matrix"]
                comb = ops.block_transform(u5, Dm, 1.0, fam["tail"])
            st_t = ops.groupnorm_stats(comb, 8)
            a_t = ops.groupnorm_apply(comb, st_t, sd["out_conv.0.weight"], sd["out_conv.0.bias"], ops.ACT_SILU, out_dtype=dt)
            pred = ops.out_conv_tanh(a_t, sd["out_conv.2.weight"], sd["out_conv.2.bias"])
            recon = ops.lincomb(xt, 1.0, pred, 1.0)                       # xt + pred, webp_training.py:515 / avif.py:570
            from . import losses
            loss = (losses.avif_frequency_aware_loss if avif else losses.frequency_aware_loss)(recon, x0)

            # ---------------- backward ----------------
            dpred = (T.avif_frequency_aware_loss_backward if avif else T.frequency_aware_loss_backward)(recon, x0)
            da = T.out_conv_tanh_backward(a_t, pred, dpred, sd["out_conv.2.weight"], G["out_conv.2.weight"], G["out_conv.2.bias"])
            dcomb = T.groupnorm_backward(comb, da, st_t, sd["out_conv.0.weight"], sd["out_conv.0.bias"], ops.ACT_SILU,
                                         G["out_conv.0.weight"], G["out_conv.0.bias"])
            if avif:
                dxt_tail = ops.lincomb(dcomb, fam["tail"])
                du5 = self._tgate_bwd("avif_layer", dxt_tail, tail_tape, sd, P["tail"], dt, impl, res=dcomb)
            else:
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
        f = f"{p}.freq_guide"
        if self.model.family == "avif":
            h3, h3_op = ops.gemm(ao, W["out"], co, impl, out_dtype=F32, out2_dtype=dt, bias=sd[f"{p}.attn.out_proj.bias"], res=h2)
            e = self._avif_freq_fwd(f, h3, h3_op, boost, sd, W, dt, impl, tp)
        else:
            h3 = ops.gemm(ao, W["out"], co, impl, out_dtype=F32, bias=sd[f"{p}.attn.out_proj.bias"], res=h2)
            d = ops.block_transform(h3, sd[f"{f}.dct.dct_matrix"], 0.0, 1.0, out_dtype=dt)
            g1 = ops.gemm(d, W["g1"], co, impl, bias=W["g1_b"], act=ops.ACT_LRELU02, freq_mode=1, bs=fam["bs"], low=fam["low"])
            # gate values g = sigmoid(z) kept for the backward (the inference path fuses g*s*d + h3 into this epilogue)
            g = ops.gemm(g1, W["g2"], co, impl, bias=sd[f"{f}.low_freq_attn.2.bias"], bias2=sd[f"{f}.high_freq_attn.2.bias"],
                         act=ops.ACT_SIGMOID, freq_mode=2, bs=fam["bs"], low=fam["low"])
            e = ops.gemm(g1, W["g2"], co, impl, bias=sd[f"{f}.low_freq_attn.2.bias"], bias2=sd[f"{f}.high_freq_attn.2.bias"],
                         act=ops.ACT_SIGMOID, freq_mode=2, bs=fam["bs"], low=fam["low"], img_scale=boost, mul=d, res=h3)
            tp.update(d=d, g1=g1, g=g)
        out = ops.conv3x3(e, W["fo"], co, impl, out_dtype=F32, bias=sd[f"{f}.conv_out.bias"], res=sc)
        tp.update(h1=h1, st2=st2, a2d=a2d, dseed=dseed, h2_op=h2_op, qkv=qkv, ao=ao, lse=lse, e=e, co=co)
        return out, tp

    # ---- AVIF family: learned transform with its sigmoid "quantisation" gate (AVIFAdaptiveTransform, avif.py:185-243) ---------
    def _tgate_fwd(self, a, x, sd, W, dt, impl, scale=None, res=None):
        """x (fp32 stream) -> xt = tr * sigmoid(q2(relu(q0(tr)))), tr = T_c X T_c^T;  returns (xt in dt, tape), or with
        `res` (and a per-image `scale`) the fp32 stream tensor res + scale * xt (the UNet tail, avif_inference.py:383)."""
        co = x.shape[-1]
        Tw = sd[f"{a}.transform_weights"]
        tr = ops.block_transform(x, Tw, 0.0, 1.0, out_dtype=dt)
        q1 = ops.gemm(tr, W["q0"], co, impl, bias=sd[f"{a}.quantization.0.bias"], act=ops.ACT_RELU)
        g = ops.gemm(q1, W["q2"], co, impl, bias=sd[f"{a}.quantization.2.bias"], act=ops.ACT_SIGMOID)       # gate kept for the backward
        xt = ops.gemm(q1, W["q2"], co, impl, out_dtype=None if res is None else F32, bias=sd[f"{a}.quantization.2.bias"],
                      act=ops.ACT_SIGMOID, img_scale=scale, mul=tr, res=res)
        return xt, dict(x=x, tr=tr, q1=q1, g=g)

    def _tgate_bwd(self, a, dxt, tp, sd, W, dt, impl, res):
        """Gradient of _tgate_fwd: parameter gradients into self.grads, returns res + d x."""
        G = self.grads
        op = lambda z: self._op(z, dt)
        co = dxt.shape[-1]
        ones = self._ones(dxt.shape[0], dxt.device)
        dz, dtr = T.gate_backward(dxt, tp["g"], tp["tr"], ones, 1, 1)          # xt = g * tr: dz = dxt*tr*g(1-g), dtr = dxt*g
        dz_op = op(dz)
        T.wgrad(dz_op, tp["q1"], G[f"{a}.quantization.2.weight"], 1)
        T.colsum(dz, G[f"{a}.quantization.2.bias"])
        dq1 = ops.gemm(dz_op, W["q2_t"], co, impl, out_dtype=F32)
        dpre = T.relu_mask_backward(dq1, tp["q1"])
        dpre_op = op(dpre)
        T.wgrad(dpre_op, tp["tr"], G[f"{a}.quantization.0.weight"], 1)
        T.colsum(dpre, G[f"{a}.quantization.0.bias"])
        dtr = ops.gemm(dpre_op, W["q0_t"], co, impl, out_dtype=F32, res=dtr)
        T.block_transform_wgrad(tp["x"], dtr, sd[f"{a}.transform_weights"], G[f"{a}.transform_weights"])
        return ops.lincomb(res, 1.0, ops.block_transform(dtr, W["Tt"], 0.0, 1.0), 1.0)

    def _ones(self, n, dev):
        if getattr(self, "_ones_buf", None) is None or self._ones_buf.shape[0] < n:
            self._ones_buf = torch.ones((max(n, 64),), dtype=F32, device=dev)
        return self._ones_buf

    def _avif_freq_fwd(self, f, h3, h3_op, boost, sd, W, dt, impl, tp):
        """AVIFFreqAwareBlock.forward up to `x + enhanced` (avif.py:284-321) with the tape of its backward."""
        Bn, _, _, co = h3.shape
        xt, tg = self._tgate_fwd(f"{f}.adaptive_transform", h3, sd, W, dt, impl)
        pooled = ops.avgpool_pyramid(h3)                       # [85, B, C] fp32: AdaptiveAvgPool2d(1 | 2 | 4 | 8)
        gates = torch.empty_like(pooled)
        ms, off = [], 0
        for i, s in enumerate((1, 2, 4, 8)):
            rows = pooled[off:off + s * s].reshape(s * s * Bn, co)
            w1 = sd[f"{f}.multi_scale_attn.{i}.1.weight"].reshape(co // 4, co)
            w3 = sd[f"{f}.multi_scale_attn.{i}.3.weight"].reshape(co, co // 4)
            u = ops.linear_rows(rows, w1, sd[f"{f}.multi_scale_attn.{i}.1.bias"])
            hid = T.act_forward(u, ops.ACT_RELU)
            z = ops.linear_rows(hid, w3, sd[f"{f}.multi_scale_attn.{i}.3.bias"])
            T.act_forward(z, ops.ACT_SIGMOID, out=gates[off:off + s * s])
            ms.append((rows, u, hid, z))
            off += s * s
        c1 = ops.gemm(h3_op, W["c0"], co, impl, bias=sd[f"{f}.color_consistency.0.bias"], act=ops.ACT_RELU)
        color = ops.gemm(c1, W["c2"], co, impl, bias=sd[f"{f}.color_consistency.2.bias"], act=ops.ACT_SIGMOID, img_scale=boost[0])
        # hidden width C/2 = 32 at the C = 64 blocks is below the tcgen05 kernels' 64-wide K block / N tile: those two convs (and
        # their data gradients) take the generic kernel (inference pads the hidden layer to 64 instead, models.py::prepack)
        impl_e = ops.IMPL_SIMT if co // 2 < 64 else impl
        e1 = ops.conv3x3(h3_op, W["e0"], co // 2, impl_e, bias=sd[f"{f}.edge_preserve.0.bias"], act=ops.ACT_RELU)
        edge = ops.conv3x3(e1, W["e2"], co, impl_e, bias=sd[f"{f}.edge_preserve.2.bias"], act=ops.ACT_SIGMOID, img_scale=boost[1])
        tp.update(tg=tg, xt=xt, gates=gates, ms=ms, h3_op=h3_op, c1=c1, color=color, e1=e1, edge=edge)
        return ops.avif_combine(h3, xt, gates, color, edge)

    def _avif_freq_bwd(self, f, de, tp, boost, sd, W, dt, impl):
        """de = gradient at `x + enhanced` -> gradient at the block's h3 (all parameter gradients of the frequency block)."""
        G = self.grads
        op = lambda z: self._op(z, dt)
        Bn, _, _, co = de.shape
        dxt, dzc, dze, dattn = T.avif_combine_backward(de, tp["xt"], tp["gates"], tp["color"], tp["edge"], boost[0], boost[1])
        # edge gate: sigmoid(conv3x3(relu(conv3x3(h3))))
        dze_op = op(dze)
        T.wgrad(dze_op, tp["e1"], G[f"{f}.edge_preserve.2.weight"], 9, oihw=True)
        T.colsum(dze, G[f"{f}.edge_preserve.2.bias"])
        impl_e = ops.IMPL_SIMT if co // 2 < 64 else impl
        de1 = ops.conv3x3(dze_op, W["e2_t"], co // 2, impl_e, out_dtype=F32)
        dpe = T.relu_mask_backward(de1, tp["e1"])
        dpe_op = op(dpe)
        T.wgrad(dpe_op, tp["h3_op"], G[f"{f}.edge_preserve.0.weight"], 9, oihw=True)
        T.colsum(dpe, G[f"{f}.edge_preserve.0.bias"])
        dh3 = ops.conv3x3(dpe_op, W["e0_t"], co, impl_e, out_dtype=F32, res=de)      # + the direct path  x + enhanced
        # colour gate: sigmoid(1x1(relu(1x1(h3))))
        dzc_op = op(dzc)
        T.wgrad(dzc_op, tp["c1"], G[f"{f}.color_consistency.2.weight"], 1)
        T.colsum(dzc, G[f"{f}.color_consistency.2.bias"])
        dc1 = ops.gemm(dzc_op, W["c2_t"], co, impl, out_dtype=F32)
        dpc = T.relu_mask_backward(dc1, tp["c1"])
        dpc_op = op(dpc)
        T.wgrad(dpc_op, tp["h3_op"], G[f"{f}.color_consistency.0.weight"], 1)
        T.colsum(dpc, G[f"{f}.color_consistency.0.bias"])
        dh3 = ops.gemm(dpc_op, W["c0_t"], co, impl, out_dtype=F32, res=dh3)
        # multi-scale gates: bilinear up-sampling and adaptive pooling transposed, the four small MLPs in between
        dgates = T.avif_gates_backward(dattn)
        dpooled = torch.empty_like(dgates)
        off = 0
        for i, s in enumerate((1, 2, 4, 8)):
            rows, u, hid, z = tp["ms"][i]
            w1 = sd[f"{f}.multi_scale_attn.{i}.1.weight"].reshape(co // 4, co)
            w3 = sd[f"{f}.multi_scale_attn.{i}.3.weight"].reshape(co, co // 4)
            dz = T.act_backward(dgates[off:off + s * s].reshape(s * s * Bn, co), z, ops.ACT_SIGMOID)
            dhid = T.linear_rows_backward(dz, hid, w3, G[f"{f}.multi_scale_attn.{i}.3.weight"], G[f"{f}.multi_scale_attn.{i}.3.bias"])
            du = T.act_backward(dhid, u, ops.ACT_RELU)
            T.linear_rows_backward(du, rows, w1, G[f"{f}.multi_scale_attn.{i}.1.weight"], G[f"{f}.multi_scale_attn.{i}.1.bias"],
                                   dx_accum=dpooled[off:off + s * s].view(s * s * Bn, co).zero_())
            off += s * s
        T.avgpool_pyramid_backward(dpooled, dh3)
        # transform branch
        return self._tgate_bwd(f"{f}.adaptive_transform", dxt, tp["tg"], sd, W, dt, impl, res=dh3)

    def _block_bwd(self, p, dout, tp, t_emb, dt_emb, boost, sd, W, dt, impl, p_drop):
        fam = _FAMILY[self.model.family]
        G = self.grads
        co = tp["co"]
        bs, low = fam["bs"], fam.get("low", 0)
        f = f"{p}.freq_guide"
        op = lambda z: self._op(z, dt)
        Bn, H, Wd, _ = dout.shape
        # out = conv_out(e) + sc.  Every gradient that feeds a GEMM is cast ONCE to the operand dtype (bf16 in production)
        # and that copy serves both the data-gradient GEMM and the tensor-core weight-gradient kernel.
        dout_op = op(dout)
        T.wgrad(dout_op, tp["e"], G[f"{f}.conv_out.weight"], 9, oihw=True)
        T.colsum(dout, G[f"{f}.conv_out.bias"])
        de = ops.conv3x3(dout_op, W["fo_t"], co, impl, out_dtype=F32)
        if self.model.family == "avif":
            dh3 = self._avif_freq_bwd(f, de, tp, boost, sd, W, dt, impl)
        else:
            dh3 = self._dct_freq_bwd(f, de, tp, boost, sd, W, dt, impl, bs, low)
        # h3 = out_proj(ao) + h2
        dh3_op = op(dh3)
        T.wgrad(dh3_op, tp["ao"], G[f"{p}.attn.out_proj.weight"], 1)
        T.colsum(dh3, G[f"{p}.attn.out_proj.bias"])
        # the GEMM epilogue writes the bf16 copy the tensor-core attention backward wants next to the fp32 gradient
        bf = dt == torch.bfloat16
        dao, dao_op = ops.gemm(dh3_op, W["out_t"], co, impl, out_dtype=F32, out2_dtype=dt) if bf else (ops.gemm(dh3_op, W["out_t"], co, impl, out_dtype=F32), None)
        dqkv = T.attention_backward(tp["qkv"].view(Bn, H * Wd, 3 * co), tp["ao"].view(Bn, H * Wd, co), dao.view(Bn, H * Wd, co),
                                    tp["lse"], fam["heads"], dout_op=dao_op).view(Bn, H, Wd, 3 * co)
        dqkv_op = op(dqkv)
        T.wgrad(dqkv_op, tp["h2_op"], G[f"{p}.attn.in_proj_weight"], 1)
        T.colsum(dqkv, G[f"{p}.attn.in_proj_bias"])
        # h2 = conv2(dropout(gelu(gn2(h1)))): dh2 only feeds GEMM-shaped consumers (weight / bias / data gradient of conv2), so in
        # bf16 mode it is produced in the operand dtype right away
        dh2_op = ops.gemm(dqkv_op, W["in_t"], co, impl, out_dtype=dt, res=dh3)
        T.wgrad(dh2_op, tp["a2d"], G[f"{p}.conv2.weight"], 9, oihw=True)
        T.colsum(dh2_op, G[f"{p}.conv2.bias"])
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

    def _dct_freq_bwd(self, f, de, tp, boost, sd, W, dt, impl, bs, low):
        """WebP/JPEG frequency block: de = gradient at e = h3 + g*s*d  ->  gradient at h3."""
        G = self.grads
        op = lambda z: self._op(z, dt)
        co = tp["co"]
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
        return dh3

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
        cs, cl, cp = self._chunk_table()
        T.adamw_multi(cs, cl, cp, self.flat_grad, self.flat_m, self.flat_v, self.lr, b1, b2, self.eps, self.wd, self.step_count,
                      self._norm_acc, self.max_grad_norm)
        self.model._packed = None      # inference weight packs are stale now

    CHUNK = 8192

    def _chunk_table(self):
        """(chunk_start, chunk_len, chunk_param) device tables of ops_train.adamw_multi: every parameter tensor cut into chunks of
        CHUNK elements; parameters without a gradient (self.unused) are left out.  Rebuilt if a parameter's storage moved."""
        ptrs = tuple(p.data_ptr() for p in self.params.values())
        if self._chunks is not None and self._chunks[0] == ptrs:
            return self._chunks[1]
        starts, lens, pp, off = [], [], [], 0
        for k, p in self.params.items():
            n = p.numel()
            if k not in self.unused:
                if not p.is_contiguous() or p.dtype != F32:
                    raise RuntimeError(f"parameter {k} must be a contiguous float32 tensor")
                for c in range(0, n, self.CHUNK):
                    starts.append(off + c); lens.append(min(self.CHUNK, n - c)); pp.append(p.data_ptr() + 4 * c)
            off += n
        dev = self.flat_grad.device
        tables = (torch.tensor(starts, dtype=torch.int64, device=dev), torch.tensor(lens, dtype=torch.int32, device=dev),
                  torch.tensor(pp, dtype=torch.int64, device=dev))
        self._chunks = (ptrs, tables)
        return tables

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


def train_epoch_ddrm_avif(trainer, batches):
    """Loop form of avif.py:528-590 (`train_epoch_ddrm_avif`) over an iterable of (x0, xt, t) device batches: pred =
    model(xt, t/100, t/100), loss = avif_frequency_aware_loss(xt + pred, x0), clip 1.0, AdamW(lr 1.5e-4, wd 1e-5,
    betas (0.9, 0.99)), scheduler.step() once per epoch.  The quality schedule and the host AVIF compression of x0
    (avif.py:539-560) are the caller's data pipeline, as in the reference's loop body."""
    if trainer.model.family != "avif":
        raise ValueError("train_epoch_ddrm_avif needs a Trainer over an AVIFDiffusionModel")
    return train_epoch_ddrm_webp(trainer, batches)
