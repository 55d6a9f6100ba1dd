"""Logit statistics of every attention call of one UNet forward (decides the windows of the attention tiers):
per call the Cauchy-Schwarz bound max|q'|*max|k| per (image, head), the same after centring q and k (softmax is invariant to
the k mean, the q mean becomes a per-key weight), and the largest logit actually seen on a sample of rows.

    python tools/attn_logit_stats.py [family] [B] [res] [default|keyed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpm_image_restoration_b200 as P
from ddpm_image_restoration_b200 import ops
import bench

fam = sys.argv[1] if len(sys.argv) > 1 else "avif"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
res = int(sys.argv[3]) if len(sys.argv) > 3 else 256
init = sys.argv[4] if len(sys.argv) > 4 else "default"
TAILS = os.environ.get("TAILS", "0") == "1"
torch.manual_seed(0)
model = {"avif": P.AVIFDiffusionModel, "webp": P.WebPDiffusionModel, "jpeg": P.JPEGDiffusionModel}[fam]()
if init == "keyed":
    from oracle import weights as W
    model.load_state_dict(W.make_state_dict(fam, 0))
model = model.cuda().eval().set_precision("bf16")
y = bench.synth_batch(fam, B, res, seed=1234).cuda()
orig = ops.attention_prescaled


def spy(qkv, heads):
    Bn, L, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    q = qkv[..., :C].float().view(Bn, L, heads, hd)
    k = qkv[..., C:2 * C].float().view(Bn, L, heads, hd)
    bound = q.norm(dim=-1).amax(1) * k.norm(dim=-1).amax(1)
    qc, kc = q - q.mean(1, keepdim=True), k - k.mean(1, keepdim=True)
    cbound = qc.norm(dim=-1).amax(1) * kc.norm(dim=-1).amax(1)
    kcb = q.norm(dim=-1).amax(1) * kc.norm(dim=-1).amax(1)
    idx = torch.randperm(L, device=qkv.device)[:512]
    s = torch.einsum("bihd,bjhd->bhij", q[:, idx], k)
    smax = s.abs().amax((2, 3))
    sc = torch.einsum("bihd,bjhd->bhij", qc[:, idx], kc).abs().amax((2, 3))
    f = lambda t: f"min {t.min():6.2f} med {t.median():6.2f} max {t.max():6.2f}"
    frac = lambda t, w: float((t <= w).float().mean())
    print(f"L={L:6d} C={C:4d} hd={hd:3d} {qkv.dtype}: bound {f(bound)} | k-centred {f(kcb)} | q,k-centred {f(cbound)} | seen |s| {f(smax)} "
          f"centred seen {f(sc)} | <=2: {frac(bound, 2):.2f} / {frac(kcb, 2):.2f} / {frac(cbound, 2):.2f}  <=3: {frac(bound, 3):.2f} / {frac(cbound, 3):.2f}"
          f"  <=4: {frac(bound, 4):.2f} / {frac(cbound, 4):.2f}")
    if TAILS and L >= 16384 and hd in (8, 16):
        # tails of the centred + balanced norms: how much of the Cauchy-Schwarz bound is a few outlier rows / keys, and where they sit
        D = ((kc * kc).mean(1, keepdim=True) / (qc * qc).mean(1, keepdim=True)).pow(0.25).clamp(1 / 16, 16)
        nq, nk = (qc * D).norm(dim=-1), (kc / D).norm(dim=-1)         # [B, L, heads]
        side = int(L ** 0.5)
        yy, xx = torch.meshgrid(torch.arange(side, device=qkv.device), torch.arange(side, device=qkv.device), indexing="ij")
        border = ((yy < 2) | (yy >= side - 2) | (xx < 2) | (xx >= side - 2)).flatten()
        for name, n in (("q", nq), ("k", nk)):
            srt = n.sort(dim=1).values
            pct = lambda p: srt[:, min(L - 1, int(p * L))]
            mxv = srt[:, -1]
            top = n.topk(64, dim=1).indices                       # [B, 64, heads]
            on_border = border[top].float().mean()
            print(f"      |{name}| balanced: max {f(mxv)} | p99.9 {f(pct(0.999))} | p99 {f(pct(0.99))} | median {f(pct(0.5))} | top-64 on the 2-px border: {on_border:.2f}")
        # bound if the top 0.1 % rows and keys were handled exactly
        b999 = nq.sort(dim=1).values[:, int(0.999 * L)] * nk.sort(dim=1).values[:, int(0.999 * L)]
        b99 = nq.sort(dim=1).values[:, int(0.99 * L)] * nk.sort(dim=1).values[:, int(0.99 * L)]
        print(f"      balanced bound {f(nq.amax(1) * nk.amax(1))} | without the top 0.1% rows+keys {f(b999)} | without the top 1% {f(b99)}")
    if qkv.dtype == torch.float16:
        out, tiers = orig(qkv, heads, return_tiers=True)
        print("      polynomial sets chosen (-1 = quadratic tiers):", dict(zip(*[t.tolist() for t in tiers.flatten().unique(return_counts=True)])))
        return out
    return orig(qkv, heads)


ops.attention_prescaled = spy
for tval in (0.9, 0.5, 0.1):
    print(f"--- t = {tval}")
    t = torch.full((B,), tval, device="cuda")
    with torch.no_grad():
        model(y, t, t)
