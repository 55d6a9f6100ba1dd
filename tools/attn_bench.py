"""Times ddpmir_attention alone (tuning aid):  python tools/attn_bench.py [hd] [heads] [L] [B] [expmode] [pre]

`pre` = 1 times ddpmir_attention_prescaled (the bounded-softmax inference path); expmode (sel + 1) << 8 picks its exp split."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddpm_image_restoration_b200 import ops, _lib

hd = int(sys.argv[1]) if len(sys.argv) > 1 else 8
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
if len(sys.argv) > 5:
    _lib.lib().ddpmir_attention_set_expmode(int(sys.argv[5]))
pre = len(sys.argv) > 6 and int(sys.argv[6]) == 1
if "ATTN_LIN" in os.environ:      # largest polynomial set of the polynomial-kernel tier (attn_lin.cu), -1 = tier off
    _lib.lib().ddpmir_attention_set_lin(int(os.environ["ATTN_LIN"]))
attn = ops.attention_prescaled if pre else ops.attention
C = hd * heads
torch.manual_seed(0)
scale = float(os.environ.get("ATTN_SCALE", "1.0"))   # 0.35: logit bound < 2, 0.8: < 11 (half-precision tiers), 1.0+: bf16 tier
qkv = torch.randn(B, L, 3 * C, device="cuda") * scale
if pre:
    qkv[..., :C] *= 1.4426950408889634 / hd ** 0.5
qkv = qkv.to(torch.float16 if os.environ.get("ATTN_F16", "1") == "1" and pre else torch.bfloat16)
for _ in range(2):
    out = attn(qkv, heads)
torch.cuda.synchronize()
if pre and qkv.dtype == torch.float16:
    _, tiers = ops.attention_prescaled(qkv, heads, return_tiers=True)
    print("polynomial sets (-1 = quadratic tiers):", dict(zip(*[t.tolist() for t in tiers.flatten().unique(return_counts=True)])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    out = attn(qkv, heads)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
scores = B * heads * L * L
print(f"[scale {scale}] attention hd={hd} heads={heads} L={L} B={B}: {ms:.3f} ms  {scores / ms / 1e9:.3f} Tscores/s  "
      f"{scores / (ms * 1e-3) / 148 / 1.965e9:.2f} scores/clk/SM@1965MHz  {4.0 * scores * hd / ms / 1e9:.1f} TFLOP/s")
