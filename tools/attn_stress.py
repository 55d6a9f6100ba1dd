"""Randomised shapes for ddpmir_attention_prescaled (tcgen05 kernel) against torch SDPA in fp32:  python tools/attn_stress.py"""
import sys, os, math, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ddpm_image_restoration_b200 import ops
random.seed(1)
bad = 0
for trial in range(60):
    hd = random.choice([8, 16]); heads = random.choice([1, 2, 4, 8]); B = random.choice([1, 2, 3, 5])
    L = 128 * random.randint(8, 96)
    C = hd * heads
    g = torch.Generator(device="cuda").manual_seed(trial)
    qkv = torch.randn(B, L, 3 * C, device="cuda", generator=g) * random.choice([0.5, 1.0, 2.0])
    c = 1.4426950408889634 / math.sqrt(hd)
    pre = qkv.clone(); pre[..., :C] *= c
    pre = pre.to(torch.bfloat16)
    ref_in = pre.float(); ref_in[..., :C] /= c
    q, k, v = ref_in.view(B, L, 3, heads, hd).permute(2, 0, 3, 1, 4)
    want = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, C)
    for rep in range(2):
        out = ops.attention_prescaled(pre, heads).float()
        rel = float((out - want).norm() / want.norm())
        if not (rel < 6e-3) or not torch.isfinite(out).all():
            bad += 1; print("BAD", trial, hd, heads, B, L, rel)
print("stress done, bad =", bad)
