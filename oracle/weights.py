"""TEST INFRASTRUCTURE ONLY -- keyed, construction-order-independent random weights.

Every tensor of a checkpoint is drawn from its own generator seeded by (seed, key name), so the oracle, the
unmodified reference and the product model all receive bit-identical parameters through `load_state_dict`,
whatever order their modules were built in.  Statistics follow PyTorch's default initialisers (uniform
+-1/sqrt(fan_in) for conv/linear, randn for the AVIF transforms, avif_inference.py:129-130) with the
normalisation affine parameters and zero-initialised biases perturbed so they are exercised too.
"""
import hashlib
import math
from typing import Dict

import torch

from .restated import BLOCKS, FREQ_BLOCKS_0409, dct_matrix


def _gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator()
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def shapes_0409() -> Dict[str, tuple]:
    """Checkpoint layout of the 0409 notebook's JPEGDiffusionModel (0409_method.ipynb#c0:L371-400)."""
    s: Dict[str, tuple] = {}
    s["time_embed.proj.0.weight"] = (1024, 256); s["time_embed.proj.0.bias"] = (1024,)
    s["time_embed.proj.2.weight"] = (256, 1024); s["time_embed.proj.2.bias"] = (256,)
    for p, ci, co in BLOCKS:
        s[f"{p}.norm1.weight"] = (ci,); s[f"{p}.norm1.bias"] = (ci,)
        s[f"{p}.conv1.weight"] = (co, ci, 3, 3); s[f"{p}.conv1.bias"] = (co,)
        s[f"{p}.time_proj.weight"] = (co, 256); s[f"{p}.time_proj.bias"] = (co,)
        s[f"{p}.norm2.weight"] = (co,); s[f"{p}.norm2.bias"] = (co,)
        s[f"{p}.conv2.weight"] = (co, co, 3, 3); s[f"{p}.conv2.bias"] = (co,)
        s[f"{p}.attn.in_proj_weight"] = (3 * co, co); s[f"{p}.attn.in_proj_bias"] = (3 * co,)
        s[f"{p}.attn.out_proj.weight"] = (co, co); s[f"{p}.attn.out_proj.bias"] = (co,)
        if ci != co:
            s[f"{p}.shortcut.weight"] = (co, ci, 1, 1); s[f"{p}.shortcut.bias"] = (co,)
        if p in FREQ_BLOCKS_0409:
            f, h = f"{p}.freq_guide", f"{p}.hfcm"
            s[f"{f}.freq_conv.weight"] = (co, co, 3, 3); s[f"{f}.freq_conv.bias"] = (co,)
            s[f"{f}.freq_attn.1.weight"] = (co // 4, co, 1, 1); s[f"{f}.freq_attn.1.bias"] = (co // 4,)
            s[f"{f}.freq_attn.3.weight"] = (co, co // 4, 1, 1); s[f"{f}.freq_attn.3.bias"] = (co,)
            s[f"{h}.high_freq_attn.0.weight"] = (co, co, 3, 3); s[f"{h}.high_freq_attn.0.bias"] = (co,)
            s[f"{h}.high_freq_attn.2.weight"] = (co, co, 3, 3); s[f"{h}.high_freq_attn.2.bias"] = (co,)
            s[f"{h}.conv_out.weight"] = (co, co, 1, 1); s[f"{h}.conv_out.bias"] = (co,)
    s["out_conv.weight"] = (3, 64, 1, 1); s["out_conv.bias"] = (3,)
    return s


def shapes(family: str) -> Dict[str, tuple]:
    """Checkpoint layout of {WebP,JPEG,AVIF}DiffusionModel (356 / 356 / 634 entries); "m0409" = the 0409 notebook's model."""
    if family == "m0409":
        return shapes_0409()
    s: Dict[str, tuple] = {}
    s["time_embed.proj.0.weight"] = (1024, 256); s["time_embed.proj.0.bias"] = (1024,)
    s["time_embed.proj.2.weight"] = (256, 1024); s["time_embed.proj.2.bias"] = (256,)
    for p, ci, co in BLOCKS:
        s[f"{p}.norm1.weight"] = (ci,); s[f"{p}.norm1.bias"] = (ci,)
        s[f"{p}.conv1.weight"] = (co, ci, 3, 3); s[f"{p}.conv1.bias"] = (co,)
        s[f"{p}.time_proj.weight"] = (co, 256); s[f"{p}.time_proj.bias"] = (co,)
        s[f"{p}.norm2.weight"] = (co,); s[f"{p}.norm2.bias"] = (co,)
        s[f"{p}.conv2.weight"] = (co, co, 3, 3); s[f"{p}.conv2.bias"] = (co,)
        s[f"{p}.attn.in_proj_weight"] = (3 * co, co); s[f"{p}.attn.in_proj_bias"] = (3 * co,)
        s[f"{p}.attn.out_proj.weight"] = (co, co); s[f"{p}.attn.out_proj.bias"] = (co,)
        f = f"{p}.freq_guide"
        if family in ("webp", "jpeg"):
            bs = 4 if family == "webp" else 8
            s[f"{f}.dct.dct_matrix"] = (bs, bs)
            for g in ("low_freq_attn", "high_freq_attn"):
                s[f"{f}.{g}.0.weight"] = (co // 2, co, 1, 1); s[f"{f}.{g}.0.bias"] = (co // 2,)
                s[f"{f}.{g}.2.weight"] = (co, co // 2, 1, 1); s[f"{f}.{g}.2.bias"] = (co,)
        else:
            a = f"{f}.adaptive_transform"
            s[f"{a}.transform_weights"] = (co, 8, 8); s[f"{a}.inverse_weights"] = (co, 8, 8)
            s[f"{a}.quantization.0.weight"] = (co, co, 1, 1); s[f"{a}.quantization.0.bias"] = (co,)
            s[f"{a}.quantization.2.weight"] = (co, co, 1, 1); s[f"{a}.quantization.2.bias"] = (co,)
            for i in range(4):
                s[f"{f}.multi_scale_attn.{i}.1.weight"] = (co // 4, co, 1, 1); s[f"{f}.multi_scale_attn.{i}.1.bias"] = (co // 4,)
                s[f"{f}.multi_scale_attn.{i}.3.weight"] = (co, co // 4, 1, 1); s[f"{f}.multi_scale_attn.{i}.3.bias"] = (co,)
            s[f"{f}.color_consistency.0.weight"] = (co, co, 1, 1); s[f"{f}.color_consistency.0.bias"] = (co,)
            s[f"{f}.color_consistency.2.weight"] = (co, co, 1, 1); s[f"{f}.color_consistency.2.bias"] = (co,)
            s[f"{f}.edge_preserve.0.weight"] = (co // 2, co, 3, 3); s[f"{f}.edge_preserve.0.bias"] = (co // 2,)
            s[f"{f}.edge_preserve.2.weight"] = (co, co // 2, 3, 3); s[f"{f}.edge_preserve.2.bias"] = (co,)
        s[f"{f}.conv_out.weight"] = (co, co, 3, 3); s[f"{f}.conv_out.bias"] = (co,)
        if ci != co:
            s[f"{p}.shortcut.weight"] = (co, ci, 1, 1); s[f"{p}.shortcut.bias"] = (co,)
    if family == "avif":
        a = "avif_layer"
        s[f"{a}.transform_weights"] = (64, 8, 8); s[f"{a}.inverse_weights"] = (64, 8, 8)
        s[f"{a}.quantization.0.weight"] = (64, 64, 1, 1); s[f"{a}.quantization.0.bias"] = (64,)
        s[f"{a}.quantization.2.weight"] = (64, 64, 1, 1); s[f"{a}.quantization.2.bias"] = (64,)
    else:
        bs = 4 if family == "webp" else 8
        s["dct_layer.dct_matrix"] = (bs, bs)
    s["out_conv.0.weight"] = (64,); s["out_conv.0.bias"] = (64,)
    s["out_conv.2.weight"] = (3, 64, 3, 3); s["out_conv.2.bias"] = (3,)
    return s


def make_state_dict(family: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    sd = {}
    for key, shp in shapes(family).items():
        g = _gen(seed, key)
        if key.endswith("dct_matrix"):
            sd[key] = dct_matrix(shp[0])
        elif "norm" in key or key.startswith("out_conv.0"):
            if key.endswith("weight"):
                sd[key] = 1.0 + 0.1 * torch.randn(shp, generator=g)
            else:
                sd[key] = 0.1 * torch.randn(shp, generator=g)
        elif key.endswith("transform_weights") or key.endswith("inverse_weights"):
            sd[key] = torch.randn(shp, generator=g)
        elif key.endswith("bias"):
            fan_in = {"in_proj_bias": shp[0] // 3}.get(key.split(".")[-1], None)
            # bias bound follows the matching weight's fan_in; look it up from the sibling weight
            wkey = key[:-4] + "weight" if not key.endswith("in_proj_bias") else key.replace("in_proj_bias", "in_proj_weight")
            wshape = shapes_cache(family)[wkey]
            fan = math.prod(wshape[1:])
            sd[key] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan)
        else:
            fan = math.prod(shp[1:])
            sd[key] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan)
    return sd


_SHAPES = {}


def shapes_cache(family: str):
    if family not in _SHAPES:
        _SHAPES[family] = shapes(family)
    return _SHAPES[family]


def synthetic_images(b: int, h: int, w: int, seed: int = 1234) -> torch.Tensor:
    """smooth + texture uint8 RGB images, returned as [-1,1] fp32 NCHW (pre-codec)."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    imgs = []
    for n in range(b):
        chans = []
        for c in range(3):
            ph = float(torch.rand(1, generator=g)) * 6.28
            base = 127 + 100 * torch.sin(xx / 17 + c + ph) * torch.cos(yy / 23 + 0.5 * n)
            chans.append(base + 8 * torch.randn(h, w, generator=g))
        imgs.append(torch.stack(chans))
    u8 = torch.stack(imgs).clamp(0, 255).to(torch.uint8)
    return u8.float() / 255.0 * 2 - 1
