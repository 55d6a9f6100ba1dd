"""Relative L2 of the bf16 UNet forward against the restated CPU oracle at 128^2, per attention exp split
(tuning aid):  python tools/unet_parity.py [family] [sel ...]   (sel as in ddpmir_attention_set_expmode >> 8, minus 1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ddpm_image_restoration_b200 as P
from ddpm_image_restoration_b200 import _lib
from oracle import restated as R, weights as Wt

fam = sys.argv[1] if len(sys.argv) > 1 else "webp"
sels = [int(a) for a in sys.argv[2:]] or [-1]
cls = {"webp": P.WebPDiffusionModel, "jpeg": P.JPEGDiffusionModel, "avif": P.AVIFDiffusionModel}[fam]
sd = Wt.make_state_dict(fam, 0)
m = cls()
m.load_state_dict(sd)
m = m.cuda().eval()
g = torch.Generator().manual_seed(5)
x = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
t = torch.tensor([0.7, 0.2])
ref = R.unet_forward(sd, x, t, None, fam)
rel = lambda a, b: float((a - b).norm() / b.norm())
for sel in sels:
    _lib.lib().ddpmir_attention_set_expmode(-1 if sel < 0 else ((sel + 1) << 8))
    out = m(x.cuda(), t.cuda()).cpu()
    print(f"{fam} 128^2 bf16 sel={sel}: rel-L2 {rel(out, ref):.3e}")
