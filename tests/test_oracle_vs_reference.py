"""Pins the restated oracle directly against the unmodified reference sources (only where /root/reference exists,
i.e. in the build container; the GPU box relies on the committed fixtures these same checks produced)."""
import pytest
import torch

from oracle import reference_loader as rl
from oracle import restated as R
from oracle import weights as W
from util import rel

pytestmark = pytest.mark.skipif(not rl.available(), reason="reference tree not mounted")


def test_webp_unet_and_codec_against_reference():
    ns = rl.load_webp()
    m = ns["WebPDiffusionModel"]().eval()
    sd = W.make_state_dict("webp", 0)
    assert list(m.state_dict().keys()) == list(W.shapes("webp").keys())
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 3, 32, 32, generator=g) * 0.5
    t = torch.tensor([0.45])
    with torch.no_grad():
        ref = m(x, t)
    assert rel(R.unet_forward(sd, x, t, None, "webp"), ref) < 5e-6
    img = W.synthetic_images(1, 32, 32)
    assert torch.equal(ns["webp_compress"](img, 10), R.codec_roundtrip(img, 10, "webp"))
    assert rel(R.phase_consistency(x, img, 0.7), ns["phase_consistency"](x, img, 0.7)) < 1e-6


def test_avif_checkpoint_layout_against_reference():
    ns = rl.load_avif()
    m = ns["AVIFDiffusionModel"]()
    ref = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref == W.shapes("avif")
