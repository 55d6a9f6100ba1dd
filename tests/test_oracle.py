"""CPU tests (no GPU): the restated oracle against the fixtures minted from the UNMODIFIED reference
(tests/golden/*.npz, oracle/make_golden.py) and known-answer vectors."""
import numpy as np
import pytest
import torch

from oracle import restated as R
from oracle import weights as W
from util import rel

NOISE_SEED = 7


def philox_noise(i, like):
    return torch.from_numpy(R.philox_normal(NOISE_SEED, i, like.numel()).astype(np.float32)).view_as(like)


def coin(i):
    w = R.philox4x32_10(np.array([i, 0, 0, 0x636F696E], dtype=np.uint32), np.array([NOISE_SEED, 0], dtype=np.uint32))
    return float(w[0]) * 2.0 ** -32


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    P = lambda c, k: [int(v) for v in R.philox4x32_10(np.array(c, dtype=np.uint32), np.array(k, dtype=np.uint32))]
    assert P([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert P([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert P([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    z = R.philox_normal(3, 1, 1 << 18)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


@pytest.mark.parametrize("fam", ["webp", "jpeg", "avif"])
def test_unet_oracle_vs_reference_fixture(golden, fam):
    d = golden(f"unet_{fam}_32.npz")
    sd = W.make_state_dict(fam, 0)
    x, t, lvl = (torch.from_numpy(d[k]) for k in ("x", "t", "level"))
    assert rel(R.unet_forward(sd, x, t, lvl, fam), torch.from_numpy(d["out"])) < 5e-6
    taps = {}
    assert rel(R.unet_forward(sd, x, t, None, fam, taps), torch.from_numpy(d["out_nolevel"])) < 5e-6
    for name, key in (("d1", "tap_down1"), ("d3", "tap_down3"), ("u5", "tap_up5")):
        assert rel(taps[name][:, ::4, ::2, ::2], torch.from_numpy(d[key])) < 5e-6


def test_unet_oracle_64_fixture(golden):
    d = golden("unet_webp_64.npz")
    sd = W.make_state_dict("webp", 0)
    x, t, lvl = (torch.from_numpy(d[k]) for k in ("x", "t", "level"))
    assert rel(R.unet_forward(sd, x, t, lvl, "webp"), torch.from_numpy(d["out"])) < 5e-6


def test_low_mask_equals_block_loop():
    """The static mask against a literal restatement of the reference's loop semantics, ragged shapes included."""
    for bs, low in ((4, 3), (8, 4)):
        for h, w in ((16, 16), (1, 1), (2, 2), (5, 7), (9, 4), (8, 24)):
            m = torch.zeros(h, w, dtype=torch.bool)
            for i in range(0, h, bs):
                ie = min(i + bs, h)
                for j in range(0, w, bs):
                    je = min(j + bs, w)
                    ls = max(1, min(low, min(ie - i, je - j)))
                    m[i:i + ls, j:j + ls] = True
            assert torch.equal(m, R.low_mask(h, w, bs, low)), (bs, h, w)


def test_op_fixtures(golden):
    d = golden("ops.npz")
    x, xn = torch.from_numpy(d["x"]), torch.from_numpy(d["xn"])
    assert torch.equal(R.codec_roundtrip(x, 10, "webp"), torch.from_numpy(d["webp_q10"]))
    assert torch.equal(R.codec_roundtrip(x, 20, "avif"), torch.from_numpy(d["avif_q20"]))
    assert torch.equal(R.codec_roundtrip(x, 10, "jpeg"), torch.from_numpy(d["jpeg_q10"]))
    assert torch.equal(R.codec_roundtrip(x, 50, "jpeg"), torch.from_numpy(d["jpeg_q50"]))
    assert rel(R.phase_consistency(xn, torch.from_numpy(d["webp_q10"]), 0.7), torch.from_numpy(d["phase_a07"])) < 1e-6
    assert rel(R.svd_structure_preservation(xn, 0.6), torch.from_numpy(d["svd_k06"])) < 1e-5
    assert rel(R.svd_structure_preservation(xn, 0.01), torch.from_numpy(d["svd_k01"])) < 1e-5
    assert abs(float(R.color_loss_conv_deep(xn * 1.5, x)) - float(d["color_deep"])) < 1e-7
    z, t = torch.from_numpy(d["z"]), torch.from_numpy(d["t"])
    assert torch.equal(R.ddrm_update(xn, torch.from_numpy(d["webp_q10"]), x, z, t, 0.2, 0.85, 1.0), torch.from_numpy(d["update_eta1"]))
    assert torch.equal(R.ddrm_update(xn, torch.from_numpy(d["webp_q10"]), x, z, t, 0.2, 0.85, 0.6), torch.from_numpy(d["update_eta06"]))


@pytest.mark.parametrize("fam", ["webp", "jpeg", "avif"])
def test_ddrm_sampler_oracle_vs_reference_fixture(golden, fam):
    d = golden(f"ddrm_{fam}_32.npz")
    sd = W.make_state_dict(fam, 0)
    y = torch.from_numpy(d["y"])
    out = R.ddrm_sample(lambda x, t, l: R.unet_forward(sd, x, t, l, fam), y, int(d["quality"]), int(d["steps"]), fam,
                        noise_fn=philox_noise)
    # the uint8 truncation inside the codec hop can flip a level on 1e-7 differences -> PSNR-level comparison
    assert abs(R.psnr(out, torch.from_numpy(d["clean"])) - float(d["psnr_out"])) < 0.02
    assert R.psnr(out, torch.from_numpy(d["out"])) > 45.0


def test_gmm_sampler_oracle_vs_reference_fixture(golden):
    d = golden("gmm_jpeg_32.npz")
    sd = W.make_state_dict("jpeg", 0)
    out = R.gmm_sample(lambda x, t, l: R.unet_forward(sd, x, t, l, "jpeg"), torch.from_numpy(d["y"]), int(d["steps"]),
                       noise_fn=philox_noise, coin_fn=coin)
    assert rel(out, torch.from_numpy(d["out"])) < 2e-5


def test_ssim_and_losses_sanity():
    g = torch.Generator().manual_seed(0)
    a = torch.rand(2, 3, 32, 32, generator=g)
    assert abs(float(R.ssim(a, a)) - 1.0) < 1e-6
    b = (a + 0.1 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    assert 0.0 < float(R.ssim(a, b)) < 1.0
    assert float(R.color_preservation_loss(a * 2 - 1, a * 2 - 1)) < 1e-6
    betas, alphas, abar = R.ddpm_schedule(100)
    assert abs(float(betas[0]) - 1e-4) < 1e-9 and abs(float(betas[-1]) - 0.02) < 1e-9


def test_trajectory256_fixture_is_consistent(golden):
    """BASELINE config 1 golden (one 256x256 WebP q=10 image, 80 steps; minted with the restated oracle)."""
    d = golden("traj256_webp.npz")
    assert d["out"].shape == (1, 3, 256, 256) and int(d["steps"]) == 80 and int(d["quality"]) == 10
    clean = torch.from_numpy(d["clean_u8"]).float() / 255 * 2 - 1
    assert abs(R.psnr(torch.from_numpy(d["out"]).float(), clean) - float(d["psnr_out"])) < 0.01


def test_dct_jpeg_projection_vs_reference_fixture(golden):
    """DCTProcessor.jpeg_compress (dct.ipynb#c2:L100-139), run with the reference's own scalar loops -> dct_jpeg.npz."""
    d = golden("dct_jpeg.npz")
    x = torch.from_numpy(d["x"])
    for q in (10, 50, 90):
        out = R.dct_jpeg_project(x, q)
        assert (out - torch.from_numpy(d[f"q{q}"])).abs().max() < 1e-3      # 0..255 scale
    # quality scaling of the tables (L105-112)
    qy, qc = R.jpeg_quant_tables(10)
    assert qy[0, 0] == 80 and qc[7, 7] == 495 and R.jpeg_quant_tables(100)[0].max() == 1
    # projection: applying it twice changes nothing beyond rounding
    once = R.dct_jpeg_project(x, 50)
    assert (R.dct_jpeg_project(once, 50) - once).abs().max() < 1e-3


def test_m0409_unet_and_gmm_oracle_vs_reference_fixture(golden):
    """The 0409 notebook's own UNet (HFCM, FrequencyAwareBlock; 0409_method.ipynb#c0:L184-428) and its GMM sampler on it."""
    sd = W.make_state_dict("m0409", 0)
    d = golden("unet_m0409_32.npz")
    x, t, lvl = (torch.from_numpy(d[k]) for k in ("x", "t", "level"))
    assert rel(R.unet0409_forward(sd, x, t, lvl), torch.from_numpy(d["out"])) < 2e-5
    assert rel(R.unet0409_forward(sd, x, t), torch.from_numpy(d["out_nolevel"])) < 2e-5
    d = golden("gmm_m0409_32.npz")
    out = R.gmm_sample(lambda x, t, l: R.unet0409_forward(sd, x, t, l), torch.from_numpy(d["y"]), int(d["steps"]),
                       noise_fn=philox_noise, coin_fn=coin)
    assert rel(out, torch.from_numpy(d["out"])) < 2e-5


def test_jpeg_exact_restatement_vs_pillow():
    """oracle/jpeg_exact.py (libjpeg-turbo's integer pipeline restated) against Pillow's own JPEG round trip: bit-exact."""
    from oracle import jpeg_exact as J
    rng = np.random.default_rng(5)
    for trial in range(18):
        if trial < 12:
            H, W = 16 * int(rng.integers(1, 4)), 16 * int(rng.integers(1, 5))
        else:       # sizes that need MCU edge padding, down to widths the decoder upsamples by replication
            H, W = int(rng.integers(1, 50)), int(rng.integers(1, 60)) if trial % 2 else int(rng.integers(1, 6))
        if trial % 3 == 0:
            img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        elif trial % 3 == 1:
            yy, xx = np.mgrid[0:H, 0:W]
            base = 127 + 110 * np.sin(xx / rng.uniform(3, 20)) * np.cos(yy / rng.uniform(3, 20))
            img = (np.stack([base, base[::-1], 255 - base], -1) + rng.normal(0, 15, (H, W, 3))).clip(0, 255).astype(np.uint8)
        else:
            img = np.zeros((H, W, 3), np.uint8); img[:, W // 2:] = 255; img[H // 2:, :, 1] = 77
        for q in (1, 10, 30, 31, 50, 75, 95, 100):
            sub = 0 if q > 30 else 2          # jpeg_compress: 4:4:4 above quality 30, 4:2:0 otherwise
            assert np.array_equal(J.pil_roundtrip(img, q, "RGB", sub), J.roundtrip_rgb(img, q, sub == 2)), (trial, q)
