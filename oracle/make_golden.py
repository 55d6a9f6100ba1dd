"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.npz by running the UNMODIFIED reference (loaded from
/root/reference by oracle/reference_loader.py) on seeded inputs with keyed weights.  Run in the build
container only:   python -m oracle.make_golden [--trajectory256]

Fixtures (all small):
  unet_{webp,jpeg,avif}_32.npz   reference UNet forward, B=2, 32x32 (ragged DCT blocks at the 1x1..4x4 levels)
  unet_{webp,avif}_64.npz        reference UNet forward, B=1, 64x64 (the resolution the reference ships with)
  ddrm_{webp,jpeg,avif}_32.npz   reference DDRM sampler, 6 steps, noise injected = oracle.restated.philox_normal
  gmm_jpeg_32.npz                reference GaussianMixtureSampler (0409), 8 steps, with SVD guide + phase consistency
  ops.npz                        codec round trips, phase_consistency, svd_structure_preservation, colour losses
  unet_m0409_{32,64}.npz, gmm_m0409_32.npz   the 0409 notebook's own UNet and its GMM sampler on that model
  dct_jpeg.npz                   DCTProcessor.jpeg_compress (dct.ipynb#c2) at q 10/50/90 on two 16x24 images
  traj256_webp.npz               (--trajectory256) BASELINE config 1: one 256x256 WebP q=10 image, 80 steps, computed
                                 with the restated oracle (the verbatim reference cannot allocate 68.7 GB at 256x256)
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

from . import reference_loader as rl
from . import restated as R
from . import weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
NOISE_SEED = 7


class _InjectNoise:
    """Replace torch.randn_like / torch.rand inside the reference samplers by keyed, order-free noise."""

    def __init__(self, steps_iter):
        self.steps = list(steps_iter)
        self.k = 0

    def __enter__(self):
        self._randn_like, self._rand = torch.randn_like, torch.rand
        def randn_like(x, **kw):
            i = self.steps[self.k]; self.k += 1
            return philox_noise(i, x)
        def rand(*a, **kw):
            i = self.steps[self.k]   # coin drawn before that step's randn_like
            return torch.tensor([coin(i)])
        torch.randn_like, torch.rand = randn_like, rand
        return self

    def __exit__(self, *a):
        torch.randn_like, torch.rand = self._randn_like, self._rand


def philox_noise(i, like):
    z = R.philox_normal(NOISE_SEED, i, like.numel()).astype(np.float32)
    return torch.from_numpy(z).view_as(like)


def coin(i):
    # keyed uniform in [0,1): word 0 of Philox(counter=(i,0,0,0x636f696e), key=seed)
    w = R.philox4x32_10(np.array([i, 0, 0, 0x636F696E], dtype=np.uint32), np.array([NOISE_SEED, 0], dtype=np.uint32))
    return float(w[0]) * 2.0 ** -32


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    arrs = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    np.savez_compressed(os.path.join(OUT, name), **arrs)
    print("wrote", name, {k: v.shape for k, v in arrs.items()})


def unet_goldens():
    loaders = {"webp": (rl.load_webp, "WebPDiffusionModel"), "jpeg": (rl.load_jpeg, "JPEGDiffusionModel"),
               "avif": (rl.load_avif, "AVIFDiffusionModel")}
    for fam, (loader, cls) in loaders.items():
        ns = loader()
        m = ns[cls]().eval()
        m.load_state_dict(W.make_state_dict(fam, 0))
        for hw, b in ((32, 2), (64, 1)):
            if hw == 64 and fam == "jpeg":
                continue
            g = torch.Generator().manual_seed(100 + hw)
            x = torch.randn(b, 3, hw, hw, generator=g) * 0.5
            t = torch.tensor([0.37, 0.8][:b])
            lvl = torch.tensor([0.2, 0.9][:b])
            taps = {}
            hooks = []
            for name in ("down1", "down3", "bottleneck", "up5"):
                mod = dict(m.named_modules())[name]
                hooks.append(mod.register_forward_hook(lambda _m, _i, o, n=name: taps.__setitem__(n, o.detach())))
            with torch.no_grad():
                out = m(x, t, lvl)
                out_nolevel = m(x, t)
            for h in hooks:
                h.remove()
            save(f"unet_{fam}_{hw}.npz", x=x, t=t, level=lvl, out=out, out_nolevel=out_nolevel,
                 **{f"tap_{k.replace('.', '_')}": v[:, ::4, ::2, ::2] for k, v in taps.items()})


def sampler_goldens():
    specs = [("webp", rl.load_webp, "WebPDiffusionModel", "DDRMWebPSampler", 10),
             ("jpeg", rl.load_jpeg, "JPEGDiffusionModel", "DDRMJPEGSampler", 10),
             ("avif", rl.load_avif, "AVIFDiffusionModel", "DDRMAVIFSampler", 20)]
    clean = W.synthetic_images(2, 32, 32)
    for fam, loader, cls, scls, q in specs:
        ns = loader()
        m = ns[cls]().eval()
        m.load_state_dict(W.make_state_dict(fam, 0))
        y = R.codec_roundtrip(clean, q, R.DDRM[fam]["codec"])
        steps = 6
        with _InjectNoise(range(steps - 1, 0, -1)):
            out = ns[scls](m).sample(y.clone(), q, steps=steps)
        save(f"ddrm_{fam}_32.npz", clean=clean, y=y, quality=q, steps=steps, out=out,
             psnr_in=R.psnr(y, clean), psnr_out=R.psnr(out, clean))
    # GMM + SVD guide (0409) driven by the JPEG UNet
    ns9 = rl.load_0409()
    nsj = rl.load_jpeg()
    m = nsj["JPEGDiffusionModel"]().eval()
    m.load_state_dict(W.make_state_dict("jpeg", 0))
    y = R.codec_roundtrip(clean, 10, "jpeg")
    steps = 8
    with _InjectNoise(range(steps - 1, 0, -1)):
        out = ns9["GaussianMixtureSampler"](m, num_timesteps=100).sample(y.clone(), steps=steps)
    save("gmm_jpeg_32.npz", clean=clean, y=y, steps=steps, out=out, psnr_out=R.psnr(out, clean))


def op_goldens():
    w, a, j, n9, cd = rl.load_webp(), rl.load_avif(), rl.load_jpeg(), rl.load_0409(), rl.load_conv_deep_color_loss()
    x = W.synthetic_images(2, 64, 64, seed=4321)
    d = dict(x=x)
    d["webp_q10"] = w["webp_compress"](x, 10)
    d["avif_q20"] = a["avif_compress"](x, 20)
    d["jpeg_q10"] = j["jpeg_compress"](x, 10)
    d["jpeg_q50"] = j["jpeg_compress"](x, 50)
    g = torch.Generator().manual_seed(5)
    xn = x + 0.1 * torch.randn(x.shape, generator=g)
    d["xn"] = xn
    d["phase_a07"] = w["phase_consistency"](xn, d["webp_q10"], 0.7)
    d["phase_a08"] = a["phase_consistency"](xn, d["avif_q20"], 0.8)
    d["svd_k06"] = n9["svd_structure_preservation"](xn, 0.6)
    d["svd_k01"] = n9["svd_structure_preservation"](xn, 0.01)
    d["color_deep"] = cd["color_loss"](xn * 1.5, x)
    # update rule, evaluated with the reference's own expression order (webp_inference.py:584-592)
    z = torch.randn(x.shape, generator=g)
    t = torch.full((2,), 37).float() / 80
    d["z"] = z
    d["t"] = t
    x_prime = xn - d["webp_q10"] + x
    d["update_eta1"] = 1.0 * x_prime + (1 - 1.0) * xn + 0.85 * (z * (t * 0.2).view(-1, 1, 1, 1))
    d["update_eta06"] = 0.6 * x_prime + (1 - 0.6) * xn + 0.85 * (z * (t * 0.2).view(-1, 1, 1, 1))
    save("ops.npz", **d)


def m0409_goldens():
    """unet_m0409_{32,64}.npz + gmm_m0409_32.npz: the 0409 notebook's own UNet (HFCM / FrequencyAwareBlock) and its
    GaussianMixtureSampler running on that native model."""
    nsm = rl.load_0409_model()
    nsm["device"] = torch.device("cpu")
    m = nsm["JPEGDiffusionModel"]().eval()
    m.load_state_dict(W.make_state_dict("m0409", 0))
    for hw, b in ((32, 2), (64, 1)):
        g = torch.Generator().manual_seed(300 + hw)
        x = torch.randn(b, 3, hw, hw, generator=g) * 0.5
        t = torch.tensor([0.37, 0.8][:b])
        lvl = torch.tensor([0.2, 0.9][:b])
        with torch.no_grad():
            save(f"unet_m0409_{hw}.npz", x=x, t=t, level=lvl, out=m(x, t, lvl), out_nolevel=m(x, t))
    ns9 = rl.load_0409()
    clean = W.synthetic_images(2, 32, 32)
    y = R.codec_roundtrip(clean, 10, "jpeg")
    steps = 8
    with _InjectNoise(range(steps - 1, 0, -1)):
        out = ns9["GaussianMixtureSampler"](m, num_timesteps=100).sample(y.clone(), steps=steps)
    save("gmm_m0409_32.npz", clean=clean, y=y, steps=steps, out=out, psnr_out=R.psnr(out, clean))


def dct_goldens():
    """dct_jpeg.npz: DCTProcessor.jpeg_compress (dct.ipynb#c2) on 0..255 images, run with the reference's own scalar loops."""
    P = rl.load_dct_processor()["DCTProcessor"](torch.device("cpu"))
    x = (W.synthetic_images(2, 16, 24, seed=99) * 127.5 + 127.5).contiguous()
    save("dct_jpeg.npz", x=x, q10=P.jpeg_compress(x, quality=10), q50=P.jpeg_compress(x, quality=50), q90=P.jpeg_compress(x, quality=90))


def unet256():
    """unet_{webp,avif,jpeg}_256.npz: one UNet forward at the resolution the bench is quoted on (B=1, L = 65 536 tokens
    at full resolution), restated oracle (the verbatim reference needs 68.7 GB for one score tensor at this size)."""
    for fam in ("avif", "webp", "jpeg"):
        sd = W.make_state_dict(fam, 0)
        g = torch.Generator().manual_seed(256 + len(fam))
        clean = W.synthetic_images(1, 256, 256, seed=77)
        x = clean + 0.1 * torch.randn(clean.shape, generator=g)      # image-like x_t, as the sampler feeds it
        t = torch.tensor([0.44])
        t0 = time.time()
        with torch.no_grad():
            out = R.unet_forward(sd, x, t, None, fam)
        save(f"unet_{fam}_256.npz", x=x, t=t, out=out, cpu_seconds=time.time() - t0)


def gmm256():
    """gmm_jpeg_256.npz: GaussianMixtureSampler (0409) at 256x256, 6 steps (i = 5, 4 take the SVD guide at rank 213 / 170,
    i = 5 the phase consistency), JPEG UNet, restated oracle."""
    sd = W.make_state_dict("jpeg", 0)
    clean = W.synthetic_images(1, 256, 256, seed=1234)
    y = R.codec_roundtrip(clean, 10, "jpeg")
    steps = 6
    t0 = time.time()
    trace = []
    with torch.no_grad():
        out = R.gmm_sample(lambda x, t, l: R.unet_forward(sd, x, t, l, "jpeg"), y, steps, philox_noise, coin, trace=trace)
    save("gmm_jpeg_256.npz", clean_u8=R.quantize_u8(clean), y=y, steps=steps, out=out, x_after_first=trace[0],
         psnr_out=R.psnr(out, clean), cpu_seconds=time.time() - t0)


def trajectory256(n_images=1, fam="webp"):
    q, steps = {"webp": (10, 80), "avif": (20, 75), "jpeg": (10, 80)}[fam]   # BASELINE configs[0], [1], [2]'s codec/quality
    sd = W.make_state_dict(fam, 0)
    clean = W.synthetic_images(n_images, 256, 256, seed=1234)
    y = R.codec_roundtrip(clean, q, fam)
    t0 = time.time()
    trace = []
    def model_fn(x, t, lvl):
        o = R.unet_forward(sd, x, t, lvl, fam)
        print(f"  step done {time.time() - t0:.0f}s", flush=True)
        return o
    out = R.ddrm_sample(model_fn, y, q, steps, fam, noise_fn=philox_noise, trace=trace)
    name = f"traj256_{fam}.npz" if n_images == 1 else f"traj256x{n_images}_{fam}.npz"
    save(name, clean_u8=R.quantize_u8(clean), y_u8=R.quantize_u8(y), quality=q, steps=steps,
         out=out.half(), psnr_in=R.psnr(y, clean), psnr_out=R.psnr(out, clean),
         psnr_each=np.array([R.psnr(out[i:i + 1], clean[i:i + 1]) for i in range(n_images)]),
         psnr_trace=np.array([R.psnr(z, clean) for z in trace]), cpu_seconds=time.time() - t0,
         cpu_threads=torch.get_num_threads())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectory256", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--images", type=int, default=1)
    ap.add_argument("--family", default="webp")
    ap.add_argument("--unet256", action="store_true")
    ap.add_argument("--gmm256", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    if args.threads:
        torch.set_num_threads(args.threads)
    if not rl.available():
        sys.exit("reference tree not mounted; goldens can only be minted in the build container")
    if args.trajectory256:
        trajectory256(args.images, args.family)
    elif args.unet256:
        unet256()
    elif args.gmm256:
        gmm256()
    else:
        for fn in (op_goldens, unet_goldens, sampler_goldens, dct_goldens, m0409_goldens):
            if not args.only or args.only in fn.__name__:
                fn()
