"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference sources from /root/reference.

This file exists only in the build container (the GPU box has no /root/reference).  It is used by
`oracle/make_golden.py` to mint the committed fixtures under tests/golden/ and by the CPU tests that
pin `oracle/restated.py` against the reference itself.  Nothing in the product package imports it.

Recipe: the reference scripts cannot be imported (dataset code runs at import,
third-party metric packages are missing), so we insert stub modules and `exec` the line ranges that hold
the codec wrappers, the UNet classes and the samplers, straight from the read-only tree.  No reference
source is copied into this repository.
"""
import json
import os
import sys
import types

REF_ROOT = os.environ.get("DDPMIR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "webp_inference.py"))


def _install_stubs():
    def _stub(name, **attrs):
        if name in sys.modules:
            return
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m

    def _missing(*a, **k):
        raise RuntimeError("third-party metric stubbed out in the oracle loader")

    _stub("pytorch_msssim", ssim=_missing)
    _stub("lpips", LPIPS=_missing)
    _stub("matplotlib", pyplot=types.ModuleType("matplotlib.pyplot"))
    _stub("matplotlib.pyplot")
    _stub("pillow_avif")
    _stub("pytorch_fid", fid_score=None)
    try:
        import tqdm  # noqa: F401
    except Exception:  # pragma: no cover
        _stub("tqdm", tqdm=lambda it, **k: it)


def _slice(lines, ranges):
    out = []
    for a, b in ranges:
        out.extend(lines[a - 1:b])
    return "\n".join(out)


def _exec(src, name, prelude=""):
    ns = {"__name__": name}
    code = prelude + src
    exec(compile(code, name, "exec"), ns)
    return ns


def _quiet_tqdm(ns):
    # the notebook samplers wrap their loop in tqdm(); keep logs quiet
    if "tqdm" in ns:
        ns["tqdm"] = lambda it, **k: it


def load_webp():
    """webp_inference.py 1-19 + 79-473 (+506-602 are duplicates of the same functions)."""
    _install_stubs()
    lines = open(os.path.join(REF_ROOT, "webp_inference.py"), encoding="utf-8").read().split("\n")
    ns = _exec(_slice(lines, [(1, 19), (79, 473)]), "ref_webp")
    _quiet_tqdm(ns)
    return ns


def load_avif():
    """avif_inference.py 1-21 + 63-459."""
    _install_stubs()
    lines = open(os.path.join(REF_ROOT, "avif_inference.py"), encoding="utf-8").read().split("\n")
    ns = _exec(_slice(lines, [(1, 21), (63, 459)]), "ref_avif")
    _quiet_tqdm(ns)
    return ns


def _nb_cell(nb, idx):
    cells = json.load(open(os.path.join(REF_ROOT, nb), encoding="utf-8"))["cells"]
    return "".join(cells[idx]["source"]).split("\n")


def load_jpeg():
    """svd.ipynb cell 1, lines 1-17 + 20-385 (cell 1 relies on cell 0's `from torch import nn`)."""
    _install_stubs()
    lines = _nb_cell("svd.ipynb", 1)
    ns = _exec(_slice(lines, [(1, 17), (20, 385)]), "ref_jpeg", prelude="from torch import nn\n")
    _quiet_tqdm(ns)
    return ns


def load_0409():
    """experiments/code/0409_method.ipynb: cell 0 lines 1-19 + 44-82 + 320-368 (codec, colour loss, SVD, phase)
    and cell 1 lines 389-449 (GaussianMixtureSampler with self.model)."""
    _install_stubs()
    c0 = _nb_cell("experiments/code/0409_method.ipynb", 0)
    c1 = _nb_cell("experiments/code/0409_method.ipynb", 1)
    src = _slice(c0, [(1, 19), (44, 82), (320, 368)]) + "\n" + _slice(c1, [(389, 449)])
    ns = _exec(src, "ref_0409", prelude="from torch import nn\n")
    _quiet_tqdm(ns)
    return ns


def load_conv_deep_color_loss():
    """experiments/code/conv_deep.ipynb cell 0 lines 60-73 (`color_loss`)."""
    _install_stubs()
    c0 = _nb_cell("experiments/code/conv_deep.ipynb", 0)
    ns = _exec(_slice(c0, [(60, 73)]), "ref_conv_deep",
               prelude="import torch\nimport torch.nn.functional as F\nfrom torch import nn\n")
    return ns


def load_avif_loss(ssim_fn):
    """avif.py lines 126-164 (`avif_frequency_aware_loss`), executed verbatim.  Its `ssim` is pytorch_msssim's (not installed,
    no version pinned by the reference): the caller passes the stand-in, so everything BUT the SSIM term is pinned by this."""
    _install_stubs()
    lines = open(os.path.join(REF_ROOT, "avif.py"), encoding="utf-8").read().split("\n")
    ns = _exec(_slice(lines, [(126, 164)]), "ref_avif_loss", prelude="import torch\nimport torch.nn.functional as F\n")
    ns["ssim"] = ssim_fn
    return ns


def load_webp_loss(ssim_fn):
    """webp_training.py lines 105-132 (`frequency_aware_loss`), executed verbatim with the SSIM stand-in (see load_avif_loss)."""
    _install_stubs()
    lines = open(os.path.join(REF_ROOT, "webp_training.py"), encoding="utf-8").read().split("\n")
    ns = _exec(_slice(lines, [(105, 132)]), "ref_webp_loss", prelude="import torch\nimport torch.nn.functional as F\n")
    ns["ssim"] = ssim_fn
    return ns


def load_dct_processor():
    """experiments/code/dct.ipynb cell 2 lines 43-139 (`DCTProcessor`: the pure-torch JPEG simulator with quant tables)."""
    _install_stubs()
    c2 = _nb_cell("experiments/code/dct.ipynb", 2)
    # The reference calls torch.cos() on Python floats (L81, L96), which raises TypeError -- the notebook's own run stopped
    # there.  To execute the algorithm as written, the class sees a `torch` whose cos() wraps scalars in tensors; every
    # other attribute is the real torch.
    prelude = ("import numpy as np\nimport torch as _torch\n"
               "class _TorchScalarCos:\n"
               "    def __getattr__(self, name):\n        return getattr(_torch, name)\n"
               "    @staticmethod\n    def cos(v):\n        return _torch.cos(_torch.as_tensor(v, dtype=_torch.float64))\n"
               "torch = _TorchScalarCos()\n")
    ns = _exec(_slice(c2, [(43, 139)]), "ref_dct", prelude=prelude)
    return ns


def load_0409_model():
    """experiments/code/0409_method.ipynb cell 0 lines 1-19 (imports, device) + 84-318 + 371-428: TimeEmbedding, DCTLayer, HFCM,
    FrequencyAwareBlock, ResAttnBlock and the notebook's JPEGDiffusionModel (the SVD/GMM solver's native UNet)."""
    _install_stubs()
    c0 = _nb_cell("experiments/code/0409_method.ipynb", 0)
    ns = _exec(_slice(c0, [(1, 19), (84, 318), (371, 428)]), "ref_0409_model")
    return ns
