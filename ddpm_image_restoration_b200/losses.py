"""Training-loss kernels exposed with the reference's names.

color_l1 is the channel-weighted L1 term shared by color_preservation_loss (0409_method.ipynb#c0:L64-82) and
color_loss (conv_deep.ipynb#c0:L60-73), as one fused reduction kernel.  The SSIM term of color_preservation_loss
(third-party pytorch_msssim) and frequency_aware_loss (webp_training.py:105-132) are SURVEY section 8(f) "next"
rows and are not built yet -- asking for them raises instead of silently computing something else.
"""
from . import ops


def color_loss(pred, target):
    """conv_deep.ipynb#c0:L60-73: 0.25*L1_R + 0.5*L1_G + 0.25*L1_B on clamped [0,1] images (forward value)."""
    return ops.color_l1(pred.contiguous().float(), target.contiguous().float())


def color_preservation_loss(pred, target, include_ssim=True):
    """0409_method.ipynb#c0:L64-82.  Only the colour term is implemented (include_ssim=False)."""
    if include_ssim:
        raise NotImplementedError("the SSIM term (pytorch_msssim) is not built yet; pass include_ssim=False for the "
                                  "channel-weighted L1 term")
    return color_loss(pred, target)


def frequency_aware_loss(pred, target):
    raise NotImplementedError("frequency_aware_loss (webp_training.py:105-132) is a later-round item (SURVEY 8(f))")
