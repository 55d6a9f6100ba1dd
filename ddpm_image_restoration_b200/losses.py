"""Loss functions with the reference's names, as fused CUDA reductions.  FORWARD VALUES ONLY for now: the training
step (webp_training.py:476-537) needs the backward kernels of every UNet op, which are not built yet, so these return
plain tensors without autograd history.

  color_loss                conv_deep.ipynb#c0:L60-73        0.25 L1_R + 0.5 L1_G + 0.25 L1_B on clamped [0,1] images
  color_preservation_loss   0409_method.ipynb#c0:L64-82      color_loss + 0.5 (1 - SSIM)
  frequency_aware_loss      webp_training.py:105-132         MSE + 0.5 sum_c [MSE |rfft2| + 0.5 MSE angle rfft2] + 0.3 (1 - SSIM)

SSIM follows pytorch_msssim.ssim's published algorithm (gaussian window 11, sigma 1.5, valid filtering); the package is
not pinned by the reference nor installed here, so that term's parity is against the oracle's restatement only.
"""
from . import ops


def _prep(pred, target):
    if pred.shape != target.shape or pred.dim() != 4:
        raise ValueError("expected two [B,C,H,W] batches of the same shape")
    return pred.contiguous().float(), target.contiguous().float()


def color_loss(pred, target):
    pred, target = _prep(pred, target)
    return ops.color_l1(pred, target)


def color_preservation_loss(pred, target, include_ssim=True):
    pred, target = _prep(pred, target)
    loss = ops.color_l1(pred, target)
    if include_ssim:
        loss = loss + 0.5 * (1.0 - ops.ssim(pred, target, clamp01=True))
    return loss


def frequency_aware_loss(pred, target):
    pred, target = _prep(pred, target)
    B, C, H, W = pred.shape
    if C != 3:
        raise ValueError("frequency_aware_loss sums over the 3 colour channels")
    spatial = ops.mse(pred, target)
    terms = ops.freq_loss_terms(pred, target)                       # sums over all B*3 planes
    count = float(B * H * (W // 2 + 1))                             # elements of one channel's rfft2
    freq = (terms[0] + 0.5 * terms[1]).float() / count
    return spatial + 0.5 * freq + 0.3 * (1.0 - ops.ssim(pred, target, clamp01=False))
