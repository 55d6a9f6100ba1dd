"""Loss functions with the reference's names, as fused CUDA reductions.  These return the loss VALUE as a plain tensor
(no autograd history: nothing in this package uses autograd); the matching gradient kernels are in ops_train
(`frequency_aware_loss_backward`, `color_preservation_loss_backward`, `huber_color_loss_backward`) and are what
training.Trainer calls for the training step (webp_training.py:476-537).

  color_loss                conv_deep.ipynb#c0:L60-73        0.25 L1_R + 0.5 L1_G + 0.25 L1_B on clamped [0,1] images
  color_preservation_loss   0409_method.ipynb#c0:L64-82      color_loss + 0.5 (1 - SSIM)
  frequency_aware_loss      webp_training.py:105-132         MSE + 0.5 sum_c [MSE |rfft2| + 0.5 MSE angle rfft2] + 0.3 (1 - SSIM)
  avif_frequency_aware_loss avif.py:126-164                  MSE + 0.3 sum_c [MSE |fft2| + 0.3 MSE angle fft2] + 0.4 (1 - SSIM) + 0.2 edge

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


def avif_frequency_aware_loss(pred, target):
    """avif.py:126-164: spatial MSE + 0.3 * sum over channels of [MSE |fft2| + 0.3 MSE angle fft2] + 0.4 (1 - SSIM) + 0.2 *
    gradient_loss (MSE of the absolute vertical / horizontal neighbour differences), on the [0,1] images."""
    pred, target = _prep(pred, target)
    B, C, H, W = pred.shape
    if C != 3:
        raise ValueError("avif_frequency_aware_loss sums over the 3 colour channels")
    spatial = ops.mse(pred, target)
    terms = ops.freq_loss_terms(pred, target, full_spectrum=True)   # sums over all B*3 planes, full spectrum
    freq = (terms[0] + 0.3 * terms[1]).float() / float(B * H * W)   # one channel's fft2 has B*H*W coefficients
    edge = ops.edge_loss_terms(pred, target)
    edge_loss = (edge[0] / float(B * C * (H - 1) * W) + edge[1] / float(B * C * H * (W - 1))).float()
    return spatial + 0.3 * freq + 0.4 * (1.0 - ops.ssim(pred, target, clamp01=False)) + 0.2 * edge_loss


def huber_loss(pred, target, delta=1.0):
    """nn.HuberLoss(reduction='mean', delta=1.0), the `huber_loss_fn` of 0409_method.ipynb#c0:L438."""
    pred, target = pred.contiguous().float(), target.contiguous().float()
    if pred.shape != target.shape:
        raise ValueError("huber_loss: shapes differ")
    return ops.huber(pred, target, delta)


def color_weight_for_epoch(epoch):
    """0409_method.ipynb#c0:L573: the colour term's weight grows with the epoch, min(1.0, 0.2 + 0.02 * epoch)."""
    return min(1.0, 0.2 + epoch * 0.02)


def huber_color_loss(pred_noise, noise, xt, x0, epoch=0):
    """The 0409 notebook's training loss (#c0:L567-574): huber(pred_noise, x0 - xt) + w(epoch) * color_preservation_loss(xt +
    pred_noise, x0).  Returns (loss, huber, colour) like the loop that logs both terms."""
    h = huber_loss(pred_noise, noise)
    col = color_preservation_loss(ops.lincomb(xt.contiguous().float(), 1.0, pred_noise.contiguous().float(), 1.0), x0)
    return h + color_weight_for_epoch(epoch) * col, h, col
