"""B200-native (sm_100a) DDPM image-restoration sampler: drop-in for the PyTorch hot path of
Azure0413/DDPM_Image_Restoration (UNet forward, per-step update, data-consistency / guidance, colour loss).

Everything on the device runs through libddpmir.so (hand-written CUDA, C ABI in include/ddpmir.h); there is no
CPU or PyTorch-op fallback.  Build with `python -m ddpm_image_restoration_b200.build`.
"""
from .codec import DCTProcessor, avif_compress, jpeg_compress, webp_compress  # noqa: F401
from . import method0409  # noqa: F401  (the 0409 notebook's own UNet: method0409.JPEGDiffusionModel)
from .losses import avif_frequency_aware_loss, color_loss, color_preservation_loss, frequency_aware_loss  # noqa: F401
from .models import AVIFDiffusionModel, JPEGDiffusionModel, WebPDiffusionModel  # noqa: F401
from .samplers import (DDRMAVIFSampler, DDRMJPEGSampler, DDRMWebPSampler, GaussianMixtureSampler,  # noqa: F401
                       phase_consistency, svd_structure_preservation)

__all__ = ["WebPDiffusionModel", "JPEGDiffusionModel", "AVIFDiffusionModel", "DDRMWebPSampler", "DDRMJPEGSampler",
           "DDRMAVIFSampler", "GaussianMixtureSampler", "webp_compress", "avif_compress", "jpeg_compress", "DCTProcessor",
           "phase_consistency", "svd_structure_preservation", "color_loss", "color_preservation_loss",
           "frequency_aware_loss", "avif_frequency_aware_loss"]
