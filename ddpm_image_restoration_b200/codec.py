"""Host codec round trip -- the sampler's data-consistency operator (webp_inference.py:506-528,
avif_inference.py:64-98, svd.ipynb#c1:L20-44).

The codec itself stays Pillow's libwebp / libavif / libjpeg-turbo (bit-exactness with those libraries is the
parity requirement), but around it the reference's serial per-image Python loop and fp32 transfers are replaced by
  * on-device truncating uint8 quantisation + NCHW->HWC (ddpmir_quantize_u8_hwc), so only 1 byte/sample crosses PCIe,
  * pinned staging buffers and asynchronous copies,
  * a thread pool over images (Pillow releases the GIL inside encode/decode),
  * raw uint8 decoder output consumed directly by the fused update kernel (ddpmir_ddrm_update, codec_u8_hwc=1).
Unlike the reference's avif_compress there is no silent JPEG fallback: an AVIF failure raises.
"""
import concurrent.futures as cf
import io
import os
import sys
import threading
import time

import numpy as np
import torch

from . import ops

_POOL = None
_POOL_THREADS = None
_AVIF_KW = None           # extra save() arguments of the AVIF encoder, decided once per process (see _avif_save_kwargs)
CPU_SECONDS = [0.0]       # wall time spent inside _roundtrip_one, summed over the pool's threads (bench.py: codec_cpu_ms_per_step)


def host_threads():
    n = os.environ.get("DDPMIR_CODEC_THREADS")
    if n:
        return max(1, int(n))
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def _codec_thread_init():
    """Codec threads (and the encoder workers they spawn, which inherit it) yield to the thread that launches the GPU kernels: with
    every core busy encoding, a launcher thread that has to queue for a core shows up as idle gaps on the GPU.  SCHED_IDLE where
    the kernel grants it (any normal-priority thread preempts at once), nice 10 otherwise."""
    tid = threading.get_native_id()
    try:
        os.sched_setscheduler(tid, os.SCHED_IDLE, os.sched_param(0))
        return
    except (OSError, AttributeError):
        pass
    try:
        os.setpriority(os.PRIO_PROCESS, tid, 10)
    except (OSError, AttributeError):  # pragma: no cover
        pass


def set_threads(n):
    """Resize the codec thread pool (bench.py divides the host cores between ranks)."""
    global _POOL, _POOL_THREADS
    if _POOL is not None:
        _POOL.shutdown(wait=True)
    _POOL = cf.ThreadPoolExecutor(max_workers=max(1, int(n)), thread_name_prefix="ddpmir-codec", initializer=_codec_thread_init)
    _POOL_THREADS = max(1, int(n))
    # The codec threads run short Python sections (Pillow's save/open wrappers) between long GIL-free encodes; with the default
    # 5 ms switch interval the thread that launches GPU work queues for the interpreter lock behind them (measured: 10 ms per
    # micro-batch launch that takes 0.2 ms uncontended).  A short interval hands the lock over promptly.
    if sys.getswitchinterval() > 2e-4:
        sys.setswitchinterval(2e-4)


def pool():
    if _POOL is None:
        set_threads(host_threads())
    return _POOL


def pool_threads():
    pool()
    return _POOL_THREADS


def _clamp_quality(codec, quality):
    q = int(quality)
    return max(0, min(100, q)) if codec == "webp" else max(1, min(100, q))


def _avif_save_kwargs():
    """Pillow's AVIF plugin hands libavif/libaom max_threads = the host's core count for EVERY save(), so a pool of N codec
    threads spawns N x cores encoder workers per batch; on 256x256 images they only cost (measured, 8 cores, 64 images:
    0.97 s per batch against 0.36 s with max_threads=2).  The bitstream does not depend on the worker count as long as it
    is >= 2 (1 switches libaom's row multi-threading off and changes the bytes), which is checked here once per process on
    a probe image against the default; if the installed libavif ever disagrees, the default stays."""
    global _AVIF_KW
    if _AVIF_KW is None:
        from PIL import Image
        yy, xx = np.mgrid[0:64, 0:64]
        probe = np.stack([(xx * 3 + yy * c) % 256 for c in (1, 2, 5)], -1).astype(np.uint8)
        probe[16:48, 16:48] = 255 - probe[16:48, 16:48]
        enc = []
        for kw in ({}, {"max_threads": 2}):
            buf = io.BytesIO()
            Image.fromarray(probe).save(buf, format="AVIF", quality=20, **kw)
            enc.append(buf.getvalue())
        _AVIF_KW = {"max_threads": 2} if enc[0] == enc[1] else {}
    return _AVIF_KW


def _roundtrip_one(codec, q, src, dst):
    """src, dst: [H, W, 3] uint8 numpy views (dst is written in place)."""
    from PIL import Image
    t0 = time.perf_counter()
    img = Image.fromarray(src, mode="RGB")
    buf = io.BytesIO()
    if codec == "webp":
        img.save(buf, format="WEBP", quality=q)
    elif codec == "avif":
        img.save(buf, format="AVIF", quality=q, **_avif_save_kwargs())
    elif codec == "jpeg":
        img.save(buf, format="JPEG", quality=q, subsampling="4:4:4" if q > 30 else "4:2:0")
    else:
        raise ValueError(f"unknown codec {codec!r}")
    buf.seek(0)
    dec = Image.open(buf)
    if dec.mode != "RGB":
        dec = dec.convert("RGB")
    dst[...] = np.asarray(dec, dtype=np.uint8)
    CPU_SECONDS[0] += time.perf_counter() - t0     # a float += under the GIL: good enough for a counter


def submit_roundtrip(codec, quality, src_u8, dst_u8):
    """Queue the round trip of every image of src_u8 [B,H,W,3] (numpy uint8) into dst_u8; returns the futures."""
    q = _clamp_quality(codec, quality)
    p = pool()
    return [p.submit(_roundtrip_one, codec, q, src_u8[i], dst_u8[i]) for i in range(src_u8.shape[0])]


def roundtrip_u8(codec, quality, src_u8):
    dst = np.empty_like(src_u8)
    for f in submit_roundtrip(codec, quality, src_u8, dst):
        f.result()
    return dst


def _compress(x, quality, codec):
    """Drop-in for {webp,avif,jpeg}_compress(x, quality): x in [-1,1], NCHW; returns fp32 on x's device."""
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("expected an [B,3,H,W] image batch")
    if x.is_cuda:
        u8 = ops.quantize_u8_hwc(x.contiguous().float())
        if codec == "jpeg":
            # same bytes as the Pillow round trip below (ddpmir_jpeg_roundtrip_u8), without leaving the device
            return ops.u8_hwc_to_nchw(ops.jpeg_roundtrip_u8(u8, _clamp_quality(codec, quality)))
        host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(u8, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        dec = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
        for f in submit_roundtrip(codec, quality, host.numpy(), dec.numpy()):
            f.result()
        return ops.u8_hwc_to_nchw(dec.to(x.device, non_blocking=True))
    # host tensors: the whole operator is host work
    u8 = (x.float() * 127.5 + 127.5).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().numpy()
    dec = roundtrip_u8(codec, quality, u8)
    return torch.from_numpy(dec).permute(0, 3, 1, 2).contiguous().float().div(255.0).sub(0.5).mul(2.0)


def webp_compress(x, quality):
    return _compress(x, quality, "webp")


def avif_compress(x, quality):
    return _compress(x, quality, "avif")


def jpeg_compress(x, quality):
    return _compress(x, quality, "jpeg")


class DCTProcessor:
    """Drop-in for DCTProcessor (experiments/code/dct.ipynb#c2:L43-139), the reference's pure-torch JPEG simulator, whose
    scalar Python loops over every 8x8 block are replaced by one kernel (ddpmir_jpeg_dct_project).  Images are fp32 NCHW
    on the 0..255 scale, channel 0 uses the luma table and the others the chroma table, exactly as the reference does."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DCTProcessor runs on CUDA only (no CPU fallback)")

    def jpeg_compress(self, images, quality=50):
        from . import ops
        return ops.jpeg_dct_project(images.to(self.device).contiguous().float(), quality)
