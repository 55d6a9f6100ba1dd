"""Host codec throughput on this box (no GPU work): wall time of 64 AVIF(q=20) 256x256 round trips through the thread pool for
several pool sizes -- the floor under the AVIF sampler's timestep (avif_inference.py:64-98 runs them serially).

    python tools/codec_bench.py [codec] [quality]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ddpm_image_restoration_b200 import codec

name = sys.argv[1] if len(sys.argv) > 1 else "avif"
q = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(0)
yy, xx = np.mgrid[0:256, 0:256]
imgs = np.stack([(127 + 100 * np.sin(xx / 17 + n)[..., None] * np.cos(yy / 23 + n * .5)[..., None] + rng.normal(0, 8, (256, 256, 3))).clip(0, 255).astype(np.uint8)
                 for n in range(64)])
cores = os.cpu_count()
print(f"{name} q={q}, 64 images 256x256, host cores {cores}, encoder kwargs {codec._avif_save_kwargs() if name == 'avif' else {}}")
for n in sorted({max(1, cores // 4), max(1, cores // 2), max(1, cores - 2), cores, cores * 3 // 2, cores * 2}):
    codec.set_threads(n)
    codec.roundtrip_u8(name, q, imgs[:n])
    codec.CPU_SECONDS[0] = 0.0
    t0 = time.perf_counter(); p0 = time.process_time()
    for _ in range(2):
        codec.roundtrip_u8(name, q, imgs)
    wall = (time.perf_counter() - t0) / 2; cpu = (time.process_time() - p0) / 2
    print(f"  pool {n:3d}: {wall * 1e3:7.1f} ms per 64 images  (process CPU {cpu * 1e3:7.1f} ms, in-thread wall {codec.CPU_SECONDS[0] / 2 * 1e3:7.1f} ms, {wall / 64 * 1e3:5.2f} ms/image)")
