"""Samplers with the reference's API:

  DDRMWebPSampler / DDRMAVIFSampler / DDRMJPEGSampler(model).sample(x_t, quality, steps=100, eta=0.85, eta_b=1.0)
      webp_inference.py:553-602, avif_inference.py:410-459, svd.ipynb#c1:L336-385
  GaussianMixtureSampler(model, num_timesteps=100).sample(x_t, steps=100, use_phase_consistency=True,
      use_svd_guide=True, guidance_scale=1.0)                          0409_method.ipynb#c1:L390-449

Per timestep the reference runs the UNet, a serial host codec loop with fp32 transfers, ~8 element-wise kernels and
a Philox kernel.  Here one step is: UNet forward (libddpmir kernels) -> on-device uint8 quantisation -> pinned D2H ->
thread-pooled codec -> pinned H2D of raw pixels -> ONE fused update kernel (codec dequantisation, data-consistency,
eta mixing, in-kernel Philox noise) -> optional phase-consistency FFT kernels.  The batch is split into micro-batches
so the host codec of one micro-batch overlaps the UNet of the next.

Noise: by default generated in-kernel (Philox4x32-10 keyed by seed/step/element -- reproducible, independent of launch
geometry).  For parity runs pass `noise_fn(i, like) -> tensor` (and `coin_fn(i) -> float` for the GMM sampler), which
replaces torch.randn_like / torch.rand(1) of the reference.
"""
import torch

from . import codec as _codec
from . import ops

_DDRM = {
    "webp": dict(codec="webp", sigma=0.2, q_thr=15, period=5, alpha=0.7),   # webp_inference.py:588,595-597
    "jpeg": dict(codec="jpeg", sigma=0.2, q_thr=20, period=5, alpha=0.7),   # svd.ipynb#c1:L371,378-380
    "avif": dict(codec="avif", sigma=0.15, q_thr=30, period=3, alpha=0.8),  # avif_inference.py:445,452-454
}


_STAGERS = None


def _stagers(n):
    """Threads that wait (for a CUDA event, then for the codec futures of one micro-batch): blocked almost all the time."""
    global _STAGERS
    import concurrent.futures as cf
    if _STAGERS is None or _STAGERS._max_workers < n:
        _STAGERS = cf.ThreadPoolExecutor(max_workers=max(n, 8), thread_name_prefix="ddpmir-stage")
    return _STAGERS


class _DDRMSampler:
    family = None

    def __init__(self, model, seed=0, micro_batches=None, noise_fn=None, projection="auto", use_graphs=True):
        """projection: "auto" (default) = "device" where it exists (JPEG family), else
        "codec"; "codec" = the reference's host codec round trip (Pillow); "device" = JPEG family only: the same
        round trip computed on the GPU with libjpeg-turbo's integer arithmetic (ddpmir_jpeg_roundtrip_u8) -- bit-identical
        pixels, no host hop; "dct" = opt-in DCT-domain projection as the reference's DCTProcessor defines it
        (dct.ipynb#c2:L100-139) -- NOT libjpeg, results differ from the codec path and are checked by PSNR."""
        if projection not in ("auto", "codec", "dct", "device"):
            raise ValueError(projection)
        if projection == "device" and self.family != "jpeg":
            raise ValueError("projection='device' exists for the JPEG codec only (WebP and AVIF stay on the host)")
        self.model = model
        self.seed = seed
        self.micro_batches = micro_batches
        self.noise_fn = noise_fn
        self.projection = projection
        self.last_stats = {}
        # CUDA graphs of the per-micro-batch GPU work (UNet forward + uint8 quantisation): a forward is ~450 kernel launches, and
        # with the codec pool keeping every core busy the launching thread took 45 ms to issue what the GPU executes in 32 ms
        self.use_graphs = use_graphs
        self._graphs = {}

    def _chunks(self, B):
        n = self.micro_batches
        if n is None:
            n = 4 if B >= 16 else (2 if B >= 4 else 1)
        n = max(1, min(n, B))
        size = (B + n - 1) // n
        return [(s, min(B, s + size)) for s in range(0, B, size)]

    def begin(self, x_t, quality, steps=100, eta=0.85, eta_b=1.0):
        """Set up one trajectory (staging buffers, cached reference phasors); returns the mutable state."""
        cfg = _DDRM[self.family]
        if not x_t.is_cuda:
            raise RuntimeError("the B200 sampler runs on CUDA only (no CPU fallback)")
        self.model.eval()
        self.model.prepack()              # (re)validates the packed weights once per trajectory
        dev = x_t.device
        x_t = x_t.contiguous().float().clone()
        B, C, H, W = x_t.shape
        chunks = self._chunks(B)
        use_phase = quality < cfg["q_thr"]
        projection = self.projection
        if projection == "auto":
            projection = "device" if self.family == "jpeg" else "codec"
        st = dict(projection=projection, cfg=cfg, x_t=x_t, x_alt=torch.empty_like(x_t), y=x_t.clone(), quality=quality, steps=steps, eta=eta, eta_b=eta_b,
                  chunks=chunks, use_phase=use_phase, h2d=0, d2h=0, codec_s=0.0, pending=[None] * len(chunks), stage=[None] * len(chunks),
                  phasor=ops.phase_reference(x_t) if (use_phase and steps > cfg["period"]) else None)
        # staging: device uint8 buffers and pinned host buffers, one set per micro-batch
        mk = lambda s, e, **kw: torch.empty((e - s, H, W, C), dtype=torch.uint8, **kw)
        st["dev_u8"] = [mk(s, e, device=dev) for s, e in chunks]
        st["pin_src"] = [mk(s, e, pin_memory=True) for s, e in chunks]
        st["pin_dst"] = [mk(s, e, pin_memory=True) for s, e in chunks]
        st["dev_dec"] = [mk(s, e, device=dev) for s, e in chunks]
        # blocking events: a stager that waits for its micro-batch sleeps instead of spinning on a core the codec needs
        st["events"] = [torch.cuda.Event(blocking=True) for _ in chunks]
        return st

    def _graph_for(self, st, k):
        """Captured GPU work of micro-batch k (static input/output buffers), or None when graphs are off / capture failed.
        Cached across trajectories; rebuilt when the model's packed weights (whose addresses the graph holds) change."""
        if not self.use_graphs:
            return None
        s, e = st["chunks"][k]
        x_t = st["x_t"]
        key = (k, e - s, tuple(x_t.shape[1:]), x_t.device, self.model.precision)
        ent = self._graphs.get(key)
        # hot path: no Python loop over the parameters here (begin() has validated the packed weights for this trajectory; the
        # launching thread shares the interpreter lock with the codec threads)
        if ent is not None and ent.get("packed") is self.model._packed and self.model._packed is not None:
            return ent
        if ent is not None and ent.get("failed"):
            return None
        packed = self.model.prepack()
        try:
            dev = x_t.device
            # inputs AND outputs live outside the graphs' shared memory pool: inside it only intermediates, which are dead when a
            # replay ends (an output placed there could be overwritten by another micro-batch's replay before it is consumed)
            ent = dict(packed=packed, x=torch.empty((e - s,) + tuple(x_t.shape[1:]), dtype=torch.float32, device=dev),
                       x_theta=torch.empty((e - s,) + tuple(x_t.shape[1:]), dtype=torch.float32, device=dev),
                       t=torch.empty((e - s,), dtype=torch.float32, device=dev),
                       u8=torch.empty((e - s, x_t.shape[2], x_t.shape[3], x_t.shape[1]), dtype=torch.uint8, device=dev))
            ent["x"].copy_(x_t[s:e]); ent["t"].fill_(0.5)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side), torch.no_grad():          # warm-up outside the capture (one-time attribute / table set-up)
                ops.quantize_u8_hwc(self.model(ent["x"], ent["t"], ent["t"]), out=ent["u8"])
            torch.cuda.current_stream(dev).wait_stream(side)
            if "pool" not in self._graphs:
                self._graphs["pool"] = torch.cuda.graph_pool_handle()      # the micro-batches replay one after another: one pool
            g = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES[0]
            with torch.cuda.graph(g, pool=self._graphs["pool"]), torch.no_grad():
                ent["x_theta"].copy_(self.model(ent["x"], ent["t"], ent["t"]))
                ops.quantize_u8_hwc(ent["x_theta"], out=ent["u8"])
            ent["launches"] = ops.LAUNCHES[0] - n0
            ent["graph"] = g
            self._graphs[key] = ent
            return ent
        except Exception as exc:                                     # stay correct: eager launches
            import warnings
            warnings.warn(f"CUDA graph capture of the sampler's UNet forward failed ({exc!r}); launching eagerly")
            self._graphs[key] = dict(failed=True)
            return None

    def _enqueue_unet(self, st, k, i):
        """GPU work of micro-batch k at timestep i: UNet forward, uint8 quantisation, asynchronous D2H, event."""
        import time
        t_enq = time.perf_counter()
        s, e = st["chunks"][k]
        x = st["x_t"][s:e]
        ent = self._graph_for(st, k) if st["projection"] == "codec" else None
        if ent is not None:
            ent["x"].copy_(x)
            ent["t"].fill_(float(i) / st["steps"])
            ent["graph"].replay()
            ops.LAUNCHES[0] += ent["launches"]
            x_theta, t, u8 = ent["x_theta"], ent["t"], ent["u8"]
        else:
            t = torch.full((e - s,), float(i) / st["steps"], dtype=torch.float32, device=x.device)
            x_theta = self.model(x, t, t)
            u8 = ops.quantize_u8_hwc(x_theta, out=st["dev_u8"][k])
        st["pin_src"][k].copy_(u8, non_blocking=True)
        st["events"][k].record()
        st["d2h"] += u8.numel()
        st["pending"][k] = (x_theta, t, i)
        st["enqueue_s"] = st.get("enqueue_s", 0.0) + time.perf_counter() - t_enq    # host time to launch one micro-batch's kernels

    def _stage(self, st, k):
        """Host side of micro-batch k, run by a stager thread: wait for its pixels to arrive in pinned memory, push them through
        the codec pool, wait for the last image.  Returns how long the codec part took."""
        import time
        st["events"][k].synchronize()
        t0 = time.perf_counter()
        for f in _codec.submit_roundtrip(st["cfg"]["codec"], st["quality"], st["pin_src"][k].numpy(), st["pin_dst"][k].numpy()):
            f.result()
        return time.perf_counter() - t0

    def _launch(self, st, k, i):
        """GPU work of micro-batch k for timestep i, followed by its host stage."""
        self._enqueue_unet(st, k, i)
        st["stage"][k] = (i, _stagers(len(st["chunks"])).submit(self._stage, st, k))

    def step(self, st, i, prefetch=True):
        """One sampler timestep i (steps-1 ... 0) over the whole batch: webp_inference.py:566-600.

        Micro-batches are independent trajectories, and each runs as its own chain UNet -> pinned D2H -> codec pool -> pinned H2D
        -> fused update -> UNet of the next timestep ...  The chains are event-driven: a stager thread per micro-batch waits for
        the pixels and for the codec, the launching thread takes whichever micro-batch finishes first, updates it and at once
        launches its next UNet forward (a CUDA graph replay) and stage -- so the GPU works on one micro-batch while the host
        cores encode the others, across timestep boundaries.  A call returns when every micro-batch has reached timestep i."""
        import concurrent.futures as cf
        import time
        cfg, y, chunks = st["cfg"], st["y"], st["chunks"]
        B, C, H, W = st["x_t"].shape
        if st["projection"] != "codec":
            return self._step_dct(st, i)
        with torch.no_grad():
            for k in range(len(chunks)):
                if st["stage"][k] is None or st["stage"][k][0] != i:
                    self._launch(st, k, i)
            x_cur, x_new = st["x_t"], st["x_alt"]
            todo = {st["stage"][k][1]: k for k in range(len(chunks))}
            while todo:
                t0 = time.perf_counter()
                done, _ = cf.wait(list(todo), return_when=cf.FIRST_COMPLETED)
                st["codec_s"] += time.perf_counter() - t0              # launcher thread blocked on the host stages
                for fut in done:
                    k = todo.pop(fut)
                    st["stage_s"] = st.get("stage_s", 0.0) + fut.result()
                    s, e = chunks[k]
                    st["dev_dec"][k].copy_(st["pin_dst"][k], non_blocking=True)
                    st["h2d"] += st["pin_dst"][k].numel()
                    x_theta, t, _ = st["pending"][k]
                    st["pending"][k] = None
                    st["stage"][k] = None
                    z = None
                    if self.noise_fn is not None and i > 0:
                        z = self.noise_fn(i, x_cur)[s:e].contiguous()
                    # the flat NCHW element index inside the FULL batch keys the noise (noise_offset), so the result
                    # does not depend on the micro-batch split
                    ops.ddrm_update(x_theta, st["dev_dec"][k], y[s:e], t, cfg["sigma"], st["eta"], st["eta_b"], z=z,
                                    last_step=(i == 0), seed=self.seed, step=i, out=x_new[s:e], noise_offset=s * C * H * W)
                    if i > 0 and st["use_phase"] and i % cfg["period"] == 0:
                        x_new[s:e].copy_(ops.phase_consistency_cached(x_new[s:e], st["phasor"][s * C:e * C], cfg["alpha"]))
                    if prefetch and i > 0:
                        # x_new is the current state of micro-batch k from here on
                        st["x_t"], st["x_alt"] = x_new, x_cur
                        self._launch(st, k, i - 1)
                        st["x_t"], st["x_alt"] = x_cur, x_new
        st["x_t"], st["x_alt"] = x_new, x_cur
        return st["x_t"]

    def _step_dct(self, st, i):
        """The same timestep with a device-side data-consistency operator instead of the host codec."""
        cfg, y = st["cfg"], st["y"]
        B, C, H, W = st["x_t"].shape
        x_cur, x_new = st["x_t"], st["x_alt"]
        with torch.no_grad():
            for s, e in st["chunks"]:
                t = torch.full((e - s,), float(i) / st["steps"], dtype=torch.float32, device=x_cur.device)
                x_theta = self.model(x_cur[s:e], t, t)
                if st["projection"] == "device":     # raw decoder bytes, exactly what the host codec would hand back
                    proj = ops.jpeg_roundtrip_u8(ops.quantize_u8_hwc(x_theta), st["quality"])
                else:
                    proj = ops.jpeg_dct_project(x_theta, st["quality"], 127.5, 127.5)
                z = None
                if self.noise_fn is not None and i > 0:
                    z = self.noise_fn(i, x_cur)[s:e].contiguous()
                ops.ddrm_update(x_theta, proj, y[s:e], t, cfg["sigma"], st["eta"], st["eta_b"], z=z, last_step=(i == 0),
                                seed=self.seed, step=i, out=x_new[s:e], noise_offset=s * C * H * W)
                if i > 0 and st["use_phase"] and i % cfg["period"] == 0:
                    x_new[s:e].copy_(ops.phase_consistency_cached(x_new[s:e], st["phasor"][s * C:e * C], cfg["alpha"]))
        st["x_t"], st["x_alt"] = x_new, x_cur
        return st["x_t"]

    def sample(self, x_t, quality, steps=100, eta=0.85, eta_b=1.0):
        st = self.begin(x_t, quality, steps, eta, eta_b)
        for i in range(steps - 1, -1, -1):
            self.step(st, i)
        self.last_stats = dict(h2d_bytes=st["h2d"], d2h_bytes=st["d2h"], steps=steps, micro_batches=len(st["chunks"]),
                               codec_threads=_codec.pool_threads(), codec_seconds=st["codec_s"])
        return st["x_t"]


class DDRMWebPSampler(_DDRMSampler):
    family = "webp"


class DDRMJPEGSampler(_DDRMSampler):
    family = "jpeg"


class DDRMAVIFSampler(_DDRMSampler):
    family = "avif"


def phase_consistency(x, ref, alpha=0.7):
    """Drop-in for phase_consistency(x, ref, alpha), webp_inference.py:531-550."""
    return ops.phase_consistency_cached(x.contiguous().float(), ops.phase_reference(ref.contiguous().float()), alpha)


def svd_structure_preservation(x, k_ratio=0.5):
    """Drop-in for svd_structure_preservation(x, k_ratio), 0409_method.ipynb#c0:L321-346."""
    h, w = x.shape[-2:]
    k = max(1, int(min(h, w) * k_ratio))
    return ops.svd_lowrank(x.contiguous().float(), k)


class GaussianMixtureSampler:
    def __init__(self, model, num_timesteps=100, seed=0, noise_fn=None, coin_fn=None):
        self.model = model
        self.num_timesteps = num_timesteps
        self.seed = seed
        self.noise_fn = noise_fn
        self.coin_fn = coin_fn

    def begin(self, x_t, steps=100, use_phase_consistency=True, use_svd_guide=True, guidance_scale=1.0):
        """Set up one trajectory; returns the mutable state that step() advances (bench.py times single steps)."""
        if not x_t.is_cuda:
            raise RuntimeError("the B200 sampler runs on CUDA only (no CPU fallback)")
        self.model.eval()
        x_t = x_t.contiguous().float().clone()
        y = x_t.clone()
        return dict(x_t=x_t, y=y, steps=steps, use_phase=use_phase_consistency, use_svd=use_svd_guide, scale=guidance_scale,
                    phasor=ops.phase_reference(y) if use_phase_consistency else None)

    def step(self, st, i):
        """One timestep i (steps-1 ... 0) of 0409_method.ipynb#c1:L406-447 over the whole batch."""
        x_t, y, steps = st["x_t"], st["y"], st["steps"]
        B, dev = x_t.shape[0], x_t.device
        with torch.no_grad():
            t = torch.full((B,), float(i) / self.num_timesteps, dtype=torch.float32, device=dev)
            pred = self.model(x_t, t, t)
            prior, g = None, 0.0
            if st["use_svd"] and i > steps // 2:
                k_ratio = i / steps
                prior = svd_structure_preservation(x_t, k_ratio)
                g = k_ratio * 0.3
            if i > 0:
                p_cons = max(0.2, min(0.8, i / steps))
                # the reference draws torch.rand(1) from the CPU generator (0409_method.ipynb#c1:L432)
                coin = self.coin_fn(i) if self.coin_fn is not None else torch.rand(1).item()
                z = self.noise_fn(i, x_t).contiguous() if self.noise_fn is not None else None
                x_n = ops.gmm_update(x_t, pred, y, prior, g, z=z, use_first=coin < p_cons,
                                     noise_scale=0.1 * i / steps * st["scale"], seed=self.seed, step=i)
                if st["use_phase"] and i % 5 == 0:
                    x_n = ops.phase_consistency_cached(x_n, st["phasor"], 0.6 + 0.3 * (1 - i / steps))
                st["x_t"] = x_n
            else:
                st["x_t"] = ops.gmm_update(x_t, pred, y, prior, g, last_step=True)
        return st["x_t"]

    def sample(self, x_t, steps=100, use_phase_consistency=True, use_svd_guide=True, guidance_scale=1.0):
        st = self.begin(x_t, steps, use_phase_consistency, use_svd_guide, guidance_scale)
        for i in range(steps - 1, -1, -1):
            self.step(st, i)
        return st["x_t"]
