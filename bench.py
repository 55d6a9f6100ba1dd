#!/usr/bin/env python
"""Headline benchmark: restored images/sec for the full DDRM restoration loop at 256x256 (BASELINE.json).

Workload at N=1 = BASELINE.json configs[1]: AVIF(q=20), batch 64 per GPU, 256x256, 75 sampler timesteps per
trajectory (init_t = clamp(100-q, 15, 75), avif_inference.py:534-535), bf16 operands, random-init weights,
synthetic images compressed offline with the real codec.  One "step" = one sampler timestep over the batch
(UNet forward -> uint8 quantise -> host codec round trip -> fused update [-> phase consistency]); the default
K = 75 timed steps are exactly one trajectory.  images/s = batch / (75 x ms_per_step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Weak scaling: every rank restores its own batch of 64 images (images are independent -> no collective on the
data path; torch.distributed is used only for the barrier and the max-over-ranks time).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRAJ_STEPS = {"avif": 75, "webp": 80, "jpeg": 80}   # clamp(100-q, lo, hi) at q = 20 / 10 / 10
QUALITY = {"avif": 20, "webp": 10, "jpeg": 10}
# analytic FLOPs per image per UNet forward at 256x256
UNET_GF = {"avif": 2605.05, "webp": 2564.31, "jpeg": 2564.31}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--family", default="avif", choices=["avif", "webp", "jpeg", "mixed"],
                    help="mixed = BASELINE configs[4]: a third of the batch per codec (JPEG q=10 / WebP q=10 / AVIF q=20), each on its own UNet")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--micro-batches", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch the UNet forwards eagerly instead of replaying CUDA graphs")
    ap.add_argument("--codec-threads", type=int, default=0, help="host codec pool size per rank (default: host cores / ranks)")
    ap.add_argument("--no-overlap", action="store_true", help="train workload: one all-reduce after the backward instead of buckets under it")
    ap.add_argument("--projection", default="auto", choices=["auto", "codec", "dct", "device"],
                    help="data-consistency step: auto (default: device JPEG codec for --family jpeg, host codec otherwise), the "
                         "reference's host codec, the bit-exact device JPEG round trip "
                         "(--family jpeg only) or the opt-in DCT-domain projection (DCTProcessor.jpeg_compress, SURVEY 8f-1)")
    ap.add_argument("--attn-expmode", type=int, default=None, help="tuning: 0 = fp32 ex2, 1 = packed bf16x2 ex2")
    ap.add_argument("--train-family", default="webp", choices=["webp", "jpeg", "avif"],
                    help="train workload: model family (BASELINE configs[3] is webp; avif = train_epoch_ddrm_avif, avif.py:528-590)")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (ncu --profile-from-start off captures exactly them)")
    ap.add_argument("--profile-ops", action="store_true", help="tuning: print CUDA-event time per GEMM/conv/attention shape")
    ap.add_argument("--attn-lin", type=int, default=None, help="tuning: largest polynomial set of the polynomial-kernel attention tier, -1 = off")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the short runs of BASELINE's other configurations that the default N=1 run appends as `secondary`")
    ap.add_argument("--workload", default="sample", choices=["sample", "train", "gmm"],
                    help="sample = the headline DDRM restoration loop (BASELINE configs[1]); gmm = BASELINE configs[2], the SVD-guided "
                         "GaussianMixtureSampler on a JPEG(q=10) batch; train = BASELINE configs[3], "
                         "one webp_training.py optimizer step per step (WebP UNet 64x64, batch 32/GPU, NCCL gradient all-reduce)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="wall budget of the reference arm")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(family, batch, res, seed):
    """Synthetic images compressed offline with the real codec -> y in [-1,1], fp32 NCHW (host)."""
    import torch
    from ddpm_image_restoration_b200 import codec
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(res).float(), torch.arange(res).float(), indexing="ij")
    ph = torch.rand(batch, 3, 1, 1, generator=g) * 6.28
    base = 127 + 100 * torch.sin(xx / 17 + ph) * torch.cos(yy / 23 + torch.arange(batch).view(-1, 1, 1, 1) * 0.5)
    img = (base + 8 * torch.randn(batch, 3, res, res, generator=g)).clamp(0, 255).to(torch.uint8)
    clean = img.float() / 255.0 * 2 - 1
    fn = {"avif": codec.avif_compress, "webp": codec.webp_compress, "jpeg": codec.jpeg_compress}[family]
    return fn(clean, QUALITY[family]).contiguous()


def cpu_step_seconds(family, res, threads):
    """One sampler timestep on ONE image with the CPU oracle: UNet forward + codec round trip + update."""
    import torch
    from oracle import restated as R
    from oracle import weights as W
    torch.set_num_threads(threads)
    sd = W.make_state_dict(family, 0)
    x = W.synthetic_images(1, res, res, seed=99)
    y = R.codec_roundtrip(x, QUALITY[family], R.DDRM[family]["codec"])
    t = torch.full((1,), 0.5)
    g = torch.Generator().manual_seed(1)

    def one():
        t0 = time.perf_counter()
        x_theta = R.unet_forward(sd, y, t, t.clone(), family)
        c = R.codec_roundtrip(x_theta, QUALITY[family], R.DDRM[family]["codec"])
        z = torch.randn(y.shape, generator=g)
        R.ddrm_update(x_theta, c, y, z, t, R.DDRM[family]["sigma"])
        return time.perf_counter() - t0
    return one


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot allocate its 68.7 GB
    attention scores at 256x256) on all host threads, one image x one timestep per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    fam = args.family
    threads = os.cpu_count() or 1
    one = cpu_step_seconds(fam, args.res, threads)
    K = args.steps if args.steps is not None else 2
    W = args.warmup
    t_first = one()                       # first warm-up doubles as the cost estimate
    done_warm = 1
    while done_warm < W and (done_warm + 1) * t_first < 0.25 * args.cpu_budget_s:
        one(); done_warm += 1
    k_fit = max(1, int((args.cpu_budget_s - done_warm * t_first) / max(t_first, 1e-3)))
    k_run = min(K, k_fit)
    times = [one() for _ in range(k_run)]
    s = sum(times) / len(times)
    traj = TRAJ_STEPS[fam]
    val = 1.0 / (traj * s)
    sample = (f"1 image x 1 sampler timestep per step (UNet fwd + {fam} round trip + update) at {args.res}x{args.res}, "
              f"restated oracle (SDPA attention), extrapolated x{traj} timesteps; timed {k_run} of {K} requested steps "
              f"within a {args.cpu_budget_s:.0f}s budget")
    line = {"impl": "reference", "metric": "restored images/sec (256^2, full DDPM loop)", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": k_run, "steps_requested": K, "warmup": done_warm, "ms_per_step": s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{fam}_inference.py DDRM sampling, {fam.upper()}(q={QUALITY[fam]}) {args.res}x{args.res}, "
                                   f"{traj} timesteps/trajectory, CPU", "batch_per_step": 1},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def train_line(args, dev, rank, world, steps=None):
    """BASELINE configs[3]: webp_training.py training step, bf16 operands, batch 256 sharded over 8 GPUs (32 per GPU,
    weak scaling), NCCL all-reduce of the flat fp32 gradient in buckets under the backward.  Needs the process group of a
    multi-rank run to exist already; every rank must call it.  Returns the line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200 import codec, ops
    from ddpm_image_restoration_b200.training import Trainer
    Bn, res = (args.batch if args.batch != 64 else 32), (args.res if args.res != 256 else 64)
    torch.manual_seed(0)
    tfam = getattr(args, "train_family", "webp")
    model = {"webp": P.WebPDiffusionModel, "jpeg": P.JPEGDiffusionModel, "avif": P.AVIFDiffusionModel}[tfam]().to(dev).set_precision("bf16")
    tr = Trainer(model, seed=rank, overlap_allreduce=not args.no_overlap)
    g = torch.Generator().manual_seed(100 + rank)
    x0 = (torch.rand(Bn, 3, res, res, generator=g) * 2 - 1)
    xt = {"webp": codec.webp_compress, "jpeg": codec.jpeg_compress, "avif": codec.avif_compress}[tfam](x0, 30).contiguous().pin_memory()
    x0 = x0.pin_memory()
    t = (torch.randint(1, 100, (Bn,), generator=g).float() / 100.0).pin_memory()
    K = steps if steps is not None else 10
    Wm = max(3, args.warmup)

    def one():
        a, b_, c = xt.to(dev, non_blocking=True), t.to(dev, non_blocking=True), x0.to(dev, non_blocking=True)
        return tr.train_step(a, b_, c)
    for _ in range(Wm):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ops.LAUNCHES[0]
    clocks = ClockSampler(dev.index); clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if getattr(args, "profile_range", False):
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(K):
        loss = one()
    lossv = float(loss)          # device -> host read of the step result
    e1.record()
    torch.cuda.synchronize()
    if getattr(args, "profile_range", False):
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    sync_err = 0.0
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        # replicas must still hold identical parameters after K averaged steps (checks the bucketed all-reduce)
        chk = torch.stack([p.detach().double().abs().sum() for p in model.parameters()]).sum().view(1)
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        sync_err = float((hi - lo) / hi)
    if rank != 0:
        return None
    val = world * Bn * K / (ms / 1e3)
    nparam = sum(p.numel() for p in model.parameters())
    script = {"webp": "webp_training.py", "jpeg": "svd.ipynb", "avif": "avif.py"}[tfam]
    return {"metric": f"training images/sec ({script} step, {res}^2)", "value": val, "unit": "images/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{script} training step, {tfam.upper()} UNet {res}x{res}, batch {Bn}/GPU, "
                                   f"{'avif_' if tfam == 'avif' else ''}frequency_aware_loss, clip+AdamW, fp32 gradient all-reduce (NCCL) in " + ("one piece after" if args.no_overlap else "buckets under") + " the backward",
                       "allreduce_bytes": nparam * 4, "allreduce_collectives_per_step": tr.buckets.collectives / (K + Wm)},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 2 * xt.numel() * 4 + Bn * 4,
                    "d2h_bytes_per_step": 4.0 / K},
            "gpu_launches": ops.LAUNCHES[0] - launches0, "clocks": clk, "loss": lossv, "replica_param_mismatch": sync_err}


def run_train(args):
    """`--workload train`: the training line on its own (secondary metric; the default run folds the same line into
    `secondary` at every N, so the driver's 1/2/4/8-GPU records carry the data-parallel training step too)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = train_line(args, dev, rank, world, args.steps)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- per-op accounting (one profiled step): algorithmic bytes / FLOPs of every op class that goes through ops._timed -------
def op_work(name, tag):
    """-> ("flops" | "bytes", amount) of one call: the ALGORITHMIC work, not what the kernel moves."""
    if name in ("conv3x3", "gemm"):
        B, H, W, K, N, io = tag
        return "flops", 2.0 * B * H * W * K * N * (9 if name == "conv3x3" else 1)
    if name.startswith("igemm_tc_stream_kernel"):
        B, H, W = tag                                               # amount is summed per call by summarize_ops
        return "bytes", 0.0
    if name == "attention":
        B, L, C, heads = tag
        return "flops", 4.0 * B * float(L) * L * C                 # QK^T + PV of the reference's softmax attention
    if name == "groupnorm_stats":
        B, HW, C, es = tag
        return "bytes", B * HW * C * es
    if name == "groupnorm_apply":
        B, HW, C, es, eo = tag
        return "bytes", B * HW * C * (es + eo)
    if name == "block_transform":
        B, HW, C, es, eo = tag
        return "bytes", B * HW * C * (es + eo)
    if name == "maxpool2":
        B, HW, C, es = tag
        return "bytes", B * HW * C * es * 1.25
    if name == "upsample2_concat":
        B, HW, C1, C2, es = tag
        return "bytes", B * HW * es * (C1 + 4 * C2 + 4 * (C1 + C2))
    if name == "avgpool_pyramid":
        B, HW, C, es = tag
        return "bytes", B * HW * C * es
    if name == "avif_combine":
        B, HW, C, eh, ex = tag
        return "bytes", B * HW * C * (eh + 4 * ex)
    if name == "conv_input":
        B, Cin, H, W, N, ks, eo = tag
        return "bytes", B * H * W * (Cin * 4 + N * eo)
    if name == "out_conv_tanh":
        B, HW, Cin, N, es = tag
        return "bytes", B * HW * (Cin * es + N * 4)
    if name == "ddrm_update":
        B, C, H, W, u8 = tag
        return "bytes", B * C * H * W * (12 + (1 if u8 else 4))
    if name == "quantize_u8_hwc":
        B, C, H, W = tag
        return "bytes", B * C * H * W * 5
    if name == "svd_lowrank":
        n, H, W, k = tag
        return "bytes", n * H * W * 8
    return "bytes", 0.0


def igemm_stream_kernel(name, tag, num_sms=148):
    """Name of the persistent weight-resident kernel a conv3x3/gemm call of this shape runs on, or None for the tiled kernel
    (mirror of the dispatch rule in csrc/conv_tc.cu: ddpmir_igemm_tc -- whole weight in shared memory, >= 2 pixel tiles per SM)."""
    B, H, W, K, N, io = tag
    taps = 9 if name == "conv3x3" else 1
    if N % 16 or N > 256 or N * taps * K * 2 > 112 * 1024 or -(-B * H * W // 128) < 2 * num_sms:
        return None
    return f"igemm_tc_stream_kernel<{128 if N <= 64 else 256 if N <= 128 else 512}>"


def summarize_ops(timed, pk):
    """{class: {...}} of one profiled step + the list of groups, largest first.  A group is an (op, shape) pair, except that
    every conv3x3/gemm call that runs on the persistent streaming kernel is grouped by KERNEL and activation extent: those
    layers (K <= 576, N <= 256 over >= 296 pixel tiles) are bound by moving their operands, so the group's work is the
    algorithmic BYTES of its calls (operand in, every output / residual / gate tensor once), not their FLOP."""
    classes, groups = {}, {}
    for name, lst in timed.items():
        for ms, tag in lst:
            kind, amount = op_work(name, tag)
            c = classes.setdefault(name, {"ms": 0.0, "calls": 0, "kind": kind, "work": 0.0})
            c["ms"] += ms; c["calls"] += 1; c["work"] += amount
            gkey, gkind, gamount = (name, tag), kind, amount
            if name in ("conv3x3", "gemm"):
                kern = igemm_stream_kernel(name, tag)
                if kern:
                    gkey, gkind, gamount = (kern, tag[:3]), "bytes", float(tag[0] * tag[1] * tag[2] * tag[5])
            g = groups.setdefault(gkey, {"ms": 0.0, "calls": 0, "kind": gkind, "work": 0.0})
            g["ms"] += ms; g["calls"] += 1; g["work"] += gamount
            g["work_total"] = g["work"]
    for c in list(classes.values()) + list(groups.values()):
        rate = c["work"] / max(c["ms"], 1e-9) * 1e3
        if c["kind"] == "flops":
            c["tflops"] = rate / 1e12
        else:
            c["gbs"] = rate / 1e9; c["frac_of_hbm_peak"] = rate / 1e9 / pk["hbm"]
        del c["work"]
    return classes, sorted(groups.items(), key=lambda kv: -kv[1]["ms"])


def traffic_from_profile(kernel_key):
    """DRAM bytes per launch of a kernel from an ncu --set full capture, if profiles/ncu_traffic.json has this exact kernel and
    shape (written by tools/ncu_traffic.py from the .ncu-rep; carries its own provenance).  Never a literal in this file."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.isfile(p):
        return None, None
    d = json.load(open(p)).get(kernel_key)
    return (d["dram_bytes"], d["source"]) if d else (None, None)


class SampleJob:
    """One codec family's share of the batch: model, sampler, synthetic degraded images (host, pinned)."""

    def __init__(self, fam, batch, res, seed, dev, args):
        import ddpm_image_restoration_b200 as P
        self.fam, self.q, self.traj = fam, QUALITY[fam], TRAJ_STEPS[fam]
        model = {"avif": P.AVIFDiffusionModel, "webp": P.WebPDiffusionModel, "jpeg": P.JPEGDiffusionModel}[fam]()
        self.model = model.to(dev).eval().set_precision("bf16")
        self.cls = {"avif": P.DDRMAVIFSampler, "webp": P.DDRMWebPSampler, "jpeg": P.DDRMJPEGSampler}[fam]
        self.y_host = synth_batch(fam, batch, res, seed=seed).pin_memory()
        self.dev, self.args = dev, args
        self._sampler = None

    def sampler(self):
        """One sampler per job: it owns the captured CUDA graphs of the micro-batch forwards."""
        if self._sampler is None:
            self._sampler = self.cls(self.model, seed=7, micro_batches=self.args.micro_batches, projection=self.args.projection,
                                     use_graphs=not self.args.no_graphs)
        return self._sampler


def run_sample_steps(job, n_steps, warm, host_io, barrier, timed_filter=None):
    """Times n_steps timesteps (after `warm` untimed ones) -> (device ms, wall ms, stats).  host_io: the trajectory starts
    from the pinned host batch and ends with the restored batch copied back to the host (the e2e arm)."""
    import torch
    from ddpm_image_restoration_b200 import codec, ops
    sampler = job.sampler()
    traj, q, dev, y_host = job.traj, job.q, job.dev, job.y_host
    st = sampler.begin(y_host.to(dev, non_blocking=True), q, steps=traj)
    i = traj - 1
    for _ in range(warm):
        # the e2e arm starts a fresh trajectory inside the timed region: nothing may be pre-enqueued for it
        sampler.step(st, i, prefetch=not host_io); i = (i - 1) % traj
    if host_io:
        i = traj - 1
    barrier()
    ops.LAUNCHES[0] = 0
    st["h2d"] = st["d2h"] = 0; st["codec_s"] = 0.0; st["enqueue_s"] = 0.0
    codec.CPU_SECONDS[0] = 0.0
    if timed_filter is not None:
        ops.timing_begin(timed_filter)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = getattr(job.args, "profile_range", False) and not host_io
    if prof:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    e0.record()
    if host_io:
        st = sampler.begin(y_host.to(dev, non_blocking=True), q, steps=traj)   # the user-facing call starts from host images
        st["h2d"] += y_host.numel() * 4
    for n in range(n_steps):
        sampler.step(st, i, prefetch=not (host_io and n == n_steps - 1)); i = (i - 1) % traj
    if not host_io:        # where the trajectory stands now: the state the per-op profile (profile_one_step) is taken on
        job.profile_x, job.profile_i = st["x_t"], i
    if host_io:
        out_host = torch.empty(y_host.shape, dtype=torch.float32, pin_memory=True)
        out_host.copy_(st["x_t"], non_blocking=True); st["d2h"] += y_host.numel() * 4
    e1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    wall = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    timed = ops.timing_end() if timed_filter is not None else {}
    return ms, wall, dict(launches=ops.LAUNCHES[0], h2d=st["h2d"], d2h=st["d2h"], codec_s=st["codec_s"], timed=timed,
                          codec_cpu_s=codec.CPU_SECONDS[0], enqueue_s=st.get("enqueue_s", 0.0), finite=bool(torch.isfinite(st["x_t"]).all()))


def profile_one_step(job, barrier):
    """The GPU work of one sampler timestep with a CUDA-event pair around EVERY op of libddpmir, outside the headline timing and
    WITHOUT the host codec: the UNet forward of every micro-batch, the uint8 quantisation and the fused update on a stand-in for
    the decoded pixels.  (Inside the real step the codec threads keep every core busy and the launcher thread's delays would be
    booked on whatever small kernel waits for its launch.)  Also returns the polynomial-tier verdicts of every tiered attention
    call, which the roofline needs to count the arithmetic that was actually executed."""
    import torch
    from ddpm_image_restoration_b200 import ops
    sampler = job.sampler()
    # the iterate the timed run ended on (the attention tiers depend on the data: the degraded input itself would flatter them)
    x = getattr(job, "profile_x", None)
    x = job.y_host.to(job.dev) if x is None else x.clone()
    i0 = getattr(job, "profile_i", job.traj - 1)
    chunks = sampler._chunks(x.shape[0])
    cfg_sigma = 0.15 if job.fam == "avif" else 0.2

    def gpu_step(i):
        for s_, e_ in chunks:
            t = torch.full((e_ - s_,), float(i) / job.traj, dtype=torch.float32, device=x.device)
            x_theta = job.model(x[s_:e_], t, t)
            u8 = ops.quantize_u8_hwc(x_theta)
            ops.ddrm_update(x_theta, u8, x[s_:e_], t, cfg_sigma, seed=7, step=i)
    with torch.no_grad():
        gpu_step(i0)
        barrier()
        ops.timing_begin(lambda name, tag: True)
        ops.TIER_LOG = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gpu_step(i0); e1.record()
        barrier()
    tiers, ops.TIER_LOG = ops.TIER_LOG, None
    return e0.elapsed_time(e1), ops.timing_end(), tiers


# features (padded to 128) of the polynomial-kernel tier's monomial maps, attn_lin_tc.cu: {head_dim: {degree: F}}
LIN_FEATURES = {8: {3: 256, 4: 640, 5: 1408, 6: 3200}, 16: {3: 1024, 4: 5504}}
LIN_SET_DEGREE = [3, 3, 3, 4, 4, 5, 6]


def attention_executed_flops(tiers):
    """Multiply-adds x 2 the attention calls of the profiled step really executed on the tensor cores: per (image, head), the
    polynomial tier's two contractions (features x keys x N and rows x features x N, N = 16 / 32 value columns) or, where it
    declined, the quadratic tier's 4 L^2 head_dim."""
    total, hist = 0.0, {}
    for B, L, C, heads, verdicts in tiers:
        hd = C // heads
        nb = 16 if hd == 8 else 32
        for v in verdicts.flatten().tolist():
            hist[v] = hist.get(v, 0) + 1
            total += 4.0 * L * L * hd if v < 0 else 2.0 * 2.0 * L * LIN_FEATURES[hd][LIN_SET_DEGREE[v]] * nb
    return total, hist


def run_gmm(args, dev, steps_sample=None):
    """BASELINE configs[2]: the SVD-guided GaussianMixtureSampler (0409_method.ipynb#c1:L390-449) on a JPEG(q=10) batch at
    256x256, 91 timesteps (init_t + 1), SVD low-rank guide on timesteps i > 45 (rank int(256 * i / 91)), phase consistency every
    5th step.  A step = one timestep over the batch; timesteps with and without the SVD guide are timed separately and weighted
    45 : 46 as in the full trajectory (steps_sample = how many of each kind are timed; None = the whole trajectory)."""
    import torch
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200 import ops
    B, res, steps = args.batch, args.res, 91
    torch.manual_seed(0)
    model = P.JPEGDiffusionModel().to(dev).eval().set_precision("bf16")
    y = synth_batch("jpeg", B, res, seed=4321).to(dev)
    smp = P.GaussianMixtureSampler(model, num_timesteps=100, seed=7)
    st = smp.begin(y, steps=steps)
    with_svd = list(range(steps - 1, steps // 2, -1))
    without = list(range(steps // 2, -1, -1))
    if steps_sample is not None:
        with_svd, without = with_svd[:steps_sample], without[:steps_sample]
    for i in (steps - 1, steps - 2, steps // 2):      # warm-up (both kinds)
        smp.step(st, i)
    torch.cuda.synchronize()
    ops.LAUNCHES[0] = 0
    ops.timing_begin(lambda name, tag: name == "svd_lowrank")
    res_ms = {}
    for kind, lst in (("svd", with_svd), ("plain", without)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in lst:
            smp.step(st, i)
        e1.record(); torch.cuda.synchronize()
        res_ms[kind] = e0.elapsed_time(e1) / len(lst)
    svd = ops.timing_end().get("svd_lowrank", [])
    n_svd, n_plain = steps - 1 - steps // 2, steps // 2 + 1
    traj_ms = n_svd * res_ms["svd"] + n_plain * res_ms["plain"]
    svd_ms = sum(m for m, _ in svd) / max(1, len(svd))
    return {"metric": "restored images/sec (256^2, SVD-guided GMM solver)", "value": B / (traj_ms / 1e3), "unit": "images/s",
            "ms_per_step": traj_ms / steps, "ms_per_step_with_svd": res_ms["svd"], "ms_per_step_without_svd": res_ms["plain"],
            "svd_ms_per_call": svd_ms, "svd_planes": 3 * B, "svd_share_of_trajectory": n_svd * svd_ms / traj_ms,
            "steps_timed": len(with_svd) + len(without), "gpu_launches": ops.LAUNCHES[0],
            "finite": bool(torch.isfinite(st["x_t"]).all()),
            "config": {"workload": f"0409_method.ipynb GaussianMixtureSampler, JPEG(q=10) batch {B} {res}x{res}, 91 timesteps, "
                                   "SVD guide on i > 45, phase consistency every 5th step, device JPEG; JPEG UNet of svd.ipynb"}}


def run_secondary(args, dev, pk):
    """Short runs of BASELINE's other configurations, folded into the headline line (each also runs stand-alone:
    --family webp --batch 1, --workload gmm, --workload train, --family mixed --res 512)."""
    import copy
    import torch
    out = {}

    def guarded(name, fn):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:                      # a secondary line must never take the headline down
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        out[name]["bench_wall_s"] = time.perf_counter() - t0
        torch.cuda.empty_cache()

    def sample_line(fam, batch, res, k):
        a = copy.copy(args); a.micro_batches = None if batch < 16 else args.micro_batches
        fams = ["jpeg", "webp", "avif"] if fam == "mixed" else [fam]
        tot_img, tot_s, per = 0, 0.0, {}
        for f in fams:
            job = SampleJob(f, batch, res, 99, dev, a)
            ms, _, stt = run_sample_steps(job, k, 3, False, lambda: torch.cuda.synchronize())
            per[f] = {"ms_per_step": ms / k, "images_per_s": batch / (job.traj * ms / k / 1e3), "finite": stt["finite"]}
            tot_img += batch; tot_s += job.traj * ms / k / 1e3
            del job
            torch.cuda.empty_cache()
        return {"value": tot_img / tot_s, "unit": "images/s", "per_family": per, "steps_timed": k,
                "config": {"workload": f"DDRM sampling {fam} batch {batch}/family {res}x{res}"}}

    guarded("config1_webp_q10_256_single_image", lambda: sample_line("webp", 1, 256, 10))
    def gmm():
        a = copy.copy(args); a.batch, a.res = 64, 256
        return run_gmm(a, dev, steps_sample=3)
    guarded("config3_gmm_svd_jpeg_q10_256_b64", gmm)
    guarded("config5_mixed_512_b4_per_family", lambda: sample_line("mixed", 4, 512, 3))
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        return run_train(args)

    import torch
    import torch.distributed as dist
    from ddpm_image_restoration_b200 import codec, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    host_cores = os.cpu_count() or 1
    # codec pool per rank = the rank's share of the host cores.  (Leaving a quarter of them to the launcher / stager threads gained
    # 3 % on a 16-vCPU box at N = 1 but lost 18 % on a 24-core box at N = 2: 9 instead of 12 threads per rank, 335 vs 273 ms/step.)
    codec.set_threads(args.codec_threads if args.codec_threads else max(1, host_cores // world))
    if args.attn_expmode is not None:
        from ddpm_image_restoration_b200 import _lib
        _lib.lib().ddpmir_attention_set_expmode(args.attn_expmode)
    if args.attn_lin is not None:
        from ddpm_image_restoration_b200 import _lib
        _lib.lib().ddpmir_attention_set_lin(args.attn_lin)
    pk = peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "gmm":
        line = run_gmm(args, dev, steps_sample=args.steps)
        v = torch.tensor([line["value"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        if rank == 0:
            line.update({"value": float(v.item()), "n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                         "dtype": "bf16", "data": "synthetic"})
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    torch.manual_seed(0)
    fams = ["jpeg", "webp", "avif"] if args.family == "mixed" else [args.family]
    per_fam = args.batch // len(fams)
    jobs = [SampleJob(f, per_fam, args.res, 1234 + rank, dev, args) for f in fams]
    B_total = per_fam * len(fams)
    Wm = max(3, args.warmup)
    H = Wd = args.res

    if args.profile_ops and rank == 0:
        step_ms, timed, _ = profile_one_step(jobs[-1], barrier)
        classes, groups = summarize_ops(timed, pk)
        tot = sum(c["ms"] for c in classes.values())
        print(f"# per-op CUDA-event times of the GPU work of one sampler step ({step_ms:.1f} ms total, {tot:.1f} ms inside libddpmir ops)", file=sys.stderr)
        for (name, tag), g_ in groups[:48]:
            rate = f"{g_['tflops']:8.1f} TFLOP/s" if "tflops" in g_ else f"{g_['gbs']:8.0f} GB/s"
            print(f"#   {g_['ms']:9.3f} ms  x{g_['calls']:3d}  {rate}  {name:16s} {tag}", file=sys.stderr)

    clocks = ClockSampler(local)
    clocks.start()
    ms_f, wall_f, stats_f = [], [], []
    for job in jobs:
        K = args.steps if args.steps is not None else job.traj
        ms, wall, stt = run_sample_steps(job, K, Wm, False, barrier)
        ms_f.append(reduce_max(max(ms, 0.0)) / K); wall_f.append(wall / K); stats_f.append(stt)
    clk = clocks.stop()
    e2e_f, stats_e2e = [], []
    for job in jobs:
        K = args.steps if args.steps is not None else job.traj
        ms_e, wall_e, stt = run_sample_steps(job, K, 1, True, barrier)
        e2e_f.append(reduce_max(max(wall_e, ms_e)) / K); stats_e2e.append(stt)
    K = args.steps if args.steps is not None else jobs[0].traj
    # images / second: every family's share of the batch needs traj_f timesteps of ms_f each
    value = world * B_total / sum(j.traj * m / 1e3 for j, m in zip(jobs, ms_f))
    e2e_value = world * B_total / sum(j.traj * m / 1e3 for j, m in zip(jobs, e2e_f))
    ms_per_step = sum(ms_f)

    # one profiled timestep (after the timed runs): every libddpmir op between CUDA events -> roofline objects
    roof = roof_convs = roof_attn = hbm_table = roof_top = None
    if rank == 0:
        job = jobs[-1]
        step_ms, timed, tiers = profile_one_step(job, lambda: torch.cuda.synchronize())
        classes, groups = summarize_ops(timed, pk)
        conv_ms = sum(classes[n]["ms"] for n in ("conv3x3", "gemm") if n in classes)
        conv_fl = sum(classes[n]["tflops"] * classes[n]["ms"] for n in ("conv3x3", "gemm") if n in classes)   # TFLOP/s * ms = GFLOP
        if conv_ms > 0:
            roof_convs = {"bound": "tensor", "kernel": "igemm_tc_kernel / igemm_tc_stream_kernel (every conv3x3 and 1x1 GEMM of one timestep)",
                          "achieved": conv_fl / conv_ms, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                          "frac": conv_fl / conv_ms / pk["tf_sustained"], "ms_per_step": conv_ms,
                          "gflop_per_step": conv_fl, "launches": sum(classes[n]["calls"] for n in ("conv3x3", "gemm") if n in classes)}
        exec_fl, set_hist = attention_executed_flops(tiers)
        if "attention" in classes:
            a = classes["attention"]
            roof_attn = {"ms_per_step": a["ms"], "launches": a["calls"], "reference_equivalent_tflops": a["tflops"],
                         "polynomial_sets": {str(k): v for k, v in sorted(set_hist.items())},
                         "note": "reference_equivalent = 4 L^2 C FLOP of the reference's softmax attention per call / measured time.  "
                                 "(image, head) pairs whose logit bound fits a polynomial set (attn_lin*.cu) cost O(L F head_dim) "
                                 "instead, so this is an equivalent rate, not tensor-pipe utilisation; polynomial_sets counts the "
                                 "verdicts of the tiered calls (-1 = quadratic tiers)"}
        hbm_table = {n: {"ms_per_step": c["ms"], "calls": c["calls"], "gbs": c["gbs"], "frac_of_hbm_peak": c["frac_of_hbm_peak"]}
                     for n, c in classes.items() if c["kind"] == "bytes" and c["ms"] > 0 and c.get("gbs", 0) > 0}
        # the dominant kernel = the (op, shape) group with the most time in the step
        (dname, dtag), dg = groups[0]
        key = f"{dname}{list(dtag)}"
        traffic, tsrc = traffic_from_profile(key)
        kind, work = op_work(dname, dtag)
        avg_ms = dg["ms"] / dg["calls"]
        if dname.startswith("igemm_tc_stream_kernel"):
            work = dg["work_total"] / dg["calls"]            # mean algorithmic bytes per launch of this kernel in the step
        if dname == "attention":
            # executed tensor-core arithmetic of exactly these calls (their verdicts), not the reference's 4 L^2 C
            mine = [t_ for t_ in tiers if (t_[0], t_[1], t_[2], t_[3]) == tuple(dtag)]
            if mine:
                work = attention_executed_flops(mine)[0] / len(mine)
        if kind == "flops":
            ach = work / (avg_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "peak_source": pk["src"] + " (sustained bf16 GEMM)"}
        else:
            ach = work / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                    "peak_source": pk["src"] + " (copy bandwidth)"}
        roof.update({"kernel": key, "launch_ms": avg_ms, "launches_per_step": dg["calls"], "share_of_step": dg["ms"] / step_ms,
                     "traffic": traffic, "traffic_source": tsrc, "algorithmic_work": work,
                     "profiled_step_ms": step_ms})
        if dname.startswith("igemm_tc_stream_kernel"):
            roof["note"] = ("persistent weight-resident tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu) over every conv3x3 / 1x1 GEMM of "
                            "the full-resolution levels that it serves (the launches differ in K, N and epilogue tensors): achieved = "
                            "sum of their algorithmic bytes (bf16 operand in + each output / residual / gate tensor once) / sum of "
                            "their times; launch_ms and algorithmic_work are per-launch means.  The tensor-roofline view of all "
                            "convs is roofline_convs; the attention calls are roofline_attention")
        if dname == "attention":
            roof["reference_equivalent_tflops"] = op_work(dname, dtag)[1] / (avg_ms * 1e-3) / 1e12
            roof["note"] = ("one attention call of the full-resolution blocks = pre-pass + polynomial-kernel tier (attn_lin_tc.cu: "
                            "attn_lin_state_tc_kernel, attn_lin_reduce_tc_kernel, attn_lin_out_tc_kernel) + the quadratic tiers for what "
                            "it declines.  achieved = the tensor-core FLOP these kernels EXECUTED (2 x 2 L F N per (image, head), F "
                            "features, N value columns) / time: the kernels are bound by generating the features on the fp32 pipe "
                            "(one multiply + half a convert per feature), the tensor pipe follows.  reference_equivalent_tflops = the "
                            "reference's 4 L^2 C FLOP / the same time")

        # An attention call is a pipeline of ~15 kernels (pre-pass, per-degree state / reduce / output, quadratic tiers).  The
        # largest SINGLE kernel of the step (first in the ncu launch list, profiles/r2_step_launches_summary.txt) is the
        # streaming conv kernel: its group gets the same treatment, against the HBM roofline, with DRAM traffic from ncu.
        roof_top = None
        for (kname, ktag), kg in groups:
            if not kname.startswith("igemm_tc_stream_kernel"):
                continue
            kkey = f"{kname}{list(ktag)}"
            ktraffic, ksrc = traffic_from_profile(kkey)
            kwork, kms = kg["work_total"] / kg["calls"], kg["ms"] / kg["calls"]
            kach = kwork / (kms * 1e-3) / 1e9
            roof_top = {"bound": "hbm", "kernel": kkey, "achieved": kach, "peak": pk["hbm"], "unit": "GB/s", "frac": kach / pk["hbm"],
                        "peak_source": pk["src"] + " (copy bandwidth)", "launch_ms": kms, "launches_per_step": kg["calls"],
                        "share_of_step": kg["ms"] / step_ms, "algorithmic_work": kwork, "traffic": ktraffic, "traffic_source": ksrc,
                        "note": "every conv3x3 / 1x1 GEMM of this activation extent that runs on the persistent weight-resident tcgen05 "
                                "kernel (csrc/conv_tc.cu): achieved = sum of algorithmic bytes (bf16 operand in + each output / residual / "
                                "gate tensor once) / sum of times; launch_ms and algorithmic_work are per-launch means over launches that "
                                "differ in K, N and epilogue tensors, `traffic` is the ncu DRAM byte count of its heaviest member (a 64->64 "
                                "3x3 conv with fp32 residual and fp32 output, 671 MB algorithmic)"}
            break

    unet_tflops = UNET_GF[jobs[0].fam] * B_total / 1e3 / (ms_per_step / 1e3) if (args.res == 256 and len(jobs) == 1) else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fam = jobs[0].fam
        one = cpu_step_seconds(fam, args.res, host_cores)
        s = one()
        cpu = {"value": 1.0 / (TRAJ_STEPS[fam] * s), "unit": "images/s", "cores": host_cores, "kind": "port",
               "sample": f"1 image x 1 sampler timestep (UNet fwd + {fam} round trip + update) = {s:.2f} s, x{TRAJ_STEPS[fam]} timesteps"}
    secondary = None
    if rank == 0 and world == 1 and args.secondary and args.family == "avif" and args.res == 256:
        secondary = run_secondary(args, dev, pk)

    if rank == 0:
        fam_txt = "+".join(f"{j.fam.upper()}(q={j.q})" for j in jobs)
        line = {
            "metric": "restored images/sec (256^2, full DDPM loop)" if args.res == 256 else f"restored images/sec ({args.res}^2, full DDPM loop)",
            "value": value, "unit": "images/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"DDRM sampling ({'/'.join(j.fam for j in jobs)}_inference.py): batch {B_total}/GPU {fam_txt} {H}x{Wd}, "
                                   f"{'/'.join(str(j.traj) for j in jobs)} timesteps/trajectory, step = one timestep over the batch",
                       "batch_per_gpu": B_total, "timesteps_per_trajectory": [j.traj for j in jobs], "micro_batches": args.micro_batches,
                       "projection": args.projection,
                       "codec_threads_per_rank": codec.pool_threads(), "host_cores": host_cores,
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": sum(s_["h2d"] for s_ in stats_e2e) / K,
                    "d2h_bytes_per_step": sum(s_["d2h"] for s_ in stats_e2e) / K,
                    "note": "DDRM sampler public API from pinned host images to pinned host result; wall clock"},
            "gpu_launches": sum(s_["launches"] for s_ in stats_f), "clocks": clk, "roofline": roof, "roofline_top_kernel": roof_top, "roofline_convs": roof_convs,
            "roofline_attention": roof_attn, "hbm_kernels": hbm_table, "cpu_baseline": cpu,
            "unet_tflops": unet_tflops, "unet_frac_of_bf16_sustained": (unet_tflops / pk["tf_sustained"]) if unet_tflops else None,
            "codec_wait_ms_per_step": sum(s_["codec_s"] for s_ in stats_f) * 1e3 / K,
            "codec_cpu_ms_per_step": sum(s_["codec_cpu_s"] for s_ in stats_f) * 1e3 / K,
            "launch_host_ms_per_step": sum(s_["enqueue_s"] for s_ in stats_f) * 1e3 / K,
            "wall_ms_per_step": sum(wall_f), "finite": all(s_["finite"] for s_ in stats_f), "secondary": secondary,
        }
        # what bounds the step: the reference's data-consistency term is a host codec round trip of every image at every
        # timestep (avif_inference.py:438 / webp_inference.py:581); the pool's thread-time per step / its threads is a floor on
        # the step no GPU kernel can lower.  gpu_ms_per_step = the libddpmir work of one step, profiled on the iterate the timed run ended on (rank 0).
        cthreads = max(1, codec.pool_threads())
        floor = line["codec_cpu_ms_per_step"] / cthreads
        gpu_ms = roof["profiled_step_ms"] if roof else None
        line["limiter"] = {"codec_thread_ms_per_step": line["codec_cpu_ms_per_step"], "codec_threads": cthreads,
                           "host_codec_floor_ms_per_step": floor, "gpu_ms_per_step": gpu_ms,
                           "bound": "host codec (Pillow round trip mandated by the reference's sampler; scales with host cores per GPU, not with GPUs)"
                                    if floor > 0.75 * ms_per_step else "gpu"}
    else:
        line = None
    if args.secondary and args.family == "avif" and args.res == 256 and args.batch == 64:
        # BASELINE configs[3] (data-parallel training step) at THIS run's N, so the 1/2/4/8-GPU records carry the gradient
        # all-reduce over NVLink too.  Every rank takes part; a watchdog prints the headline without it if a rank gets stuck.
        import threading
        TRAIN_KEY = "config4_train_webp64_b32_per_gpu"

        def give_up():
            if rank == 0:
                line["secondary"] = dict(line["secondary"] or {}, **{TRAIN_KEY: {"error": "timed out after 240 s"}})
                print(json.dumps(line), flush=True)
            os._exit(0)
        dog = threading.Timer(240.0, give_up); dog.daemon = True; dog.start()
        del jobs, job
        gc.collect(); torch.cuda.empty_cache()
        t0 = time.perf_counter()
        try:
            tl = train_line(args, dev, rank, world)
        except Exception as e:                      # a secondary line must never take the headline down
            tl = {"error": f"{type(e).__name__}: {e}"}
        dog.cancel()
        if rank == 0:
            tl["bench_wall_s"] = time.perf_counter() - t0
            line["secondary"] = dict(line["secondary"] or {}, **{TRAIN_KEY: tl})
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
