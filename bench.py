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
    ap.add_argument("--family", default="avif", choices=["avif", "webp", "jpeg"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--micro-batches", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="train workload: one all-reduce after the backward instead of buckets under it")
    ap.add_argument("--projection", default="auto", choices=["auto", "codec", "dct", "device"],
                    help="data-consistency step: auto (default: device JPEG codec for --family jpeg, host codec otherwise), the "
                         "reference's host codec, the bit-exact device JPEG round trip "
                         "(--family jpeg only) or the opt-in DCT-domain projection (DCTProcessor.jpeg_compress, SURVEY 8f-1)")
    ap.add_argument("--attn-expmode", type=int, default=None, help="tuning: 0 = fp32 ex2, 1 = packed bf16x2 ex2")
    ap.add_argument("--profile-ops", action="store_true", help="tuning: print CUDA-event time per GEMM/conv/attention shape")
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample = the headline DDRM restoration loop (BASELINE configs[1]); train = BASELINE configs[3], "
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


def run_train(args):
    """BASELINE configs[3]: webp_training.py training step, bf16 operands, batch 256 sharded over 8 GPUs (32 per GPU,
    weak scaling), one NCCL all-reduce of the flat fp32 gradient per step.  Secondary line (not the headline metric)."""
    import torch
    import torch.distributed as dist
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200 import codec, ops
    from ddpm_image_restoration_b200.training import Trainer
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Bn, res = (args.batch if args.batch != 64 else 32), (args.res if args.res != 256 else 64)
    torch.manual_seed(0)
    model = P.WebPDiffusionModel().to(dev).set_precision("bf16")
    tr = Trainer(model, seed=rank, overlap_allreduce=not args.no_overlap)
    g = torch.Generator().manual_seed(100 + rank)
    x0 = (torch.rand(Bn, 3, res, res, generator=g) * 2 - 1)
    xt = codec.webp_compress(x0, 30).contiguous().pin_memory()
    x0 = x0.pin_memory()
    t = (torch.randint(1, 100, (Bn,), generator=g).float() / 100.0).pin_memory()
    K = args.steps if args.steps is not None else 10
    Wm = max(3, args.warmup)

    def one():
        a, b_, c = xt.to(dev, non_blocking=True), t.to(dev, non_blocking=True), x0.to(dev, non_blocking=True)
        return tr.train_step(a, b_, c)
    for _ in range(Wm):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ops.LAUNCHES[0] = 0
    clocks = ClockSampler(local); clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = one()
    lossv = float(loss)          # device -> host read of the step result
    e1.record()
    torch.cuda.synchronize()
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
    if rank == 0:
        val = world * Bn * K / (ms / 1e3)
        nparam = sum(p.numel() for p in model.parameters())
        print(json.dumps({"metric": "training images/sec (webp_training.py step, 64^2)", "value": val, "unit": "images/s", "n_gpus": world,
                          "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                          "config": {"workload": f"webp_training.py training step, WebP UNet {res}x{res}, batch {Bn}/GPU, "
                                                 "frequency_aware_loss, clip+AdamW, fp32 gradient all-reduce (NCCL) in " + ("one piece after" if args.no_overlap else "buckets under") + " the backward",
                                     "allreduce_bytes": nparam * 4},
                          "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 2 * xt.numel() * 4 + Bn * 4,
                                  "d2h_bytes_per_step": 4.0 / K},
                          "gpu_launches": ops.LAUNCHES[0], "clocks": clk, "loss": lossv, "replica_param_mismatch": sync_err}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        return run_train(args)

    import torch
    import torch.distributed as dist
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200 import codec, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fam = args.family
    traj = TRAJ_STEPS[fam]
    K = args.steps if args.steps is not None else traj
    Wm = max(3, args.warmup)
    host_cores = os.cpu_count() or 1
    codec.set_threads(max(1, host_cores // world))

    if args.attn_expmode is not None:
        from ddpm_image_restoration_b200 import _lib
        _lib.lib().ddpmir_attention_set_expmode(args.attn_expmode)
    torch.manual_seed(0)
    model = {"avif": P.AVIFDiffusionModel, "webp": P.WebPDiffusionModel, "jpeg": P.JPEGDiffusionModel}[fam]()
    model = model.to(dev).eval().set_precision("bf16")
    sampler_cls = {"avif": P.DDRMAVIFSampler, "webp": P.DDRMWebPSampler, "jpeg": P.DDRMJPEGSampler}[fam]
    y_host = synth_batch(fam, args.batch, args.res, seed=1234 + rank).pin_memory()
    B, C, H, Wd = y_host.shape
    q = QUALITY[fam]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n_steps, warm, host_io):
        """Times n_steps timesteps (after `warm` untimed ones) -> (device ms, stats).  host_io: the trajectory starts
        from the pinned host batch and ends with the restored batch copied back to the host (the e2e arm)."""
        sampler = sampler_cls(model, seed=7, micro_batches=args.micro_batches, projection=args.projection)
        y_dev = y_host.to(dev, non_blocking=True)
        st = sampler.begin(y_dev, q, steps=traj)
        i = traj - 1
        for _ in range(warm):
            # the e2e arm starts a fresh trajectory inside the timed region: nothing may be pre-enqueued for it
            sampler.step(st, i, prefetch=not host_io); i = (i - 1) % traj
        if host_io:
            i = traj - 1
        barrier()
        ops.LAUNCHES[0] = 0
        st["h2d"] = st["d2h"] = 0; st["codec_s"] = 0.0
        L_full = H * Wd
        ops.timing_begin(lambda name, tag: name == "attention" and tag[1] == L_full)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if host_io:
            st = sampler.begin(y_host.to(dev, non_blocking=True), q, steps=traj)   # the user-facing call starts from host images
            st["h2d"] += y_host.numel() * 4
        for n in range(n_steps):
            sampler.step(st, i, prefetch=not (host_io and n == n_steps - 1)); i = (i - 1) % traj
        if host_io:
            out_host = torch.empty(y_host.shape, dtype=torch.float32, pin_memory=True)
            out_host.copy_(st["x_t"], non_blocking=True); st["d2h"] += y_host.numel() * 4
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        attn = ops.timing_end().get("attention", [])
        return ms, wall, dict(launches=ops.LAUNCHES[0], h2d=st["h2d"], d2h=st["d2h"], codec_s=st["codec_s"], attn=attn,
                              finite=bool(torch.isfinite(st["x_t"]).all()))

    if args.profile_ops and rank == 0:
        sampler = sampler_cls(model, seed=7, micro_batches=args.micro_batches, projection=args.projection)
        st = sampler.begin(y_host.to(dev), q, steps=traj)
        for i in (traj - 1, traj - 2):
            sampler.step(st, i)
        torch.cuda.synchronize()
        ops.timing_begin(lambda name, tag: True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sampler.step(st, traj - 3); e1.record()
        torch.cuda.synchronize()
        agg = {}
        for name, lst in ops.timing_end().items():
            for ms, tag in lst:
                k = (name, tag)
                a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
        tot = sum(v[1] for v in agg.values())
        print(f"# per-op CUDA-event times of one sampler step ({e0.elapsed_time(e1):.1f} ms total, {tot:.1f} ms in GEMM/conv/attention)", file=sys.stderr)
        for (name, tag), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"#   {ms:9.3f} ms  x{n:3d}  {name:10s} {tag}", file=sys.stderr)

    clocks = ClockSampler(local)
    clocks.start()
    ms, wall, stats = run(K, Wm, host_io=False)
    clk = clocks.stop()
    ms_e2e, wall_e2e, stats_e2e = run(K, 1, host_io=True)

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = reduce_max(max(ms, 0.0)); ms_e2e = reduce_max(max(wall_e2e, ms_e2e))
    ms_per_step = ms / K
    value = world * B / (traj * ms_per_step / 1e3)
    e2e_value = world * B / (traj * (ms_e2e / K) / 1e3)

    # roofline of the dominant kernel: the full-resolution attention launches (L = H*W tokens)
    pk = peaks()
    heads = 8 if fam == "avif" else 4
    attn = stats["attn"]
    roof = None
    if attn:
        avg_ms = sum(m for m, _ in attn) / len(attn)
        bsz = attn[0][1][0]
        flops = 4.0 * (H * Wd) ** 2 * 64 * bsz      # 4 L^2 C per image (QK^T + PV), C = 64 at full resolution
        ach = flops / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": f"attn_tc_kernel<hd={64 // heads}> (tcgen05/TMEM bounded-softmax self-attention, full resolution, L={H * Wd})",
                "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                # dram__bytes_read+write of this launch from profiles/r1_attn_tc_ncu_metrics.csv (ncu --set full, same shape)
                "traffic": 542.1e6 if (fam == "avif" and bsz == 16 and H * Wd == 65536) else None,
                "algorithmic_bytes": bsz * H * Wd * 64 * 2 * 4,
                "peak_source": pk["src"] + " (sustained bf16 GEMM)", "launch_ms": avg_ms,
                "launches_timed": len(attn), "share_of_step": sum(m for m, _ in attn) / ms,
                "exp_pipe": {"scores_per_s": heads * (H * Wd) ** 2 * bsz / (avg_ms * 1e-3),
                             "mufu_ceiling_scores_per_s": 16 * 148 * (clk["sm_mhz"] or 1965.0) * 1e6,
                             "frac": heads * (H * Wd) ** 2 * bsz / (avg_ms * 1e-3) / (16 * 148 * (clk["sm_mhz"] or 1965.0) * 1e6)},
                "note": "this kernel is bound by the exp (MUFU) + issue pipes, not the tensor pipe: head_dim 8/16 gives 16-32 "
                        "FLOP per exp.  exp_pipe.frac is against the MUFU-only ceiling (16 exp2/clk/SM); the kernel evaluates "
                        "every other score pair on the FMA/ALU pipes (packed bf16), so it can exceed 1.  See DESIGN.md section 4 and "
                        "profiles/r1_ncu_summary.md"}
    unet_tflops = UNET_GF[fam] * B / 1e3 / (ms_per_step / 1e3) if args.res == 256 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        one = cpu_step_seconds(fam, args.res, host_cores)
        s = one()
        cpu = {"value": 1.0 / (traj * s), "unit": "images/s", "cores": host_cores, "kind": "port",
               "sample": f"1 image x 1 sampler timestep (UNet fwd + {fam} round trip + update) = {s:.2f} s, x{traj} timesteps"}

    if rank == 0:
        line = {
            "metric": "restored images/sec (256^2, full DDPM loop)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{fam}_inference.py DDRM sampling: batch {B}/GPU {fam.upper()}(q={q}) {H}x{Wd}, "
                                   f"{traj} timesteps/trajectory, step = one timestep over the batch",
                       "batch_per_gpu": B, "timesteps_per_trajectory": traj, "micro_batches": args.micro_batches,
                       "projection": args.projection,
                       "codec_threads_per_rank": codec.pool_threads(), "host_cores": host_cores,
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": stats_e2e["h2d"] / K,
                    "d2h_bytes_per_step": stats_e2e["d2h"] / K,
                    "note": "DDRM sampler public API from pinned host images to pinned host result; wall clock"},
            "gpu_launches": stats["launches"], "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "unet_tflops": unet_tflops, "unet_frac_of_bf16_sustained": (unet_tflops / pk["tf_sustained"]) if unet_tflops else None,
            "codec_wait_ms_per_step": stats["codec_s"] * 1e3 / K, "wall_ms_per_step": wall / K, "finite": stats["finite"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
