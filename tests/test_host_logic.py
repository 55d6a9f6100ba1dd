"""Host-side logic on the CPU: checkpoint layout, codec thread pool, sampler micro-batching, sharding (gloo, 2 ranks)."""
import os

import numpy as np
import pytest
import torch

from oracle import restated as R
from oracle import weights as W


@pytest.mark.parametrize("fam,cls,n", [("webp", "WebPDiffusionModel", 356), ("jpeg", "JPEGDiffusionModel", 356), ("avif", "AVIFDiffusionModel", 634)])
def test_checkpoint_layout(fam, cls, n):
    import ddpm_image_restoration_b200 as P
    m = getattr(P, cls)()
    sd = m.state_dict()
    assert len(sd) == n
    assert {k: tuple(v.shape) for k, v in sd.items()} == W.shapes(fam)
    m.load_state_dict(W.make_state_dict(fam, 0))                      # raw state_dict, strict
    wrapped = {"model_state_dict": W.make_state_dict(fam, 1), "epoch": 3}   # webp_training.py:796-804 layout
    m.load_state_dict(wrapped["model_state_dict"])
    assert sum(p.numel() for p in m.parameters()) == {"webp": 114398409, "jpeg": 114398409, "avif": 158284137}[fam]


def test_gradient_bucket_spans_follow_backward_order():
    """The bucketed all-reduce relies on the flat gradient buffer being laid out in registration order, so that the blocks
    the backward finishes (out_conv, up5 ... down1, time_embed) close contiguous tails of it."""
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200.training import grad_span_starts
    m = P.WebPDiffusionModel()
    spans = grad_span_starts(m)
    order = ["time_embed", "down1", "down2", "down3", "down4", "down5", "bottleneck.0", "bottleneck.1", "bottleneck.2",
             "up1", "up2", "up3", "up4", "up5", "out_conv"]
    assert list(spans) == order and spans["time_embed"] == 0
    starts = [spans[k] for k in order]
    assert starts == sorted(starts) and len(set(starts)) == len(starts)
    total = sum(p.numel() for p in m.parameters())
    assert starts[-1] < total
    # every parameter of a block lies inside [start(block), start(next block))
    off = 0
    for k, p in m.named_parameters():
        owner = ".".join(k.split(".")[:2]) if k.startswith("bottleneck.") else k.split(".")[0]
        i = order.index(owner)
        hi = starts[i + 1] if i + 1 < len(starts) else total
        assert starts[i] <= off and off + p.numel() <= hi
        off += p.numel()


def test_checkpoint_layout_m0409():
    """0409 notebook's JPEGDiffusionModel: 282 entries, 119 873 161 parameters."""
    from ddpm_image_restoration_b200 import method0409
    m = method0409.JPEGDiffusionModel()
    sd = m.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == W.shapes("m0409") and len(sd) == 282
    m.load_state_dict(W.make_state_dict("m0409", 0))
    assert sum(p.numel() for p in m.parameters()) == 119873161
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 3, 32, 32), torch.zeros(1))     # CPU tensors: no fallback


@pytest.mark.parametrize("codec,q", [("webp", 10), ("webp", 0), ("avif", 20), ("jpeg", 10), ("jpeg", 50), ("jpeg", 150)])
def test_codec_pool_matches_oracle(codec, q):
    from ddpm_image_restoration_b200 import codec as C
    x = W.synthetic_images(5, 32, 48, seed=3)
    fn = {"webp": C.webp_compress, "avif": C.avif_compress, "jpeg": C.jpeg_compress}[codec]
    assert torch.equal(fn(x, q), R.codec_roundtrip(x, q, codec))
    C.set_threads(2)
    assert C.pool_threads() == 2
    assert torch.equal(fn(x, q), R.codec_roundtrip(x, q, codec))
    C.set_threads(C.host_threads())


def test_codec_rejects_bad_input():
    from ddpm_image_restoration_b200 import codec as C
    with pytest.raises(ValueError):
        C.webp_compress(torch.zeros(3, 8, 8), 10)
    with pytest.raises(ValueError):
        C.roundtrip_u8("gif", 10, np.zeros((1, 8, 8, 3), dtype=np.uint8))


def test_sampler_micro_batching():
    import ddpm_image_restoration_b200 as P
    s = P.DDRMAVIFSampler(model=None)
    assert s._chunks(64) == [(0, 16), (16, 32), (32, 48), (48, 64)]
    assert s._chunks(1) == [(0, 1)]
    assert s._chunks(5) == [(0, 3), (3, 5)]
    s.micro_batches = 3
    ch = s._chunks(10)
    assert ch[0][0] == 0 and ch[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))


def test_shard_range():
    from ddpm_image_restoration_b200.parallel import codec_threads_per_rank, shard_range
    for n, world in ((512, 8), (10, 4), (3, 8), (64, 1)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert codec_threads_per_rank(16, 8) == 2 and codec_threads_per_rank(4, 8) == 1


def test_cosine_warm_restarts_matches_torch():
    """parallel.CosineWarmRestarts (what Trainer.scheduler_step drives) against the scheduler the reference constructs
    (webp_training.py:776: CosineAnnealingWarmRestarts(optimizer, T_0=100, T_mult=2)), over three restarts; and a short cycle."""
    from ddpm_image_restoration_b200.parallel import CosineWarmRestarts
    for base, T0, Tm, epochs in ((2e-4, 100, 2, 750), (1.5e-4, 3, 2, 40), (1e-3, 5, 1, 23)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=base)
        ref = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=T0, T_mult=Tm)
        mine = CosineWarmRestarts(base, T0, Tm)
        assert mine.lr() == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12)
        for _ in range(epochs):
            opt.step(); ref.step()
            assert mine.step() == pytest.approx(opt.param_groups[0]["lr"], rel=1e-9, abs=1e-18)
    st = mine.state_dict()
    other = CosineWarmRestarts(9.0); other.load_state_dict(st)
    assert other.step() == mine.step()


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from ddpm_image_restoration_b200.parallel import (GradBuckets, allreduce_mean_, init_process_group_from_env, max_over_ranks,
                                                      shard_range, sum_over_ranks)
    assert init_process_group_from_env() == (rank, world, 0) and dist.get_backend() == "gloo"
    flat = torch.full((1000,), float(rank + 1))      # the training step's flat gradient buffer
    allreduce_mean_(flat)
    assert torch.allclose(flat, torch.full((1000,), 1.5))
    # the Trainer's bucketed, overlapped gradient all-reduce (GradBuckets): tails of the flat buffer become final in the
    # order the backward finishes the blocks; the result must equal ONE all-reduce of the whole buffer, bit for bit
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(10007, generator=g)
    single = grad.clone()
    allreduce_mean_(single)
    bk = GradBuckets(grad, bucket_bytes=4 * 1500)
    for lo in (9000, 8800, 7000, 6999, 3000, 100):       # block starts, descending
        bk.final_from(lo)
    assert bk.pending, "some buckets must already be in flight before finish()"
    try:
        bk.reset()
        raise AssertionError("reset() must refuse to drop pending all-reduces")
    except RuntimeError:
        pass
    bk.finish()
    assert bk.collectives == 4 and not bk.pending and bk.hi == 0      # [7000,10007) [3000,7000) [100,3000) + the head [0,100)
    assert torch.equal(grad, single)
    bk.reset()
    lo, hi = shard_range(10, rank, world)
    dist.barrier()
    t = max_over_ranks(1.0 + rank)            # the slowest rank defines the step time
    n = sum_over_ranks(hi - lo)               # units all ranks processed
    q.put((rank, t, n, lo, hi))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [2.0, 2.0]      # max over ranks
    assert [r[2] for r in res] == [10.0, 10.0]    # every unit processed exactly once
    assert (res[0][3], res[0][4], res[1][3], res[1][4]) == (0, 5, 5, 10)


def test_bench_op_grouping_and_stream_kernel_rule():
    """bench.py's per-op accounting: conv3x3 / gemm calls that the persistent streaming kernel serves are grouped by KERNEL with
    their algorithmic bytes (HBM roofline), everything else by (op, shape) with FLOP or bytes; the dispatch rule mirrors
    csrc/conv_tc.cu (whole weight <= 112 KB in shared memory, N % 16 == 0, N <= 256, >= 2 pixel tiles per SM)."""
    import bench
    full = (16, 256, 256, 64, 64, 64 * 2 + 64 * 8)           # B, H, W, K, N, bytes per pixel (bf16 in, fp32 out + fp32 residual)
    assert bench.igemm_stream_kernel("conv3x3", full) == "igemm_tc_stream_kernel<128>"
    assert bench.igemm_stream_kernel("gemm", (16, 256, 256, 64, 192, 0)) == "igemm_tc_stream_kernel<512>"
    assert bench.igemm_stream_kernel("conv3x3", (16, 256, 256, 128, 64, 0)) is None      # 9 * 128 * 64 * 2 B = 147 KB of weights
    assert bench.igemm_stream_kernel("conv3x3", (16, 8, 8, 1024, 1024, 0)) is None       # N > 256
    assert bench.igemm_stream_kernel("gemm", (1, 64, 64, 64, 64, 0)) is None             # 32 tiles < 2 per SM
    timed = {"conv3x3": [(0.25, full), (0.35, full)], "gemm": [(0.1, (16, 8, 8, 1024, 1024, 4096))],
             "groupnorm_apply": [(0.05, (16, 65536, 64, 4, 2))]}
    classes, groups = bench.summarize_ops(timed, {"hbm": 6552.3, "tf_sustained": 1373.5})
    (name, tag), g = groups[0]
    assert name == "igemm_tc_stream_kernel<128>" and tag == (16, 256, 256) and g["calls"] == 2 and g["kind"] == "bytes"
    px = 16 * 256 * 256
    assert g["work_total"] == pytest.approx(2.0 * px * full[5])
    assert g["gbs"] == pytest.approx(2.0 * px * full[5] / 0.6e-3 / 1e9)
    assert classes["conv3x3"]["tflops"] == pytest.approx(2 * 2.0 * px * 64 * 64 * 9 / 0.6e-3 / 1e12)
    kinds = {k[0]: v["kind"] for k, v in groups}
    assert kinds["gemm"] == "flops" and kinds["groupnorm_apply"] == "bytes"
    assert bench.op_work("attention", (16, 65536, 64, 8)) == ("flops", 4.0 * 16 * 65536.0 * 65536 * 64)
