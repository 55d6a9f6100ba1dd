"""Data-parallel training step on two GPUs over NCCL (BASELINE configs[3]): the gradients that come out of the
bucketed all-reduces launched under the backward must equal (a) one all-reduce after the backward and (b) a single-rank run
over the concatenated batch (every term of frequency_aware_loss is a mean over the batch and GroupNorm is per sample, so the
average of the two half-batch gradients IS the full-batch gradient).  Needs two visible GPUs: skipped on the one-GPU box,
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _rank_main(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
    import ddpm_image_restoration_b200 as P
    from ddpm_image_restoration_b200.training import Trainer
    from oracle import restated as R
    from oracle import weights as W
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    sd = W.make_state_dict("webp", 0)
    x0 = W.synthetic_images(4, 32, 32, seed=77)
    xt = R.codec_roundtrip(x0, 10, "webp")
    t = torch.tensor([37.0, 81.0, 12.0, 55.0]) / 100.0
    lo, hi = rank * 2, rank * 2 + 2

    def trainer(**kw):
        m = P.WebPDiffusionModel()
        m.load_state_dict(sd)
        return Trainer(m.to(dev).set_precision("fp32"), dropout=0.0, **kw)

    out = {}
    # (1) buckets under the backward; 8 MB buckets so that the 458 MB buffer goes out in many collectives
    tr = trainer(bucket_bytes=8 << 20)
    tr.forward_backward(xt[lo:hi].to(dev), t[lo:hi].to(dev), x0[lo:hi].to(dev), overlap=True)
    pending = len(tr.buckets.pending)
    tr.allreduce_grads()
    g_bucketed = tr.flat_grad.clone()
    out["collectives"] = tr.buckets.collectives
    out["pending_before_finish"] = pending
    # a second overlapped backward without finish() must be refused, not silently mixed into the sums
    tr.forward_backward(xt[lo:hi].to(dev), t[lo:hi].to(dev), x0[lo:hi].to(dev), overlap=True)
    try:
        tr.forward_backward(xt[lo:hi].to(dev), t[lo:hi].to(dev), x0[lo:hi].to(dev), overlap=True)
        out["refused"] = False
    except RuntimeError:
        out["refused"] = True
    tr.allreduce_grads()
    del tr
    # (2) one all-reduce after the backward
    tr = trainer(overlap_allreduce=False)
    tr.forward_backward(xt[lo:hi].to(dev), t[lo:hi].to(dev), x0[lo:hi].to(dev), overlap=True)
    tr.allreduce_grads()
    g_single = tr.flat_grad.clone()
    out["collectives_single"] = tr.buckets.collectives
    # (3) standalone forward_backward leaves the gradients local: full batch on every rank, no collective
    tr.forward_backward(xt.to(dev), t.to(dev), x0.to(dev))
    g_full = tr.flat_grad.clone()
    out["collectives_after_local"] = tr.buckets.collectives
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    out["bucketed_vs_single"] = rel(g_bucketed, g_single)
    out["bucketed_vs_full_batch"] = rel(g_bucketed, g_full)
    # both ranks must hold the same averaged gradient bit for bit (NCCL all-reduce gives every rank the same sums)
    chk = g_bucketed.double().abs().sum().view(1)
    hi_, lo_ = chk.clone(), chk.clone()
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX); dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    out["rank_mismatch"] = float(hi_ - lo_)
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_nccl_bucketed_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in (0, 1):
        o = res[rank]
        print(rank, o)
        assert o["collectives"] >= 8 and o["pending_before_finish"] >= 7      # most buckets were in flight under the backward
        assert o["collectives_single"] == 1 and o["collectives_after_local"] == 1
        assert o["refused"]
        assert o["bucketed_vs_single"] < 1e-5                                   # same sums, only the partition differs
        assert o["bucketed_vs_full_batch"] < 2e-4                               # fp32 reassociation over the batch
        assert o["rank_mismatch"] == 0.0
