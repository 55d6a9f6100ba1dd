"""Multi-GPU plumbing: one process per GPU.

Sampling: images sharded across ranks, replicated weights, NO data-path collective (images are independent; SURVEY.md
section 8(e)); torch.distributed is only used for the barrier and the max-over-ranks timing (bench.py).
Training: data parallel, ONE collective -- the gradient all-reduce over NCCL / NVLink -- issued in buckets while the
backward is still running (`GradBuckets`, used by training.Trainer).  Everything here is device-agnostic, so the CPU tests
drive the same code over gloo with world_size 2."""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for this rank (the first n_items % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def codec_threads_per_rank(host_cores, world):
    return max(1, host_cores // max(1, world))


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_mean_(flat):
    """In-place mean over ranks of one flat tensor (the training step's single gradient collective; NCCL over
    NVLink on the GPUs, gloo in the CPU tests).  Identity when torch.distributed is not initialised."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.mul_(1.0 / dist.get_world_size())
    return flat


def init_process_group_from_env(device=None):
    """(rank, world, local) from the torchrun environment; initialises torch.distributed (NCCL when a CUDA device is given,
    gloo otherwise) if WORLD_SIZE > 1.  MASTER_ADDR / MASTER_PORT come from the launcher."""
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if device is not None and getattr(device, "type", "cpu") == "cuda":
            dist.init_process_group("nccl", device_id=device)
        else:
            dist.init_process_group("gloo")
    return rank, world, local


def _world():
    import torch.distributed as dist
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


class GradBuckets:
    """Bucketed, overlapped all-reduce of ONE flat gradient buffer whose tail becomes final first.

    The backward finishes the blocks in reverse registration order, so after each block everything at offsets >= lo is
    final: `final_from(lo)` sends [lo, previous lo) as soon as it holds >= bucket_bytes (asynchronously: NCCL runs it on its
    own stream, ordered after the kernels enqueued so far, so it overlaps the rest of the backward); `finish()` sends the
    head, waits for everything and scales by 1 / world.  `reset()` starts a new step and REFUSES to drop handles that are
    still pending (a forward_backward called twice without finish() would otherwise leave the buffer partly summed)."""

    def __init__(self, flat, bucket_bytes=64 << 20):
        self.flat, self.bucket_bytes = flat, bucket_bytes
        self.pending, self.hi = [], flat.numel()
        self.collectives = 0

    def reset(self):
        if self.pending:
            raise RuntimeError("GradBuckets.reset(): all-reduces of the previous step are still pending -- call finish() "
                               "(Trainer.allreduce_grads) after every forward_backward that ran with overlap=True")
        self.hi = self.flat.numel()

    def final_from(self, lo):
        import torch.distributed as dist
        if _world() == 1 or lo >= self.hi:
            return
        if (self.hi - lo) * self.flat.element_size() >= self.bucket_bytes:
            self.pending.append(dist.all_reduce(self.flat[lo:self.hi], op=dist.ReduceOp.SUM, async_op=True))
            self.collectives += 1
            self.hi = lo

    def finish(self):
        import torch.distributed as dist
        world = _world()
        if world == 1:
            self.hi = 0
            return
        if self.hi > 0:
            self.pending.append(dist.all_reduce(self.flat[:self.hi], op=dist.ReduceOp.SUM, async_op=True))
            self.collectives += 1
            self.hi = 0
        for h in self.pending:
            h.wait()
        self.pending = []
        self.flat.mul_(1.0 / world)


class CosineWarmRestarts:
    """lr schedule of the reference's training scripts: CosineAnnealingWarmRestarts(T_0=100, T_mult=2), stepped once per
    epoch (webp_training.py:530, 776; avif.py:797).  Closed form of torch's scheduler: within a cycle of length T_i that has
    run T_cur epochs, lr = eta_min + (base_lr - eta_min) * (1 + cos(pi * T_cur / T_i)) / 2."""

    def __init__(self, base_lr, T_0=100, T_mult=2, eta_min=0.0):
        if T_0 <= 0 or T_mult < 1:
            raise ValueError("T_0 must be positive and T_mult >= 1")
        self.base_lr, self.T_0, self.T_mult, self.eta_min = base_lr, T_0, T_mult, eta_min
        self.T_cur, self.T_i, self.epoch = 0, T_0, 0

    def lr(self):
        import math
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * self.T_cur / self.T_i)) / 2

    def step(self):
        self.epoch += 1
        self.T_cur += 1
        if self.T_cur >= self.T_i:
            self.T_cur -= self.T_i
            self.T_i *= self.T_mult
        return self.lr()

    def state_dict(self):
        return dict(T_cur=self.T_cur, T_i=self.T_i, epoch=self.epoch, base_lr=self.base_lr, T_0=self.T_0,
                    T_mult=self.T_mult, eta_min=self.eta_min)

    def load_state_dict(self, d):
        for k, v in d.items():
            setattr(self, k, v)
