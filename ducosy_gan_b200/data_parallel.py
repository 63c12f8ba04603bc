"""Batch-sharded (data-parallel) CycleGAN step: one process per GPU, each with full replicas of G_A2B / G_B2A / D_A / D_B
and 1/world of the batch (BASELINE config 4).  Replaces the reference's ``nn.DataParallel`` wrapping
(modules/trainer.py:333-338: per-forward parameter broadcast + gather to GPU 0) with the one exchange the step
really has: a gradient all-reduce (mean) over NCCL/NVLink per optimiser, on a flat bucket the parameter ``.grad``
tensors are views of (no flatten / unflatten copies).

InstanceNorm is per sample, so activations need no synchronisation.  Two loss terms are *batch-global* in the
reference -- ContrastRegionLoss and ContrastEdgeLoss take an unbiased std (and a top-10 % mean) over the whole batch
tensor (trainer.py:117-127,163-181) -- so their three inputs are all-gathered (1 MB per sample each) and the terms are
evaluated on the full batch on every rank; with the mean-all-reduce that follows, weighting them by ``world`` gives
exactly the gradient of the single-process step.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .losses import l1_loss, mse_gan_loss
from .trainer import CycleGANStep


class GradBucket:
    """Flat fp32 gradient buffer; every ``p.grad`` is a view into it, ``all_reduce_mean`` is one collective."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.offsets, total = [], 0
        for p in self.params:                       # every view starts 16-byte aligned (vectorised Adam / reductions)
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self):
        self.flat.zero_()
        for p, off in zip(self.params, self.offsets):   # re-attach views an optimizer.zero_grad(set_to_none=True) may have dropped
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + off * 4:
                p.grad = self.flat[off:off + p.numel()].view_as(p)

    def all_reduce_mean(self, group=None, async_op=False):
        if not dist.is_initialized():
            return None
        world = dist.get_world_size(group)
        if world == 1:
            return None
        if dist.get_backend(group) == "nccl":
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=False)   # gloo has no AVG
        self.flat.div_(world)
        return work


class _AllGatherBatch(torch.autograd.Function):
    """[b, ...] per rank -> [world*b, ...] on every rank (rank-major).  Backward hands each rank the slice of the
    gradient that belongs to its own samples: every rank evaluates the same full-batch loss, so no reduction is due."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        x = x.contiguous()
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        ctx.b = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        r = dist.get_rank(ctx.group)
        return g[r * ctx.b:(r + 1) * ctx.b].contiguous(), None


def all_gather_batch(x, group=None):
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllGatherBatch.apply(x, group)


TERM_KEYS = ("GAN", "cycle", "id", "grad_cycle", "grad_id", "ssim", "contrast_attention", "contrast_region", "contrast_edge")


def loss_weights(lambda_cyc=10.0, lambda_id=5.0):
    """The loss mix of reference modules/trainer.py:493-512 (argmanager.py:97-98)."""
    return dict(GAN=1.0, cycle=lambda_cyc, id=lambda_id, grad_cycle=5.0, grad_id=2.5, ssim=2.0, contrast_attention=2.0,
                contrast_region=1.5, contrast_edge=1.0)


def logged_losses(terms, loss_D_A, loss_D_B, lambda_cyc=10.0, lambda_id=5.0, group=None):
    """What reference modules/trainer.py:527-531 logs, for the GLOBAL batch, from each rank's shard.

    The seven per-sample terms and the two discriminator losses are means over the samples of the rank, and the shards are
    equal, so their mean over ranks is the mean over the whole batch; the two batch-global terms (region, edge) were
    evaluated on the gathered batch and are identical on every rank, so the same mean leaves them alone.  ``G`` is then the
    reference's mix of those values -- each term counted once.  (The ``world``-weighted expression that
    ``generator_losses`` differentiates is a backward surrogate for the mean-all-reduce of the gradients, not a loss
    anybody should read.)  One small all-reduce of 11 floats; detached, graph-capturable."""
    vec = torch.stack([terms[k].detach().to(torch.float32).reshape(()) for k in TERM_KEYS]
                      + [loss_D_A.detach().to(torch.float32).reshape(()), loss_D_B.detach().to(torch.float32).reshape(())])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(vec, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
            vec = vec / dist.get_world_size(group)
    out = {k: vec[i] for i, k in enumerate(TERM_KEYS)}
    w = loss_weights(lambda_cyc, lambda_id)
    out["G"] = sum(w[k] * out[k] for k in TERM_KEYS)
    out["D_A"], out["D_B"] = vec[len(TERM_KEYS)], vec[len(TERM_KEYS) + 1]
    return out


def check_equal_shards(local_batch, group=None):
    """The mean-all-reduce of gradients and losses is the global-batch mean only for equal shards (and
    ``all_gather_into_tensor`` needs them too): refuse anything else instead of silently weighting samples unevenly."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    sizes = [None] * dist.get_world_size(group)
    dist.all_gather_object(sizes, int(local_batch), group=group)
    if len(set(sizes)) != 1:
        raise ValueError(f"data-parallel step needs equal shards on every rank, got per-rank batch sizes {sizes}")


def shard_batch(n, rank, world):
    """Contiguous sample range of this rank: [n*rank//world, n*(rank+1)//world)."""
    return n * rank // world, n * (rank + 1) // world


class DataParallelCycleGANStep(CycleGANStep):
    """``CycleGANStep`` over the local shard of the batch, with the gradient exchange between backward and optimiser step.
    Every rank must construct it with the same ``seed`` (identical replicas, as DataParallel's broadcast guarantees)."""

    def __init__(self, *args, group=None, **kw):
        if kw.get("seed") is None:
            raise ValueError("data-parallel replicas need a common seed")
        super().__init__(*args, **kw)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bucket_G = GradBucket(list(self.G_A2B.parameters()) + list(self.G_B2A.parameters()))
        self.bucket_D_A = GradBucket(list(self.D_A.parameters()))
        self.bucket_D_B = GradBucket(list(self.D_B.parameters()))
        self._checked_batch = None

    def _batch_global_losses(self, fake_B, real_B, real_A):
        """Region / edge terms on the all-gathered batch, identical on every rank.  Their weight in the differentiated
        expression is ``world`` -- a BACKWARD SURROGATE, not a logged loss: full-batch terms x world, then the mean-all-reduce of
        the gradients, gives exactly the single-process gradient (module docstring).  step() reports logged_losses()."""
        if self.world == 1:
            return super()._batch_global_losses(fake_B, real_B, real_A)
        g = self.group
        fake_B_all, real_B_all, real_A_all = all_gather_batch(fake_B, g), all_gather_batch(real_B, g), all_gather_batch(real_A, g)
        return (self.criterion_contrast_region(fake_B_all, real_B_all, real_A_all),
                self.criterion_contrast_edge(fake_B_all, real_B_all, real_A_all), float(self.world))

    def step(self, real_A, real_B, masks=None):
        """real_A / real_B / masks are this rank's shard of the batch (equal shards on every rank).  Returns the losses the
        reference logs (trainer.py:527-531) for the global batch -- identical on every rank, see ``logged_losses``."""
        if self.world > 1 and self._checked_batch != real_A.shape[0] and not torch.cuda.is_current_stream_capturing():
            check_equal_shards(real_A.shape[0], self.group)
            self._checked_batch = real_A.shape[0]
        # The three exchanges run on NCCL's own stream while the next phase computes: the discriminator losses depend on the
        # fakes (made before any update) and on the discriminators' own weights only, so moving optimizer_G.step() behind them
        # changes nothing (DUCOSY_DP_OVERLAP=0 restores the literal order of trainer.py:514-524).
        import os
        overlap = os.environ.get("DUCOSY_DP_OVERLAP", "1") != "0" and self.world > 1
        self.bucket_G.zero()
        loss_G, terms, fake_A, fake_B = self.generator_losses(real_A, real_B, masks)
        loss_G.backward()
        work_G = self.bucket_G.all_reduce_mean(self.group, async_op=overlap)
        if not overlap:
            self.optimizer_G.step()

        self.bucket_D_A.zero()   # also discards what loss_G.backward() left in the discriminators (trainer.py:517)
        loss_D_A = self._disc_loss(self.D_A, real_A, fake_A)
        loss_D_A.backward()
        work_A = self.bucket_D_A.all_reduce_mean(self.group, async_op=overlap)
        if not overlap:
            self.optimizer_D_A.step()

        self.bucket_D_B.zero()
        loss_D_B = self._disc_loss(self.D_B, real_B, fake_B)
        loss_D_B.backward()
        work_B = self.bucket_D_B.all_reduce_mean(self.group, async_op=overlap)
        if overlap:
            for work, opt in ((work_G, self.optimizer_G), (work_A, self.optimizer_D_A), (work_B, self.optimizer_D_B)):
                if work is not None:
                    work.wait()
                opt.step()
        else:
            self.optimizer_D_B.step()
        if self.world == 1:
            out = {k: v.detach() for k, v in terms.items()}
            out.update(G=loss_G.detach(), D_A=loss_D_A.detach(), D_B=loss_D_B.detach())
            return out
        return logged_losses(terms, loss_D_A, loss_D_B, self.lambda_cyc, self.lambda_id, self.group)


class GraphedCycleGANStep:
    """One whole optimisation step -- six generator passes, the discriminators, nine losses, three gradient all-reduces and
    three Adam updates, ~3000 kernel launches -- captured once in a CUDA graph and replayed.  At small per-rank batches
    (global batch 8 over 8 GPUs = one sample per rank) the eager step is bound by Python launching those kernels, not by the
    GPU; the replay is one launch.

    ``step`` must be a ``DataParallelCycleGANStep(..., capturable=True)`` (gradients live in static buckets, Adam reads
    {lr, step} from device memory).  The example batch fixes the shapes; the ``warmup`` eager steps needed before a capture
    are real optimisation steps on that batch."""

    def __init__(self, step: "DataParallelCycleGANStep", real_A, real_B, masks=None, warmup=3):
        if not all(o.capturable for o in (step.optimizer_G, step.optimizer_D_A, step.optimizer_D_B)):
            raise ValueError("construct the step with capturable=True")
        self.step = step
        self.static = [None if t is None else t.clone() for t in (real_A, real_B, masks)]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step.step(*self.static)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = step.step(*self.static)
        self._params = [p for m in (step.G_A2B, step.G_B2A, step.D_A, step.D_B) for p in m.parameters()]

    def close(self):
        """Release the graph.  Call this before ``torch.distributed.destroy_process_group()``: tearing NCCL down while a
        live graph still holds captured collectives blocks forever."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None
            self.out = None

    def __call__(self, real_A, real_B, masks=None):
        for dst, src in zip(self.static, (real_A, real_B, masks)):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        for o in (self.step.optimizer_G, self.step.optimizer_D_A, self.step.optimizer_D_B):
            o.push_lr()
        self.graph.replay()
        for p in self._params:      # the replay changed the weights behind autograd's back: keep _version-keyed caches honest
            torch.autograd.graph.increment_version(p)
        return self.out
