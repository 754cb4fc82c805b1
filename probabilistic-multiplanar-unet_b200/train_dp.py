"""Data-parallel training step: one process per GPU, gradients summed over NCCL (NVLink / NVSwitch).

The reference trains on one device (train.py); BASELINE config 4 asks for the batch-64 step split over
8 GPUs.  The loss is  sum_b CE_b + beta * mean_b KL_b  (probabilistic_unet.py:294-308): the CE part
ADDS over shards, the KL part AVERAGES.  Each rank therefore steps its shard with
``net.kl_world_size = world`` (the local KL term is weighted beta / world) and the gradients are
SUM-reduced — the result is exactly the gradient of the global objective evaluated with per-rank
BatchNorm statistics (standard data parallelism; the reference has no SyncBN either).

The exchange: 275 MB of fp32 gradients for the trainer model.  In the CUDA-graph step (graph=True, the fast path) the
gradients of a backward live in ONE flat buffer laid out in completion order and each completed range is all-reduced
inside the graph while the rest of the backward runs (train_engine.GraphedTrainStep); the eager step reduces after
backward in a few large flattened buckets.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 128 << 20) -> int:
    """SUM-all-reduce ``p.grad`` of every parameter that has one, in flat buckets.  Every rank must hold
    gradients for the same parameters (true for the ELBO step: the set is fixed by the architecture).
    Returns the number of collectives issued."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    buckets: List[List[torch.Tensor]] = [[]]
    size = 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if buckets[-1] and (size + nbytes > bucket_bytes or g.dtype != buckets[-1][0].dtype):
            buckets.append([])
            size = 0
        buckets[-1].append(g)
        size += nbytes
    for b in buckets:
        flat = torch._utils._flatten_dense_tensors(b)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        for g, synced in zip(b, torch._utils._unflatten_dense_tensors(flat, b)):
            g.copy_(synced)
    return len(buckets)


def dp_train_step(trainer, imgs, masks, optimizer, acc_steps: int = 1, clip_value: Optional[float] = 0.1,
                  group=None, step_now: bool = True, graph: bool = False):
    """train.py:85-110 for one shard: predict -> loss / acc_steps -> backward -> (all-reduce, clip, SGD step).
    Returns the local loss (sum over ranks of the local losses == the global loss).
    graph=True: forward + elbo + backward replayed as ONE CUDA graph (train_engine.GraphedTrainStep, captured on the first
    call for these input shapes; the eager step is bound by the host's ~3000 launches).  The look-at sample of
    trainer.predict() (its return value never enters the loss, probunet_trainer.py:27-39) is skipped."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if graph:
        from . import train_engine
        key = (tuple(imgs.shape), tuple(masks.shape), world, acc_steps)
        gs = getattr(trainer, "_graph_step", None)
        if gs is None or gs[0] != key:
            trainer.net.kl_world_size = world
            try:
                # one micro-step per optimizer step: the gradient exchange rides inside the graph, overlapped with the backward
                gs = (key, train_engine.GraphedTrainStep(trainer.net, imgs, masks, loss_scale=1.0 / acc_steps,
                                                         allreduce_group=group, allreduce=world > 1 and acc_steps == 1))
            finally:
                trainer.net.kl_world_size = 1
            trainer._graph_step = gs
        trainer.net.kl_world_size = world
        try:
            loss = gs[1].step(imgs, masks, accumulate=acc_steps > 1)
        finally:
            trainer.net.kl_world_size = 1
        if step_now:
            if not gs[1].ar_in_graph:
                allreduce_gradients(trainer.net.parameters(), group)
            if clip_value is not None:
                torch.nn.utils.clip_grad_value_(trainer.net.parameters(), clip_value)
            optimizer.step()
            optimizer.zero_grad()
        return loss.detach()
    # scoped to this step: the no-grad elbo() used for validation keeps the single-process weighting (beta * mean KL),
    # so logged training and validation losses stay comparable
    trainer.net.kl_world_size = world
    try:
        trainer.predict(imgs, masks)
        loss = trainer.loss(imgs, masks, None) / acc_steps
        loss.backward()
    finally:
        trainer.net.kl_world_size = 1
    if step_now:
        allreduce_gradients(trainer.net.parameters(), group)
        if clip_value is not None:
            torch.nn.utils.clip_grad_value_(trainer.net.parameters(), clip_value)
        optimizer.step()
        optimizer.zero_grad()
    return loss.detach()


def sync_batchnorm_buffers(net: torch.nn.Module, group=None) -> int:
    """Average BatchNorm running_mean / running_var over the ranks (num_batches_tracked: max) — call before a checkpoint
    or an evaluation: every rank has trained on its own shard, so the running statistics differ between ranks although
    the parameters are bit-identical, and the folded eval-mode inference path is built from exactly these buffers (what
    DDP's broadcast_buffers does for its rank-0 copy).  Returns the number of buffers reduced."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    n = 0
    for name, buf in net.named_buffers():
        if name.endswith("num_batches_tracked"):
            dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=group)
        elif buf.is_floating_point():
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            buf.div_(world)
        else:
            continue
        n += 1
    return n
