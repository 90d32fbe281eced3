"""Batch-sharded data parallelism (one process per GPU, NCCL over NVLink; gloo on CPU for tests).

The reference has no distributed code; the step shards naturally over the batch
(every sample is independent through encoders, sampling, decoders and lw[:, b]) and has exactly
one exchange: the sum of the parameter gradients.  Each stack's backward produces ONE flat
gradient buffer; as soon as it is final it is all-reduced asynchronously, so the decoders'
buckets travel while the encoders' backward is still running.  `m_iwae` is a SUM over the batch,
so a SUM all-reduce reproduces the single-process gradient of the concatenated batch; for
mean-type objectives (`elbo`) the optimiser divides by the world size (``grad_average=True``).
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist

_enabled = False
_pending: List = []


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun contract)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    enable(world > 1)
    return rank, world, local


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = bool(flag) and dist.is_initialized() and dist.get_world_size() > 1


def enabled() -> bool:
    return _enabled


def world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


_seen_this_pass: set = set()
_callback_queued = False


def _end_of_backward() -> None:
    """Runs when the autograd engine has finished the backward pass: every bucket is reduced before ANY optimiser
    (torch.optim.AdamW as in the reference scripts, or the fused one) can read a gradient."""
    global _callback_queued
    _callback_queued = False
    _seen_this_pass.clear()
    wait_all()


def first_bucket_of(params) -> bool:
    """True the first time a stack (identified by its first parameter) reports a bucket in the current backward pass."""
    if not params:
        return True
    key = params[0].data_ptr()
    if key in _seen_this_pass:
        return False
    _seen_this_pass.add(key)
    return True


def bucket_ready(flat: torch.Tensor, blocking: bool = False) -> None:
    """Called by a stack's backward when its flat gradient bucket is final.  Asynchronous by default (the bucket travels while
    the remaining backward runs); `blocking` reduces it — and everything still in flight — before returning, for buckets whose
    views autograd is going to ADD to existing gradients instead of adopting them."""
    global _callback_queued
    if not _enabled:
        return
    if not _callback_queued:
        try:
            torch.autograd.Variable._execution_engine.queue_callback(_end_of_backward)
            _callback_queued = True
        except RuntimeError:          # not inside a backward pass (a direct call): the caller waits explicitly
            pass
    if blocking:
        wait_all()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    else:
        _pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))


def wait_all() -> None:
    """Idempotent: called at the end of every backward pass and again by the fused optimiser before it reads gradients."""
    while _pending:
        _pending.pop().wait()


def shard(x, rank: int, world: int, multimodal: bool):
    """Contiguous split of a global batch along dim 0 (equal shards)."""
    def cut(t):
        n = t.shape[0] // world
        return t[rank * n:(rank + 1) * n]
    if multimodal:
        return [tuple(cut(t) for t in mod) for mod in x]
    return tuple(cut(t) for t in x)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    if world_size() > 1:
        for p in module.parameters():
            dist.broadcast(p.data, src)


def all_reduce_scalar(t: torch.Tensor, average: bool = False) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if average:
            t = t / world_size()
    return t


class _GatherBatch(torch.autograd.Function):
    """cat over ranks along dim 0, differentiable on any backend: the backward sums the gradient of the gathered tensor
    over the ranks and keeps this rank's rows."""

    @staticmethod
    def forward(ctx, t):
        parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, t.contiguous())
        ctx.n = t.shape[0]
        return torch.cat(parts, 0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        r = dist.get_rank()
        return g[r * ctx.n:(r + 1) * ctx.n]


def gather_batch(t: torch.Tensor) -> torch.Tensor:
    """All ranks' rows of `t` (equal shard sizes), with gradients flowing back to their owners; identity without DP."""
    return _GatherBatch.apply(t) if _enabled else t


def rank() -> int:
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
