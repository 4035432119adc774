"""Batch sharding of studies over the GPUs of one box (one process per GPU).

Studies are independent (BatchNorm in eval mode uses running stats, LayerNorm/softmax are per row), so the
path shards over the batch with NO data-path collective; the only exchange is gathering the [B/N, 13] logits
(plus probabilities / label vectors) on every rank - 52 B per study, latency-bound.  `torch.distributed` is
used for plumbing only: NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `n` studies owned by `rank`; the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gathers per-rank row blocks (possibly of different heights) into the full [n_total, ...] tensor,
    in rank order - the inverse of `shard_range`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    base, extra = divmod(n_total, world)
    hmax = base + (1 if extra else 0)
    pad = torch.zeros((hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * hmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        pieces.append(out[r * hmax: r * hmax + (hi - lo)])
    return torch.cat(pieces, 0)


def inference_batch_sharded(run_local, images, tokens: dict, group=None):
    """Splits a global batch over the ranks, runs `run_local(images_shard, tokens_shard) -> dict of row tensors`
    on this rank's shard and gathers every output on all ranks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(images)
    lo, hi = shard_range(n, world, rank)
    out = run_local(images[lo:hi], {k: v[lo:hi] for k, v in tokens.items()})
    return {k: gather_rows(v, n, group) for k, v in out.items()}
