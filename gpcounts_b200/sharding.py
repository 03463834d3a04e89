"""Gene sharding over the GPUs of one box.

Genes are independent, so the fit path has no collective: rank r of G fits genes r, r+G, r+2G, ... (an
interleaved partition, which balances cost because neighbouring genes are not cost-correlated) and only the
final per-gene statistics rows (D x K x 16 doubles, ~2.5 MB for 20k genes) are all-gathered - over NCCL /
NVLink when the tensors are on the GPU, over gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n, rank, world):
    """Positions of the genes owned by `rank` (interleaved partition)."""
    return np.arange(rank, n, world, dtype=np.int64)


def gather_rows(local, n_total, rank, world):
    """All-gather the per-gene rows of every rank and restore the original gene order.

    local: (n_local, ...) tensor holding the rows of positions ``shard_indices(n_total, rank, world)``."""
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    pad_shape = (per,) + tuple(local.shape[1:])
    buf = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_total, r, world)
        full[torch.as_tensor(idx, device=local.device)] = out[r][: len(idx)]
    return full
