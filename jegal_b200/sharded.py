"""Gallery-sharded retrieval over the GPUs of one box.

Pair scores are independent, so the gallery is partitioned BY CLIP into
contiguous ranges (one per rank), the query clips are replicated (one broadcast),
every rank runs K1 + K2 on its shard, and the only exchange is the per-shard
top-k list: k (score, global index) pairs per query.  The merge uses the same
ordering rule as K2 (score descending, ties to the lower GLOBAL index), so the
N-GPU result equals the 1-GPU result bit for bit on the indices.

The reference has no counterpart (no torch.distributed anywhere in it).

Two exchange paths:
  * ``retrieve_topk_sharded(..)``                 K2, one all-gather of [Q, k] values + one of
                                                  indices (NCCL), then the K2 merge kernel;
  * ``retrieve_topk_sharded(.., exchange=ex)``    with ``ex = ops.TopkExchange(Q, k)``: K2 stores each
                                                  rank's list straight into every peer's exchange block
                                                  over NVLink (IPC-mapped pointers), publishes a flag,
                                                  and the merge kernel waits on the flags — no
                                                  host-launched collective on the data path
                                                  (csrc/exchange.cu).

The local scoring / merge functions are injectable so the host logic (sharding,
offsets, collectives) is covered on CPU with the gloo backend in tests/.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous clip range [lo, hi) of `rank`: ceil-sized shards, last ones may be short/empty."""
    per = (n_clips + world - 1) // world
    lo = min(rank * per, n_clips)
    hi = min(lo + per, n_clips)
    return lo, hi


def shard_rows(cu_len: np.ndarray, lo: int, hi: int) -> Tuple[int, int]:
    """Row range of clips [lo, hi) in the packed gallery matrix."""
    return int(cu_len[lo]), int(cu_len[hi])


def _cuda_local_topk(q16, q_layout, g16, g_layout, k, mode, idx_offset, queries_are):
    from . import ops

    if g_layout.n_clips == 0:
        dev = q16.device
        return (torch.full((q_layout.n_clips, k), float("-inf"), device=dev),
                torch.full((q_layout.n_clips, k), -1, dtype=torch.int32, device=dev))
    if queries_are == "gesture":
        s = ops.simpool_allpairs(q16, q_layout, g16, g_layout, mode)
    else:
        s = ops.simpool_allpairs(g16, g_layout, q16, q_layout, mode, content_major=True)
    return ops.topk(s, k, idx_offset=idx_offset)


def _cuda_merge(vals, idxs):
    from . import ops

    return ops.topk_merge(vals.contiguous(), idxs.contiguous())


def retrieve_topk_sharded(
    q16: torch.Tensor,
    q_layout,
    shard16: torch.Tensor,
    shard_layout,
    shard_lo: int,
    k: int = 10,
    mode: str = "max_t_mean_w",
    queries_are: str = "gesture",
    group: Optional[dist.ProcessGroup] = None,
    local_topk_fn: Callable = _cuda_local_topk,
    merge_fn: Callable = _cuda_merge,
    exchange=None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Every rank returns the same merged (values [Q, k], global indices [Q, k]).

    ``q16`` are the replicated, prepped query rows; ``shard16`` / ``shard_layout`` this rank's
    gallery clips, whose first clip has global index ``shard_lo``.
    """
    if exchange is not None:  # fused K2 + NVLink exchange + merge
        from . import ops

        if shard_layout.n_clips == 0:
            s = torch.empty((q_layout.n_clips, 0), dtype=torch.float32, device=q16.device)
        elif queries_are == "gesture":
            s = ops.simpool_allpairs(q16, q_layout, shard16, shard_layout, mode)
        else:
            s = ops.simpool_allpairs(shard16, shard_layout, q16, q_layout, mode, content_major=True)
        return exchange.topk(s, idx_offset=shard_lo)
    v, i = local_topk_fn(q16, q_layout, shard16, shard_layout, k, mode, shard_lo, queries_are)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return v, i
    nq = v.shape[0]
    vals = torch.empty((world * nq, k), dtype=v.dtype, device=v.device)  # rank-major concatenation
    idxs = torch.empty((world * nq, k), dtype=i.dtype, device=i.device)
    dist.all_gather_into_tensor(vals, v.contiguous(), group=group)
    dist.all_gather_into_tensor(idxs, i.contiguous(), group=group)
    return merge_fn(vals.view(world, nq, k), idxs.view(world, nq, k))


def broadcast_queries(q_rows: Optional[torch.Tensor], n_rows: int, dtype: torch.dtype, device,
                      src: int = 0, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Replicate the packed query rows from `src` to every rank (65.5 MB for config 5)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return q_rows
    if q_rows is None:
        q_rows = torch.empty((n_rows, 512), dtype=dtype, device=device)
    dist.broadcast(q_rows, src=src, group=group)
    return q_rows
