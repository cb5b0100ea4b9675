"""Host-to-host retrieval with the gallery streamed through the GPU in chunks.

The gallery of config 5 is 1.07 GB of fp16 rows; copying it to the device before scoring would
add ~20 ms to a 46 ms step.  Here the host->device copy of chunk i+1 (copy stream, pinned
memory) overlaps K0 + K1 + K2 on chunk i (compute stream); every chunk yields a per-query
top-k list with global indices and the lists are merged at the end with K2's ordering rule —
the same plumbing the multi-GPU path uses across ranks, applied across chunks.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops
from .ops import JegalError


def chunk_schedule(n_clips: int, chunk_clips: int, ramp: bool = True):
    """Clip counts of the successive gallery chunks.  Scoring cannot start before the first chunk has arrived
    and cannot finish before the last chunk's copy + its scoring, so with ``ramp`` the chunks grow geometrically
    at the start (chunk/8, /8, /4, /2), stay at full size in the middle and shrink again at the end
    (/2, /4, /8, /8): the start-up latency is the copy of 1/8 chunk, and when the copies are the bottleneck
    (several GPUs pulling from one host) the tail after the last copy is the scoring of 1/8 chunk."""
    n, c = int(n_clips), max(1, int(chunk_clips))
    if n <= 0:
        return []
    if not ramp or c < 64:
        return [c] * (n // c) + ([n % c] if n % c else [])
    while n < 2 * c and c >= 128:
        c //= 2
    if n < 2 * c:
        return [n]
    up = [max(1, c // 8), max(1, c // 8), max(1, c // 4)]
    up.append(c - sum(up))  # the ramp sums to exactly one full chunk whatever c is (c // 8 * 8 != c in general)
    down = up[::-1]
    mid = n - sum(up) - sum(down)
    sched = up + [c] * (mid // c) + ([mid % c] if mid % c else []) + down
    assert sum(sched) == n and min(sched) > 0, (n, chunk_clips, sched)
    return sched


class StreamedGallery:
    """Pinned host gallery (packed fp16/fp32 rows + per-clip lengths) cut into clip-aligned chunks."""

    def __init__(self, rows_host: torch.Tensor, lengths: np.ndarray, chunk_clips: int = 8192, device=None,
                 idx_base: int = 0, ramp: bool = True):
        if rows_host.is_cuda or rows_host.dim() != 2 or rows_host.shape[1] != 512:
            raise JegalError("StreamedGallery: rows_host must be a host [rows, 512] tensor")
        self.rows = rows_host if rows_host.is_pinned() else rows_host.pin_memory()
        self.lengths = np.asarray(lengths, dtype=np.int64)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.idx_base = idx_base
        cu = np.concatenate([[0], np.cumsum(self.lengths)])
        self.chunks = []
        lo = 0
        sched = chunk_schedule(len(self.lengths), chunk_clips, ramp)
        if sum(sched) != len(self.lengths):
            raise JegalError(f"StreamedGallery: chunk schedule covers {sum(sched)} of {len(self.lengths)} clips")
        for size in sched:
            hi = lo + size
            self.chunks.append((lo, hi, int(cu[lo]), int(cu[hi]), ops.Layout.from_lengths(self.lengths[lo:hi])))
            lo = hi
        max_rows = max((c[3] - c[2] for c in self.chunks), default=0)
        self.stage = [torch.empty((max_rows, 512), dtype=self.rows.dtype, device=self.dev) for _ in range(2)]
        self.op16 = [torch.empty((max_rows, 512), dtype=torch.bfloat16, device=self.dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.n_clips = len(self.lengths)

    @property
    def nbytes(self) -> int:
        return self.rows.numel() * self.rows.element_size()


def _query_parts(q_layout: ops.Layout, n_parts: int):
    """Contiguous clip ranges of the query set: (clip_lo, clip_hi, row_lo, row_hi, layout) per part."""
    lengths = np.asarray(q_layout.lengths, dtype=np.int64)
    nq = len(lengths)
    n_parts = max(1, min(int(n_parts), nq)) if nq else 1
    cache = q_layout.__dict__.setdefault("_query_parts_cache", {})  # layouts own device tables: build once
    if n_parts in cache:
        return cache[n_parts]
    cu = np.concatenate([[0], np.cumsum(lengths)])
    per = (nq + n_parts - 1) // n_parts if nq else 0
    parts = []
    for lo in range(0, nq, max(per, 1)):
        hi = min(lo + per, nq)
        parts.append((lo, hi, int(cu[lo]), int(cu[hi]), q_layout if (lo == 0 and hi == nq) else ops.Layout.from_lengths(lengths[lo:hi])))
    cache[n_parts] = parts
    return parts


def retrieve_topk_streamed(q_host: torch.Tensor, q_layout: ops.Layout, gallery: StreamedGallery, k: int = 10,
                           mode: str = "max_t_mean_w", queries_are: str = "gesture",
                           q_dev: Optional[torch.Tensor] = None, q_parts: int = 1,
                           bcast_src: Optional[int] = None, group=None,
                           q_dtype: torch.dtype = torch.float16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery clips per query; queries and gallery start in (pinned) host memory.
    Returns DEVICE tensors (values [Q, k], global indices [Q, k]); call .cpu() to finish the round trip.
    ``q_dev`` may carry queries that are already on the device (e.g. after a broadcast).

    ``q_parts`` > 1 pipelines the QUERY side too: the query clips are cut into contiguous parts, part p+1
    is copied (and, with ``bcast_src`` set in a torch.distributed job, broadcast from that rank over NCCL)
    while part p is being scored against the first gallery chunk, so scoring starts after 1/q_parts of the
    query transfer instead of all of it.  On every rank but ``bcast_src`` ``q_host`` may be None (its dtype
    is then ``q_dtype``)."""
    dev = gallery.dev
    main = torch.cuda.current_stream(dev)
    nq = q_layout.n_clips
    import torch.distributed as dist

    bcast = bcast_src is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    parts = _query_parts(q_layout, q_parts if q_dev is None else 1)
    ready = [None] * len(parts)  # per part: a CUDA event (copy) or an NCCL work handle (broadcast)
    if q_dev is None:
        i_am_src = (not bcast) or dist.get_rank(group) == bcast_src
        q_dtype = q_host.dtype if q_host is not None else q_dtype
        q_dev = torch.empty((q_layout.rows, 512), dtype=q_dtype, device=dev)
        if not hasattr(gallery, "q_stream"):
            gallery.q_stream = torch.cuda.Stream(device=dev)
        gallery.q_stream.wait_stream(main)  # q_dev was allocated on main
        with torch.cuda.stream(gallery.q_stream):
            for p, (lo, hi, r0, r1, _) in enumerate(parts):
                if i_am_src:
                    q_dev[r0:r1].copy_(q_host[r0:r1], non_blocking=True)
                if bcast:  # enqueued behind the copy; the next part's copy does not wait for it
                    ready[p] = dist.broadcast(q_dev[r0:r1], src=bcast_src, group=group, async_op=True)
                else:
                    ready[p] = torch.cuda.Event()
                    ready[p].record(gallery.q_stream)
        q_dev.record_stream(gallery.q_stream)
    if not gallery.chunks:  # an empty shard still took part in the broadcast above (it is a collective)
        for r in ready:
            if r is not None and not isinstance(r, torch.cuda.Event):
                r.wait()
        return (torch.full((nq, k), float("-inf"), device=dev), torch.full((nq, k), -1, dtype=torch.int32, device=dev))
    q16 = torch.empty((q_layout.rows, 512), dtype=torch.bfloat16, device=dev)
    vals = torch.empty((len(gallery.chunks), nq, k), dtype=torch.float32, device=dev)
    idxs = torch.empty((len(gallery.chunks), nq, k), dtype=torch.int32, device=dev)

    def issue_copy(c: int):
        lo, hi, r0, r1, _ = gallery.chunks[c]
        b = c & 1
        with torch.cuda.stream(gallery.copy_stream):
            if c >= 2:
                gallery.copy_stream.wait_event(gallery.consumed[b])  # K0 of chunk c-2 has read the buffer
            gallery.stage[b][: r1 - r0].copy_(gallery.rows[r0:r1], non_blocking=True)
            gallery.copied[b].record(gallery.copy_stream)

    # the staging buffers may still be read by the previous call's K0 on `main` (the function returns device
    # tensors without synchronising): order this call's first copies behind everything enqueued so far
    gallery.copy_stream.wait_stream(main)
    issue_copy(0)
    for c, (lo, hi, r0, r1, lay) in enumerate(gallery.chunks):
        b = c & 1
        if c + 1 < len(gallery.chunks):
            issue_copy(c + 1)
        main.wait_event(gallery.copied[b])
        g16 = gallery.op16[b][: r1 - r0]
        ops.prep(gallery.stage[b][: r1 - r0], lay, out=g16)
        gallery.consumed[b].record(main)
        # the query parts only matter while they are still arriving: the first gallery chunk is scored part
        # by part, every later chunk against the whole query set in one launch
        todo = parts if c == 0 else [(0, nq, 0, q_layout.rows, q_layout)]
        for p, (qlo, qhi, qr0, qr1, qlay) in enumerate(todo):
            if c == 0:  # first use of this query part: wait for its transfer, normalise + cast it
                if ready[p] is not None:
                    if isinstance(ready[p], torch.cuda.Event):
                        main.wait_event(ready[p])
                    else:
                        ready[p].wait()
                ops.prep(q_dev[qr0:qr1], qlay, out=q16[qr0:qr1])
            if queries_are == "gesture":
                s = ops.simpool_allpairs(q16[qr0:qr1], qlay, g16, lay, mode)
            else:
                s = ops.simpool_allpairs(g16, lay, q16[qr0:qr1], qlay, mode, content_major=True)
            v, i = ops.topk(s, k, idx_offset=gallery.idx_base + lo)
            vals[c, qlo:qhi].copy_(v)
            idxs[c, qlo:qhi].copy_(i)
    if len(gallery.chunks) == 1:
        return vals[0], idxs[0]
    return ops.topk_merge(vals, idxs)
