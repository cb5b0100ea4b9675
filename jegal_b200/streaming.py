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


class StreamedGallery:
    """Pinned host gallery (packed fp16/fp32 rows + per-clip lengths) cut into clip-aligned chunks."""

    def __init__(self, rows_host: torch.Tensor, lengths: np.ndarray, chunk_clips: int = 8192, device=None,
                 idx_base: int = 0):
        if rows_host.is_cuda or rows_host.dim() != 2 or rows_host.shape[1] != 512:
            raise JegalError("StreamedGallery: rows_host must be a host [rows, 512] tensor")
        self.rows = rows_host if rows_host.is_pinned() else rows_host.pin_memory()
        self.lengths = np.asarray(lengths, dtype=np.int64)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.idx_base = idx_base
        cu = np.concatenate([[0], np.cumsum(self.lengths)])
        self.chunks = []
        for lo in range(0, len(self.lengths), chunk_clips):
            hi = min(lo + chunk_clips, len(self.lengths))
            self.chunks.append((lo, hi, int(cu[lo]), int(cu[hi]), ops.Layout.from_lengths(self.lengths[lo:hi])))
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


def retrieve_topk_streamed(q_host: torch.Tensor, q_layout: ops.Layout, gallery: StreamedGallery, k: int = 10,
                           mode: str = "max_t_mean_w", queries_are: str = "gesture",
                           q_dev: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery clips per query; queries and gallery start in (pinned) host memory.
    Returns DEVICE tensors (values [Q, k], global indices [Q, k]); call .cpu() to finish the round trip.
    ``q_dev`` may carry queries that are already on the device (e.g. after a broadcast)."""
    dev = gallery.dev
    main = torch.cuda.current_stream(dev)
    if q_dev is None:
        q_dev = q_host.to(dev, non_blocking=True)
    q16, _ = ops.prep(q_dev, q_layout)
    nq = q_layout.n_clips
    if not gallery.chunks:
        return (torch.full((nq, k), float("-inf"), device=dev), torch.full((nq, k), -1, dtype=torch.int32, device=dev))
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

    issue_copy(0)
    for c, (lo, hi, r0, r1, lay) in enumerate(gallery.chunks):
        b = c & 1
        if c + 1 < len(gallery.chunks):
            issue_copy(c + 1)
        main.wait_event(gallery.copied[b])
        g16 = gallery.op16[b][: r1 - r0]
        ops.prep(gallery.stage[b][: r1 - r0], lay, out=g16)
        gallery.consumed[b].record(main)
        if queries_are == "gesture":
            s = ops.simpool_allpairs(q16, q_layout, g16, lay, mode)
        else:
            s = ops.simpool_allpairs(g16, lay, q16, q_layout, mode, content_major=True)
        v, i = ops.topk(s, k, idx_offset=gallery.idx_base + lo)
        vals[c].copy_(v)
        idxs[c].copy_(i)
    if len(gallery.chunks) == 1:
        return vals[0], idxs[0]
    return ops.topk_merge(vals, idxs)
