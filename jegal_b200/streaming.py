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


def balanced_schedule(n_clips: int, n_chunks: int = 16, min_clips: int = 128):
    """`n_chunks` equal chunks.  The host->device copies run back to back on their own stream and chunk i is scored while
    chunk i + 1 arrives, so the step costs copy(first chunk) + max(all copies, all scoring) + scoring(last chunk) PLUS
    a bubble whenever a chunk's copy takes longer than the scoring of the chunk before it.  When scoring is the
    bottleneck (N <= 2: K1 needs 2x the gallery's PCIe time) only the first copy is exposed; when the two rates are
    close (N = 8 on a host whose eight links share ~236 GB/s, profiles/h2d_ceiling_r02_n8.json: 142 MB take 6.1 ms on
    the slow links against 5.1 ms of K1) every large chunk stalls the GPU for the difference -- the measured timeline of
    a (1, 2, 5, 4, 2, 1, 1)/16 schedule showed a 1.3 ms bubble behind its 5/16 chunk (profiles/e2e_timeline_r02_*).
    Sixteen equal chunks expose 1/16 of either side whatever the ratio; more would only add launch tails."""
    n = int(n_clips)
    if n <= 0:
        return []
    k = max(1, min(int(n_chunks), n // max(1, min_clips)))
    base, extra = divmod(n, k)
    out = [base + (1 if i < extra else 0) for i in range(k)]
    assert sum(out) == n and min(out) > 0, (n, out)
    return out


class StreamedGallery:
    """Pinned host gallery (packed fp16/fp32 rows + per-clip lengths) cut into clip-aligned chunks."""

    def __init__(self, rows_host: torch.Tensor, lengths: np.ndarray, chunk_clips: int = 8192, device=None,
                 idx_base: int = 0, ramp: bool = True, schedule=None):
        if rows_host.is_cuda or rows_host.dim() != 2 or rows_host.shape[1] != 512:
            raise JegalError("StreamedGallery: rows_host must be a host [rows, 512] tensor")
        self.rows = rows_host if rows_host.is_pinned() else rows_host.pin_memory()
        self.lengths = np.asarray(lengths, dtype=np.int64)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.idx_base = idx_base
        cu = np.concatenate([[0], np.cumsum(self.lengths)])
        self.chunks = []
        lo = 0
        sched = list(schedule) if schedule is not None else chunk_schedule(len(self.lengths), chunk_clips, ramp)
        if sum(sched) != len(self.lengths):
            raise JegalError(f"StreamedGallery: chunk schedule covers {sum(sched)} of {len(self.lengths)} clips")
        for size in sched:
            hi = lo + size
            self.chunks.append((lo, hi, int(cu[lo]), int(cu[hi]), ops.Layout.from_lengths(self.lengths[lo:hi])))
            lo = hi
        max_rows = max((c[3] - c[2] for c in self.chunks), default=0)
        # The raw rows land in a device buffer as large as the whole gallery (shard): no copy ever waits for the
        # scoring of an earlier chunk to release a staging buffer, so the PCIe link runs back to back -- with two
        # chunk-sized staging buffers the copies were throttled by the compute pipeline, and at N >= 4, where the
        # host's shared H2D ceiling makes the copies the bottleneck, every such bubble lengthened the step.
        self.stage = torch.empty((int(cu[-1]), 512), dtype=self.rows.dtype, device=self.dev)
        self.op16 = [torch.empty((max_rows, 512), dtype=torch.bfloat16, device=self.dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.copied = [torch.cuda.Event() for _ in self.chunks]
        self.n_clips = len(self.lengths)

    @property
    def nbytes(self) -> int:
        return self.rows.numel() * self.rows.element_size()


def shared_host_tensor(path: str, shape, dtype: torch.dtype, create: bool) -> torch.Tensor:
    """A host tensor backed by a shared-memory file (e.g. under /dev/shm) and registered with CUDA as page-locked:
    every process of the box that maps the same path sees the same bytes and can start asynchronous H2D copies
    from them.  The creator sizes the file; call ``release_shared_host_tensor`` before unlinking it."""
    n = int(np.prod(shape))
    item = torch.empty((), dtype=dtype).element_size()
    if create:
        with open(path, "wb") as f:
            f.truncate(max(n * item, 1))
    t = torch.from_file(path, shared=True, size=n, dtype=dtype).view(*shape)
    if n and torch.cuda.is_available():
        rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), n * item, 0)
        if int(rc) != 0:
            raise JegalError(f"cudaHostRegister({path}) failed with {int(rc)}")
    return t


def release_shared_host_tensor(t: torch.Tensor) -> None:
    if t.numel() and torch.cuda.is_available():
        torch.cuda.cudart().cudaHostUnregister(t.data_ptr())


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
                           q_dtype: torch.dtype = torch.float16, q_gather=False,
                           timeline: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k gallery clips per query; queries and gallery start in (pinned) host memory.
    Returns DEVICE tensors (values [Q, k], global indices [Q, k]); call .cpu() to finish the round trip.
    ``q_dev`` may carry queries that are already on the device (e.g. after a broadcast).

    ``q_parts`` > 1 pipelines the QUERY side too: the query clips are cut into contiguous parts, part p+1
    is copied (and, with ``bcast_src`` set in a torch.distributed job, broadcast from that rank over NCCL)
    while part p is being scored against the first gallery chunk, so scoring starts after 1/q_parts of the
    query transfer instead of all of it.  On every rank but ``bcast_src`` ``q_host`` may be None (its dtype
    is then ``q_dtype``).

    ``q_gather`` (torch.distributed job, ``q_host`` visible to EVERY rank, e.g. ``shared_host_tensor``): the query
    transfer is sharded like the gallery -- rank r copies rows r/N .. (r+1)/N of the queries over ITS OWN PCIe link and
    one all-gather over NVLink replicates them, so the 65 MB query set costs every link 1/N of its copy time
    instead of sitting in front of rank 0's gallery stream (1.2 ms of a 5.7 ms step at N = 8).  ``q_gather=True``
    uses an NCCL all-gather of the raw rows followed by K0 on every rank; passing an ``ops.QueryGather`` instead
    fuses normalise + cast + all-gather into one kernel over peer memory (C2): each rank prepares only its slice
    and stores the operand rows straight into every rank's buffer."""
    dev = gallery.dev
    main = torch.cuda.current_stream(dev)
    nq = q_layout.n_clips
    import torch.distributed as dist

    def mark(name, stream=None):  # timeline: CUDA events, read by `timeline_ms` after the caller synchronised
        if timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream or main)
            timeline.setdefault("events", []).append((name, ev))

    mark("start")

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    bcast = bcast_src is not None and multi
    fused_gather = isinstance(q_gather, ops.QueryGather) and q_dev is None
    gather = bool(q_gather) and multi and q_dev is None and not fused_gather
    q_ready16 = None  # the prepared query operand when it does not come from K0 below
    parts = _query_parts(q_layout, q_parts if (q_dev is None and not gather) else 1)
    ready = [None] * len(parts)  # per part: a CUDA event (copy) or an NCCL work handle (broadcast / all-gather)
    if fused_gather:
        if q_host is None:
            raise JegalError("q_gather needs the queries in host memory every rank can read")
        qg = q_gather
        r0, r1 = qg.slice_rows()
        mine = torch.empty((max(r1 - r0, 0), 512), dtype=q_host.dtype, device=dev)
        if not hasattr(gallery, "q_stream"):
            gallery.q_stream = torch.cuda.Stream(device=dev)
        gallery.q_stream.wait_stream(main)
        with torch.cuda.stream(gallery.q_stream):
            if r1 > r0:
                mine.copy_(q_host[r0:r1], non_blocking=True)
            mark("query slice copied", gallery.q_stream)
            q_ready16 = qg.prep_gather(mine, r0)  # K0 on the slice + NVLink stores to every rank + flag wait
            ready[0] = torch.cuda.Event()
            ready[0].record(gallery.q_stream)
        mine.record_stream(gallery.q_stream)
        parts = _query_parts(q_layout, 1)
        q_dev = q_ready16  # (only its shape matters below)
    elif gather:
        if q_host is None:
            raise JegalError("q_gather needs the queries in host memory every rank can read")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        rows = q_layout.rows
        per = (rows + world - 1) // world
        q_all = torch.empty((per * world, 512), dtype=q_host.dtype, device=dev)
        mine = torch.empty((per, 512), dtype=q_host.dtype, device=dev)
        r0 = min(rank * per, rows)
        r1 = min(r0 + per, rows)
        if not hasattr(gallery, "q_stream"):
            gallery.q_stream = torch.cuda.Stream(device=dev)
        gallery.q_stream.wait_stream(main)
        with torch.cuda.stream(gallery.q_stream):
            if r1 > r0:
                mine[: r1 - r0].copy_(q_host[r0:r1], non_blocking=True)
            mark("query slice copied", gallery.q_stream)
            ready[0] = dist.all_gather_into_tensor(q_all, mine, group=group, async_op=True)
        q_all.record_stream(gallery.q_stream)
        mine.record_stream(gallery.q_stream)
        q_dev = q_all[:rows]
    elif q_dev is None:
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
    q16 = q_ready16 if q_ready16 is not None else torch.empty((q_layout.rows, 512), dtype=torch.bfloat16, device=dev)
    vals = torch.empty((len(gallery.chunks), nq, k), dtype=torch.float32, device=dev)
    idxs = torch.empty((len(gallery.chunks), nq, k), dtype=torch.int32, device=dev)

    # every chunk's copy is enqueued up front, back to back on the copy stream; the device buffer may still be read
    # by the previous call's K0 on `main` (the function returns device tensors without synchronising)
    gallery.copy_stream.wait_stream(main)
    with torch.cuda.stream(gallery.copy_stream):
        for c, (lo, hi, r0, r1, _) in enumerate(gallery.chunks):
            gallery.stage[r0:r1].copy_(gallery.rows[r0:r1], non_blocking=True)
            gallery.copied[c].record(gallery.copy_stream)
            mark(f"chunk {c} copied", gallery.copy_stream)
    for c, (lo, hi, r0, r1, lay) in enumerate(gallery.chunks):
        b = c & 1
        main.wait_event(gallery.copied[c])
        g16 = gallery.op16[b][: r1 - r0]
        ops.prep(gallery.stage[r0:r1], lay, out=g16)
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
                if q_ready16 is None:
                    ops.prep(q_dev[qr0:qr1], qlay, out=q16[qr0:qr1])
                mark(f"query part {p} ready")
            if queries_are == "gesture":
                s = ops.simpool_allpairs(q16[qr0:qr1], qlay, g16, lay, mode)
            else:
                s = ops.simpool_allpairs(g16, lay, q16[qr0:qr1], qlay, mode, content_major=True)
            v, i = ops.topk(s, k, idx_offset=gallery.idx_base + lo)
            vals[c, qlo:qhi].copy_(v)
            idxs[c, qlo:qhi].copy_(i)
        mark(f"chunk {c} scored")
    if len(gallery.chunks) == 1:
        return vals[0], idxs[0]
    out = ops.topk_merge(vals, idxs)
    mark("chunk lists merged")
    return out


def timeline_ms(timeline: dict) -> list:
    """[(name, milliseconds since "start")] of a timeline filled by retrieve_topk_streamed; call after a synchronise."""
    ev = timeline.get("events", [])
    if not ev:
        return []
    t0 = ev[0][1]
    return [(name, round(t0.elapsed_time(e), 3)) for name, e in ev]


# ---------------------------------------------------------------------------------------------------------------
# Host-to-host spotting / active-speaker scoring: the clip set sits in page-locked host memory (a packed index,
# jegal_b200.index), its rows stream to the device in clip-aligned chunks on a copy stream and K3 / K0-means run on
# chunk i while chunk i + 1 is in flight.  With the normalisation fused into K3's operand load there is no other
# pass over the rows: the step is the PCIe transfer plus the kernels' tail on the last chunk.
class HostClips:
    """One side of a clip set: packed rows in page-locked host memory + a persistent device buffer of the same size."""

    def __init__(self, rows_host: torch.Tensor, lengths, device=None):
        if rows_host.is_cuda or rows_host.dim() != 2 or rows_host.shape[1] != 512:
            raise JegalError("HostClips: rows_host must be a host [rows, 512] tensor")
        self.rows = rows_host if rows_host.is_pinned() else rows_host.pin_memory()
        self.lengths = np.asarray(lengths, dtype=np.int64)
        self.cu = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        if int(self.cu[-1]) != self.rows.shape[0]:
            raise JegalError("HostClips: lengths do not add up to the row count")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.rows_dev = torch.empty(self.rows.shape, dtype=self.rows.dtype, device=self.dev)
        self._layouts = {}

    @classmethod
    def from_index(cls, idx, device=None) -> "HostClips":
        return cls(idx.pinned(), idx.lengths, device)

    @property
    def n(self) -> int:
        return len(self.lengths)

    @property
    def nbytes(self) -> int:
        return self.rows.numel() * self.rows.element_size()

    def layout(self, lo: int, hi: int) -> ops.Layout:
        key = (lo, hi)
        if key not in self._layouts:  # layouts own device tables: built once per chunk boundary
            self._layouts[key] = ops.Layout.from_lengths(self.lengths[lo:hi])
        return self._layouts[key]

    def chunk(self, lo: int, hi: int):
        r0, r1 = int(self.cu[lo]), int(self.cu[hi])
        return self.rows_dev[r0:r1], self.layout(lo, hi)


def clip_chunks(sides, n_chunks: int):
    """Clip ranges [lo, hi) that cut the clip set into ~equal BYTE shares over all `sides` (HostClips that list the
    same clips); the first chunk is half size so that scoring starts early."""
    n = sides[0].n
    if n == 0:
        return []
    per_clip = sum(s.lengths * s.rows.element_size() for s in sides).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(per_clip)])
    n_chunks = max(1, min(int(n_chunks), n))
    targets = cum[-1] * (np.arange(1, n_chunks * 2) / (n_chunks * 2.0))
    cuts = np.searchsorted(cum, targets)
    cuts = np.unique(np.concatenate([[0], cuts[:1], cuts[1::2], [n]]))  # first cut at 1/(2 n_chunks), then every 1/n_chunks
    return [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def _copy_chunks(sides, chunks, copy_stream, main):
    """Enqueue the H2D copies of every chunk (all sides) on `copy_stream`; one event per chunk."""
    copy_stream.wait_stream(main)  # the device buffers may still be read by kernels of the previous call
    events = []
    with torch.cuda.stream(copy_stream):
        for lo, hi in chunks:
            for s in sides:
                r0, r1 = int(s.cu[lo]), int(s.cu[hi])
                if r1 > r0:
                    s.rows_dev[r0:r1].copy_(s.rows[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            events.append(ev)
    return events


_copy_streams = {}


def _copy_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=dev)
    return _copy_streams[key]


def spot_streamed(g: HostClips, c: HostClips, word_idx, windows=None, thresh: float = 0.5, temp: float = 0.07,
                  normalize: bool = True, n_chunks: int = 8, want_heat: bool = False) -> dict:
    """evaluate_spotting.py:59-90 for a clip set in pinned host memory: H2D of chunk i + 1 overlaps K3 on chunk i.
    Returns HOST numpy arrays: pred_frame, pred_score, correct (if windows), heat (flat [sum T], if asked)."""
    if g.n != c.n:
        raise JegalError("spot_streamed: gesture and content sides must list the same clips")
    if c.n and int(c.lengths.max()) > 64:
        raise JegalError("spot_streamed: a clip has more than 64 words; use scoring.spot_batch for such sets")
    dev = g.dev
    main = torch.cuda.current_stream(dev)
    chunks = clip_chunks([g, c], n_chunks)
    wi = torch.as_tensor(np.asarray(word_idx, dtype=np.int32)).pin_memory().to(dev, non_blocking=True)
    lo_d = hi_d = None
    if windows is not None:
        lo_d = torch.as_tensor(np.asarray(windows[0], dtype=np.int32)).pin_memory().to(dev, non_blocking=True)
        hi_d = torch.as_tensor(np.asarray(windows[1], dtype=np.int32)).pin_memory().to(dev, non_blocking=True)
    events = _copy_chunks([g, c], chunks, _copy_stream(dev), main)
    outs = []
    fused = g.rows.dtype in (torch.float16, torch.bfloat16) and c.rows.dtype == g.rows.dtype
    for (lo, hi), ev in zip(chunks, events):
        main.wait_event(ev)
        g_rows, gl = g.chunk(lo, hi)
        c_rows, cl = c.chunk(lo, hi)
        kn = normalize
        if not fused:  # fp32 storage: one K0 pass (normalise + cast) per chunk
            g_rows = ops.prep(g_rows, gl, normalize=normalize)[0]
            c_rows = ops.prep(c_rows, cl, normalize=normalize)[0]
            kn = False
        outs.append(ops.spot(g_rows, gl, c_rows, cl, wi[lo:hi], tau=temp, want_heat=want_heat,
                             win_lo=None if lo_d is None else lo_d[lo:hi], win_hi=None if hi_d is None else hi_d[lo:hi],
                             thresh=thresh, normalize=kn))
    res = {}
    for k in ("pred_frame", "pred_score", "correct", "heat"):
        parts = [o[k] for o in outs if o[k] is not None]
        res[k] = torch.cat(parts).cpu().numpy() if parts else None
    if res["correct"] is not None:
        res["correct"] = res["correct"].astype(bool)
    return res


def asd_streamed(g: HostClips, c: HostClips, pair_gest, pair_cont, tracks: int, prefixes=(2, 4, 6), temp: float = 0.07,
                 mode: str = "reference", n_chunks: int = 8) -> dict:
    """evaluate_asd.py:54-127 for clip sets in pinned host memory.  mode "reference": the clip means of chunk i
    (one read-only pass, K0) are formed while chunk i + 1 is copied, then one warp per pair takes the cosine;
    a pooling mode name: K4 (normalisation fused into the load) on the pairs of every gesture chunk, against the
    content side, which is copied first.  Returns host arrays: scores [n_groups, tracks], pred {P: [n_groups]}."""
    dev = g.dev
    main = torch.cuda.current_stream(dev)
    pg_h, pc_h = np.asarray(pair_gest, dtype=np.int32), np.asarray(pair_cont, dtype=np.int32)
    pg = torch.as_tensor(pg_h).pin_memory().to(dev, non_blocking=True)
    pc = torch.as_tensor(pc_h).pin_memory().to(dev, non_blocking=True)
    cs = _copy_stream(dev)
    c_chunks = clip_chunks([c], max(1, n_chunks // 4))
    g_chunks = clip_chunks([g], n_chunks)
    cs.wait_stream(main)
    ev_c = _copy_chunks([c], c_chunks, cs, main)
    ev_g = _copy_chunks([g], g_chunks, cs, main)
    if mode == "reference":
        gm = torch.empty((g.n, 512), dtype=torch.float32, device=dev)
        cm = torch.empty((c.n, 512), dtype=torch.float32, device=dev)
        for side, chunks, events, out in ((c, c_chunks, ev_c, cm), (g, g_chunks, ev_g, gm)):
            for (lo, hi), ev in zip(chunks, events):
                main.wait_event(ev)
                rows, lay = side.chunk(lo, hi)
                ops.clip_means(rows, lay, mean_eps=1e-8, out=out[lo:hi])
        scores = ops.pair_cosine(gm, cm, pg, pc, normalize=False)
    else:
        if c.n and int(c.lengths.max()) > 64:
            raise JegalError("asd_streamed: a content clip has more than 64 words; use scoring.asd_batch")
        if g.rows.dtype not in (torch.float16, torch.bfloat16) or c.rows.dtype != g.rows.dtype:
            raise JegalError("asd_streamed: pooling modes take 16-bit stored rows (same type on both sides)")
        for ev in ev_c:
            main.wait_event(ev)
        c_rows, cl = c.chunk(0, c.n)
        scores = torch.empty((pg_h.size,), dtype=torch.float32, device=dev)
        order = np.argsort(pg_h, kind="stable")
        bounds = np.searchsorted(pg_h[order], [lo for lo, _ in g_chunks] + [g.n])
        for k, ((lo, hi), ev) in enumerate(zip(g_chunks, ev_g)):
            main.wait_event(ev)
            sel = order[bounds[k]:bounds[k + 1]]
            if sel.size == 0:
                continue
            g_rows, gl = g.chunk(lo, hi)
            contiguous = sel.size == sel[-1] - sel[0] + 1 and np.all(np.diff(sel) == 1)
            sel_d = None if contiguous else torch.from_numpy(sel).to(dev)
            pg_l = (pg[int(sel[0]):int(sel[-1]) + 1] if contiguous else pg[sel_d]) - lo
            pc_l = pc[int(sel[0]):int(sel[-1]) + 1] if contiguous else pc[sel_d]
            sc = ops.simpool_pairs(g_rows, gl, c_rows, cl, pg_l.contiguous(), pc_l.contiguous(), mode, normalize=True)["scores"]
            if contiguous:
                scores[int(sel[0]):int(sel[-1]) + 1] = sc
            else:
                scores[sel_d] = sc
    n_groups = pg_h.size // tracks
    pred = {}
    for P in prefixes:
        if P <= tracks:
            pred[P] = ops.group_softmax(scores, n_groups, P, stride=tracks, tau=temp, want_probs=False)[1]
    out = dict(scores=scores.cpu().numpy().reshape(n_groups, tracks), pred={P: v.cpu().numpy() for P, v in pred.items()})
    out["acc"] = {P: (float((v == 0).mean()) if n_groups else float("nan")) for P, v in out["pred"].items()}
    return out
