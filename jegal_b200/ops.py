"""Device-tensor wrappers over the C ABI (include/jegal_b200.h).

torch is used only as the owner of device memory and streams; every compute
step is a call into libjegal_b200.so.  All functions raise ``JegalError`` when
the library or an sm_100 device is missing — there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import JegalError, POOL_MODES

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class Context:
    """One per device; owns nothing but the error string and launch counter."""

    _by_device = {}

    def __init__(self, device: int):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise JegalError("no CUDA device: jegal_b200 has no CPU path")
        h = C.c_void_p()
        rc = self.lib.jegal_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise JegalError(f"jegal_ctx_create(device={device}) failed with {rc} (sm_100 device required)")
        self.h = h
        self.device = int(device)

    @classmethod
    def get(cls, device: Optional[int] = None) -> "Context":
        if device is None:
            device = torch.cuda.current_device()
        if device not in cls._by_device:
            cls._by_device[device] = Context(device)
        return cls._by_device[device]

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.jegal_last_error(self.h).decode()
            raise JegalError(f"{what} failed ({rc}): {msg}")

    @property
    def launches(self) -> int:
        return int(self.lib.jegal_launch_count(self.h))


class Layout:
    """Ragged layout of one packed operand (clip i = rows cu_len[i]:cu_len[i+1])."""

    def __init__(self, cu_len: Sequence[int], ctx: Optional[Context] = None):
        self.ctx = ctx or Context.get()
        cu = np.ascontiguousarray(np.asarray(cu_len, dtype=np.int32))
        if cu.ndim != 1 or cu.size < 1:
            raise JegalError("cu_len must be a 1-D array of n_clips + 1 offsets")
        self.cu_len = cu
        self.n_clips = int(cu.size - 1)
        self.rows = int(cu[-1])
        h = C.c_void_p()
        rc = self.ctx.lib.jegal_layout_create(
            self.ctx.h, cu.ctypes.data_as(C.POINTER(C.c_int32)), self.n_clips, _stream(), C.byref(h)
        )
        self.ctx.check(rc, "jegal_layout_create")
        self.h = h

    @classmethod
    def from_lengths(cls, lengths: Sequence[int], ctx: Optional[Context] = None) -> "Layout":
        cu = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lengths, dtype=np.int64), out=cu[1:])
        if cu[-1] >= 2**31:
            raise JegalError("more than 2^31 rows in one operand")
        return cls(cu.astype(np.int32), ctx)

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.cu_len)

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                self.ctx.lib.jegal_layout_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _check_rows(x: torch.Tensor, layout: Layout, what: str) -> None:
    if not x.is_cuda or x.dim() != 2 or x.shape[1] != 512 or not x.is_contiguous():
        raise JegalError(f"{what}: expected a contiguous CUDA tensor of shape [rows, 512]")
    if x.shape[0] != layout.rows:
        raise JegalError(f"{what}: {x.shape[0]} rows but the layout has {layout.rows}")


def prep(
    emb: torch.Tensor,
    layout: Layout,
    normalize: bool = True,
    out_dtype: torch.dtype = torch.bfloat16,
    want_mean_scale: bool = False,
    row_eps: float = 1e-12,
    mean_eps: float = 1e-12,
    out: Optional[torch.Tensor] = None,
    want_mean_rows: bool = False,
):
    """K0: normalise + cast rows; optionally 1/||mean row|| per clip and the unit-norm
    mean row of every clip.  Returns (rows16, inv_meannorm | None[, mean_rows16])."""
    _check_rows(emb, layout, "prep")
    if emb.dtype not in _DT or out_dtype not in (torch.bfloat16, torch.float16):
        raise JegalError("prep: unsupported dtype")
    ctx = layout.ctx
    if out is None:
        out = torch.empty((layout.rows, 512), dtype=out_dtype, device=emb.device)
    scale = torch.empty((layout.n_clips,), dtype=torch.float32, device=emb.device) if want_mean_scale else None
    mean_rows = torch.empty((layout.n_clips, 512), dtype=out_dtype, device=emb.device) if want_mean_rows else None
    rc = ctx.lib.jegal_prep(
        ctx.h, layout.h, _ptr(emb), _DT[emb.dtype], int(normalize), row_eps, mean_eps, _DT[out_dtype],
        _ptr(out), _ptr(scale), _ptr(mean_rows), _stream(),
    )
    ctx.check(rc, "jegal_prep")
    if want_mean_rows:
        return out, scale, mean_rows
    return out, scale


def clip_means(emb: torch.Tensor, layout: Layout, mean_eps: float = 1e-12, out_dtype: torch.dtype = torch.float32,
               want_rows: bool = True, want_scale: bool = False, out: Optional[torch.Tensor] = None):
    """K0 without the row output (read-only pass): per clip the unit-norm mean row ([n_clips, 512] in
    ``out_dtype``) and / or 1 / max(||mean||, mean_eps).  Returns (mean_rows | None, inv_meannorm | None)."""
    _check_rows(emb, layout, "clip_means")
    if emb.dtype not in _DT or out_dtype not in _DT:
        raise JegalError("clip_means: unsupported dtype")
    ctx = layout.ctx
    rows = None
    if out is not None:
        if tuple(out.shape) != (layout.n_clips, 512) or out.dtype not in _DT or not out.is_contiguous() or not out.is_cuda:
            raise JegalError("clip_means: bad out tensor")
        rows, out_dtype = out, out.dtype
    elif want_rows:
        rows = torch.empty((layout.n_clips, 512), dtype=out_dtype, device=emb.device)
    scale = torch.empty((layout.n_clips,), dtype=torch.float32, device=emb.device) if want_scale else None
    if rows is None and scale is None:
        raise JegalError("clip_means: nothing requested")
    rc = ctx.lib.jegal_clip_means(ctx.h, layout.h, _ptr(emb), _DT[emb.dtype], float(mean_eps), _DT[out_dtype],
                                  _ptr(rows), _ptr(scale), _stream())
    ctx.check(rc, "jegal_clip_means")
    return rows, scale


def pair_cosine(a: torch.Tensor, b: torch.Tensor, pair_a: Optional[torch.Tensor] = None,
                pair_b: Optional[torch.Tensor] = None, normalize: bool = True, eps: float = 1e-8,
                n_pairs: Optional[int] = None) -> torch.Tensor:
    """Cosine (or dot product) of listed pairs of 512-wide rows (nn.CosineSimilarity, evaluate_asd.py:45-47)."""
    for t in (a, b):
        if not t.is_cuda or t.dim() != 2 or t.shape[1] != 512 or not t.is_contiguous() or t.dtype not in _DT:
            raise JegalError("pair_cosine: expected contiguous CUDA [n, 512] matrices")
    if a.dtype != b.dtype:
        raise JegalError("pair_cosine: both matrices must have the same dtype")
    for t in (pair_a, pair_b):
        if t is not None and (t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous()):
            raise JegalError("pair_cosine: pair lists must be contiguous CUDA int32")
    if n_pairs is None:
        n_pairs = int(pair_a.numel() if pair_a is not None else pair_b.numel() if pair_b is not None
                      else min(a.shape[0], b.shape[0]))
    ctx = Context.get(a.device.index)
    scores = torch.empty((n_pairs,), dtype=torch.float32, device=a.device)
    rc = ctx.lib.jegal_pair_cosine(ctx.h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], _DT[a.dtype], _ptr(pair_a),
                                   _ptr(pair_b), n_pairs, int(bool(normalize)), float(eps), _ptr(scores), _stream())
    ctx.check(rc, "jegal_pair_cosine")
    return scores


def simpool_allpairs(
    gest_rows: torch.Tensor,
    gest_layout: Layout,
    cont_rows: torch.Tensor,
    cont_layout: Layout,
    mode: str = "mean_mean",
    gscale: Optional[torch.Tensor] = None,
    cscale: Optional[torch.Tensor] = None,
    content_major: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """K1: pooled score of every (gesture clip, content clip) pair.

    Returns [n_gest, n_cont] fp32, or [n_cont, n_gest] when ``content_major``.
    """
    _check_rows(gest_rows, gest_layout, "simpool gest")
    _check_rows(cont_rows, cont_layout, "simpool cont")
    if gest_rows.dtype != cont_rows.dtype or gest_rows.dtype not in (torch.bfloat16, torch.float16):
        raise JegalError("simpool: operands must both be bf16 or both fp16 (outputs of prep)")
    ctx = gest_layout.ctx
    nG, nC = gest_layout.n_clips, cont_layout.n_clips
    shape = (nC, nG) if content_major else (nG, nC)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=gest_rows.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise JegalError("simpool: bad out tensor")
    ld_g, ld_c = (1, nG) if content_major else (nC, 1)
    rc = ctx.lib.jegal_simpool_allpairs(
        ctx.h, gest_layout.h, _ptr(gest_rows), cont_layout.h, _ptr(cont_rows), _DT[gest_rows.dtype],
        POOL_MODES[mode], _ptr(gscale), _ptr(cscale), _ptr(out), ld_g, ld_c, _stream(),
    )
    ctx.check(rc, "jegal_simpool_allpairs")
    return out


def topk(scores: torch.Tensor, k: int, idx_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2: per-row top-k (descending, ties -> lower index)."""
    if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 2 or scores.stride(1) != 1:
        raise JegalError("topk: expected a CUDA fp32 matrix with unit column stride")
    ctx = Context.get(scores.device.index)
    nq, ng = scores.shape
    val = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    idx = torch.empty((nq, k), dtype=torch.int32, device=scores.device)
    rc = ctx.lib.jegal_topk(ctx.h, _ptr(scores), nq, ng, scores.stride(0), k, idx_offset, _ptr(val), _ptr(idx), _stream())
    ctx.check(rc, "jegal_topk")
    return val, idx


def topk_merge(vals: torch.Tensor, idxs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge [n_lists, n_q, k] sorted lists into [n_q, k]."""
    if vals.shape != idxs.shape or vals.dim() != 3 or not vals.is_contiguous() or not idxs.is_contiguous():
        raise JegalError("topk_merge: expected contiguous [n_lists, n_q, k] tensors")
    if vals.dtype != torch.float32 or idxs.dtype != torch.int32:
        raise JegalError("topk_merge: fp32 values and int32 indices")
    ctx = Context.get(vals.device.index)
    nl, nq, k = vals.shape
    oval = torch.empty((nq, k), dtype=torch.float32, device=vals.device)
    oidx = torch.empty((nq, k), dtype=torch.int32, device=vals.device)
    rc = ctx.lib.jegal_topk_merge(ctx.h, _ptr(vals), _ptr(idxs), nl, nq, k, _ptr(oval), _ptr(oidx), _stream())
    ctx.check(rc, "jegal_topk_merge")
    return oval, oidx


def rank_of_positive(
    scores: torch.Tensor, gt: Optional[torch.Tensor] = None
) -> Tuple[torch.Tensor, torch.Tensor]:
    """(#entries > positive, #entries == positive) per row; any 2-D strides."""
    if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 2:
        raise JegalError("rank_of_positive: expected a CUDA fp32 matrix")
    ctx = Context.get(scores.device.index)
    nq, ng = scores.shape
    ngt = torch.empty((nq,), dtype=torch.int32, device=scores.device)
    neq = torch.empty((nq,), dtype=torch.int32, device=scores.device)
    if gt is not None and (gt.dtype != torch.int32 or not gt.is_contiguous()):
        raise JegalError("rank_of_positive: gt must be contiguous int32")
    rc = ctx.lib.jegal_rank_of_positive(
        ctx.h, _ptr(scores), nq, ng, scores.stride(0), scores.stride(1), _ptr(gt), _ptr(ngt), _ptr(neq), _stream()
    )
    ctx.check(rc, "jegal_rank_of_positive")
    return ngt, neq


def spot(
    gest_rows: torch.Tensor,
    gest_layout: Layout,
    cont_rows: torch.Tensor,
    cont_layout: Layout,
    word_idx: torch.Tensor,
    tau: float = 0.07,
    want_heat: bool = True,
    want_full: bool = False,
    win_lo: Optional[torch.Tensor] = None,
    win_hi: Optional[torch.Tensor] = None,
    thresh: float = 0.5,
    normalize: bool = False,
    row_eps: float = 1e-12,
) -> dict:
    """K3: word spotting over n clips (clip i of both layouts).

    ``normalize=True``: the operands are the rows as stored (fp16 / bf16) and the kernel fuses
    F.normalize (evaluate_spotting.py:49-50) into the load; False: rows are used as they are.

    Returns dict(heat [sum T] | None, full [sum T_i*W_i] | None, full_off int64 [n+1] | None,
    pred_frame int32 [n], pred_score fp32 [n], correct uint8 [n] | None).
    """
    _check_rows(gest_rows, gest_layout, "spot gest")
    _check_rows(cont_rows, cont_layout, "spot cont")
    if gest_rows.dtype != cont_rows.dtype or gest_rows.dtype not in (torch.bfloat16, torch.float16):
        raise JegalError("spot: operands must both be bf16 or both fp16 (stored rows or outputs of prep)")
    ctx = gest_layout.ctx
    n = gest_layout.n_clips
    dev = gest_rows.device
    if word_idx.dtype != torch.int32 or word_idx.numel() != n or not word_idx.is_cuda:
        raise JegalError("spot: word_idx must be CUDA int32 [n_clips]")
    heat = torch.empty((gest_layout.rows,), dtype=torch.float32, device=dev) if want_heat else None
    full = full_off = None
    if want_full:
        sizes = gest_layout.lengths.astype(np.int64) * cont_layout.lengths.astype(np.int64)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        full_off = torch.from_numpy(off).to(dev)
        full = torch.empty((int(off[-1]),), dtype=torch.float32, device=dev)
    pred_frame = torch.empty((n,), dtype=torch.int32, device=dev)
    pred_score = torch.empty((n,), dtype=torch.float32, device=dev)
    correct = None
    if win_lo is not None:
        if win_hi is None or win_lo.dtype != torch.int32 or win_hi.dtype != torch.int32:
            raise JegalError("spot: win_lo/win_hi must both be int32")
        correct = torch.empty((n,), dtype=torch.uint8, device=dev)
    rc = ctx.lib.jegal_spot(
        ctx.h, gest_layout.h, _ptr(gest_rows), cont_layout.h, _ptr(cont_rows), _DT[gest_rows.dtype],
        int(bool(normalize)), float(row_eps), _ptr(word_idx), float(tau), _ptr(heat), _ptr(full), _ptr(full_off), _ptr(pred_frame), _ptr(pred_score),
        _ptr(win_lo), _ptr(win_hi), float(thresh), _ptr(correct), _stream(),
    )
    ctx.check(rc, "jegal_spot")
    return dict(heat=heat, full=full, full_off=full_off, pred_frame=pred_frame, pred_score=pred_score, correct=correct)


def spot_dense(cos: torch.Tensor, gest_layout: Layout, cont_layout: Layout, word_idx: torch.Tensor, tau: float = 0.07,
               want_full: bool = False, win_lo: Optional[torch.Tensor] = None, win_hi: Optional[torch.Tensor] = None,
               thresh: float = 0.5) -> dict:
    """K3's outputs from a dense [sum T, sum W] frame x word cosine matrix of a group of clips (only the block
    diagonal is read): the route for clips with more than 64 words.  Same dict as ``spot``."""
    if not cos.is_cuda or cos.dtype != torch.float32 or cos.dim() != 2 or cos.stride(1) != 1:
        raise JegalError("spot_dense: expected a CUDA fp32 matrix with unit column stride")
    if cos.shape[0] != gest_layout.rows or cos.shape[1] != cont_layout.rows:
        raise JegalError("spot_dense: the matrix must be [gesture rows, content rows] of the two layouts")
    ctx = gest_layout.ctx
    n, dev = gest_layout.n_clips, cos.device
    if word_idx.dtype != torch.int32 or word_idx.numel() != n or not word_idx.is_cuda:
        raise JegalError("spot_dense: word_idx must be CUDA int32 [n_clips]")
    heat = torch.empty((gest_layout.rows,), dtype=torch.float32, device=dev)
    full = full_off = None
    if want_full:
        sizes = gest_layout.lengths.astype(np.int64) * cont_layout.lengths.astype(np.int64)
        off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(sizes, out=off[1:])
        full_off = torch.from_numpy(off).to(dev)
        full = torch.empty((int(off[-1]),), dtype=torch.float32, device=dev)
    pred_frame = torch.empty((n,), dtype=torch.int32, device=dev)
    pred_score = torch.empty((n,), dtype=torch.float32, device=dev)
    correct = torch.empty((n,), dtype=torch.uint8, device=dev) if win_lo is not None else None
    rc = ctx.lib.jegal_spot_dense(ctx.h, _ptr(cos), cos.stride(0), gest_layout.h, cont_layout.h, _ptr(word_idx), float(tau),
                                  _ptr(heat), _ptr(full), _ptr(full_off), _ptr(pred_frame), _ptr(pred_score), _ptr(win_lo),
                                  _ptr(win_hi), float(thresh), _ptr(correct), _stream())
    ctx.check(rc, "jegal_spot_dense")
    return dict(heat=heat, full=full, full_off=full_off, pred_frame=pred_frame, pred_score=pred_score, correct=correct)


def simpool_pairs(
    gest_rows: torch.Tensor,
    gest_layout: Layout,
    cont_rows: torch.Tensor,
    cont_layout: Layout,
    pair_gest: Optional[torch.Tensor],
    pair_cont: Optional[torch.Tensor],
    mode: str = "mean_mean",
    gscale: Optional[torch.Tensor] = None,
    cscale: Optional[torch.Tensor] = None,
    group_size: int = 0,
    tau: float = 0.07,
    want_probs: bool = False,
    n_pairs: Optional[int] = None,
    normalize: bool = False,
    row_eps: float = 1e-12,
) -> dict:
    """K4: pooled scores of listed (gesture clip, content clip) pairs, optional per-group
    softmax(score / tau) and argmax (groups of `group_size` consecutive pairs)."""
    _check_rows(gest_rows, gest_layout, "simpool_pairs gest")
    _check_rows(cont_rows, cont_layout, "simpool_pairs cont")
    if gest_rows.dtype != cont_rows.dtype or gest_rows.dtype not in (torch.bfloat16, torch.float16):
        raise JegalError("simpool_pairs: operands must both be bf16 or both fp16 (stored rows or outputs of prep)")
    ctx = gest_layout.ctx
    dev = gest_rows.device
    for t in (pair_gest, pair_cont):
        if t is not None and (t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous()):
            raise JegalError("simpool_pairs: pair lists must be contiguous CUDA int32")
    if n_pairs is None:
        n_pairs = int(pair_gest.numel() if pair_gest is not None else pair_cont.numel() if pair_cont is not None
                      else gest_layout.n_clips)
    scores = torch.empty((n_pairs,), dtype=torch.float32, device=dev)
    probs = argmax = None
    if group_size > 0:
        argmax = torch.empty((n_pairs // group_size,), dtype=torch.int32, device=dev)
        if want_probs:
            probs = torch.empty((n_pairs,), dtype=torch.float32, device=dev)
    rc = ctx.lib.jegal_simpool_pairs(
        ctx.h, gest_layout.h, _ptr(gest_rows), cont_layout.h, _ptr(cont_rows), _DT[gest_rows.dtype],
        int(bool(normalize)), float(row_eps), POOL_MODES[mode], _ptr(gscale), _ptr(cscale), _ptr(pair_gest),
        _ptr(pair_cont), n_pairs,
        max(group_size, 1), float(tau), _ptr(scores), _ptr(probs), _ptr(argmax), _stream(),
    )
    ctx.check(rc, "jegal_simpool_pairs")
    return dict(scores=scores, probs=probs, argmax=argmax)


def group_softmax(scores: torch.Tensor, n_groups: int, group_size: int, stride: Optional[int] = None,
                  tau: float = 0.07, want_probs: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    """softmax(scores / tau) + first argmax inside groups (group g = scores[g*stride : g*stride+group_size])."""
    if not scores.is_cuda or scores.dtype != torch.float32 or not scores.is_contiguous():
        raise JegalError("group_softmax: expected a contiguous CUDA fp32 tensor")
    stride = group_size if stride is None else stride
    if scores.numel() < (n_groups - 1) * stride + group_size:
        raise JegalError("group_softmax: scores too short")
    ctx = Context.get(scores.device.index)
    probs = torch.empty((n_groups, group_size), dtype=torch.float32, device=scores.device) if want_probs else None
    argmax = torch.empty((n_groups,), dtype=torch.int32, device=scores.device)
    rc = ctx.lib.jegal_group_softmax(ctx.h, _ptr(scores), n_groups, group_size, stride, float(tau), _ptr(probs),
                                     _ptr(argmax), _stream())
    ctx.check(rc, "jegal_group_softmax")
    return probs, argmax


def segment_mean(x: torch.Tensor, seg_begin: torch.Tensor, seg_end: torch.Tensor, out: Optional[torch.Tensor] = None,
                 col_off: int = 0, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """K5: out[s, col_off:col_off+D] = mean(x[seg_begin[s]:seg_end[s]], dim=0) for every segment, one launch.

    ``x`` is a dense CUDA [rows, D] matrix (fp32/fp16/bf16, D a multiple of 8); ``seg_begin`` / ``seg_end`` are
    int32 CUDA vectors (end exclusive, ranges may overlap).  ``out`` may be a wider [n_seg, ld] matrix that
    receives the result at column ``col_off`` (the concat fusion of models/jegal.py:405-406)."""
    if not x.is_cuda or x.dim() != 2 or not x.is_contiguous() or x.dtype not in _DT:
        raise JegalError("segment_mean: x must be a contiguous CUDA [rows, D] fp32/fp16/bf16 matrix")
    for t in (seg_begin, seg_end):
        if not t.is_cuda or t.dtype != torch.int32 or t.dim() != 1 or not t.is_contiguous():
            raise JegalError("segment_mean: seg_begin / seg_end must be contiguous CUDA int32 vectors")
    n = seg_begin.numel()
    if seg_end.numel() != n:
        raise JegalError("segment_mean: seg_begin and seg_end differ in length")
    rows, dim = x.shape
    if out is None:
        out = torch.empty((n, col_off + dim), dtype=out_dtype or x.dtype, device=x.device)
    if not out.is_cuda or out.dim() != 2 or out.shape[0] != n or out.stride(1) != 1 or out.dtype not in _DT:
        raise JegalError("segment_mean: bad out tensor")
    ctx = Context.get(x.device.index)
    rc = ctx.lib.jegal_segment_mean(ctx.h, _ptr(x), _DT[x.dtype], rows, dim, _ptr(seg_begin), _ptr(seg_end), n,
                                    _ptr(out), _DT[out.dtype], out.stride(0) if n > 0 else out.shape[1], col_off,
                                    _stream())
    ctx.check(rc, "jegal_segment_mean")
    return out


class TopkExchange:
    """C1: K2 + NVLink peer-memory exchange + merge for a gallery sharded over the ranks of one box.

    Construction is collective (every rank of ``group`` must call it): the CUDA IPC handles of the
    per-rank exchange blocks are all-gathered through torch.distributed and opened.
    """

    def __init__(self, n_q: int, k: int, group=None, device: Optional[int] = None):
        import torch.distributed as dist

        self.ctx = Context.get(device)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_q, self.k = int(n_q), int(k)
        h = C.c_void_p()
        rc = self.ctx.lib.jegal_exchange_create(self.ctx.h, self.rank, self.world, self.n_q, self.k, C.byref(h))
        self.ctx.check(rc, "jegal_exchange_create")
        self.h = h
        mine = (C.c_uint8 * 64)()
        self.ctx.check(self.ctx.lib.jegal_exchange_ipc_handle(self.h, mine), "jegal_exchange_ipc_handle")
        handles = [bytes(mine)]
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            handles = gathered
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.ctx.check(self.ctx.lib.jegal_exchange_connect(self.h, buf), "jegal_exchange_connect")
        if self.world > 1:
            dist.barrier(group=group)

    def topk(self, scores: torch.Tensor, idx_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """scores: this rank's [n_q, n_local] fp32 shard.  Returns the GLOBAL top-k on every rank."""
        if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 2 or scores.shape[0] != self.n_q:
            raise JegalError("TopkExchange.topk: expected CUDA fp32 [n_q, n_local]")
        if scores.shape[1] > 0 and scores.stride(1) != 1:
            raise JegalError("TopkExchange.topk: unit column stride required")
        val = torch.empty((self.n_q, self.k), dtype=torch.float32, device=scores.device)
        idx = torch.empty((self.n_q, self.k), dtype=torch.int32, device=scores.device)
        rc = self.ctx.lib.jegal_topk_exchange(self.ctx.h, self.h, _ptr(scores), scores.shape[1],
                                              max(scores.stride(0), scores.shape[1]), idx_offset, _ptr(val), _ptr(idx), _stream())
        self.ctx.check(rc, "jegal_topk_exchange")
        return val, idx

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                self.ctx.lib.jegal_exchange_destroy(self.h)
                self.h = None
        except Exception:
            pass


class _DevArray:
    """A device allocation owned by the library, exposed through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class QueryGather:
    """C2: normalise + cast + all-gather of the replicated (query-side) operand over NVLink peer memory, one kernel.

    Construction is collective (every rank of ``group``): the CUDA IPC handles of the per-rank operand blocks are
    all-gathered through torch.distributed and opened.  ``prep_gather(slice, row0)`` takes this rank's rows
    [row0, row0 + len(slice)) of the raw [rows, 512] matrix and returns the complete normalised 16-bit operand
    (a view of library-owned memory, valid until the call after next)."""

    def __init__(self, rows: int, group=None, device: Optional[int] = None):
        import torch.distributed as dist

        self.ctx = Context.get(device)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rows = int(rows)
        self.group = group
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.jegal_qgather_create(self.ctx.h, self.rank, self.world, self.rows, C.byref(h)), "jegal_qgather_create")
        self.h = h
        mine = (C.c_uint8 * 64)()
        self.ctx.check(self.ctx.lib.jegal_qgather_ipc_handle(self.h, mine), "jegal_qgather_ipc_handle")
        handles = [bytes(mine)]
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine), group=group)
            handles = gathered
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self.ctx.check(self.ctx.lib.jegal_qgather_connect(self.h, buf), "jegal_qgather_connect")
        if self.world > 1:
            dist.barrier(group=group)

    def slice_rows(self):
        """This rank's share [r0, r1) of the rows (equal ceil-sized slices)."""
        per = (self.rows + self.world - 1) // self.world
        r0 = min(self.rank * per, self.rows)
        return r0, min(r0 + per, self.rows)

    def prep_gather(self, emb_slice: torch.Tensor, row0: int, normalize: bool = True, row_eps: float = 1e-12,
                    out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        if emb_slice.numel() and (not emb_slice.is_cuda or emb_slice.dim() != 2 or emb_slice.shape[1] != 512 or not emb_slice.is_contiguous()):
            raise JegalError("prep_gather: expected a contiguous CUDA [n, 512] slice")
        if emb_slice.dtype not in _DT or out_dtype not in (torch.bfloat16, torch.float16):
            raise JegalError("prep_gather: unsupported dtype")
        res = C.c_void_p()
        rc = self.ctx.lib.jegal_prep_gather(self.ctx.h, self.h, _ptr(emb_slice) if emb_slice.numel() else C.c_void_p(0), _DT[emb_slice.dtype],
                                            int(row0), int(emb_slice.shape[0]), int(bool(normalize)), float(row_eps), _DT[out_dtype],
                                            C.byref(res), _stream())
        self.ctx.check(rc, "jegal_prep_gather")
        dev = torch.device("cuda", self.ctx.device)
        t = torch.as_tensor(_DevArray(res.value, (self.rows, 512), "<i2"), device=dev)
        return t.view(out_dtype)

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                self.ctx.lib.jegal_qgather_destroy(self.h)
                self.h = None
        except Exception:
            pass
