"""Host-side mirror of the reference's scoring functions, running on the B200 kernels.

The reference (Sindhu-Hegde/jegal) has no operator/plugin API for this path: its
callers use module-level functions.  The functions below keep their NAMES,
ARGUMENTS and RETURN TYPES so the evaluation scripts work by swapping one import:

    reference                                              here
    evaluation/evaluate_retrieval.py:38-48  get_similarity_matrix(emb1, emb2)
    evaluation/evaluate_retrieval.py:51-65  compute_metrics(x)
    evaluation/evaluate_retrieval.py:68-84  print_computed_metrics / get_metrics
    evaluation/evaluate_spotting.py:39-57   get_attn_matrix(idx, gesture_emb, content_emb, word_boundaries, temp)
    utils/plot_heatmap.py:34-59             get_attn_matrix(gesture_emb, content_emb, word_boundaries, temp)
    evaluation/evaluate_spotting.py:59-90   get_spotting_acc(data_rows, gesture_emb, content_emb, word_boundaries, ...)
    evaluation/evaluate_asd.py:43-51        get_similarity_cos(query_emb, data_emb, temp)

plus the batched forms the kernels are built for (``score_allpairs``,
``retrieve_topk``, ``retrieval_metrics``, ``spot_batch``, ``asd_batch``): one
call for a whole directory of clips instead of a Python loop per clip.

Every function computes on the GPU through libjegal_b200.so and raises
``JegalError`` if that is impossible; none of them has a CPU implementation.
"""
from __future__ import annotations

import ast
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from .ops import JegalError, Layout

TEMP = 0.07  # evaluate_spotting.py:39, evaluate_asd.py:43, plot_heatmap.py:34
GROUPED_MAX_WORDS = 64  # K3 / K4 keep a clip's words on <= 64 accumulator columns; wider clips take the K1 route
ArrayLike = Union[np.ndarray, torch.Tensor, Sequence]

_layout_cache: Dict[Tuple[int, bytes], Layout] = {}


class _Staging:
    """One grow-only pinned host buffer for packing clip lists (cudaHostAlloc of gigabytes per call
    costs more than the copy it speeds up).  Reuse waits for the previous H2D copy out of it."""

    buf: Optional[torch.Tensor] = None
    event: Optional[torch.cuda.Event] = None

    @classmethod
    def get(cls, nbytes: int) -> torch.Tensor:
        if cls.event is not None:
            cls.event.synchronize()
        if cls.buf is None or cls.buf.numel() < nbytes:
            cls.buf = None
            cls.buf = torch.empty((max(nbytes, 1 << 20),), dtype=torch.uint8, pin_memory=True)
        return cls.buf[:nbytes]

    @classmethod
    def mark_copy(cls) -> None:
        cls.event = torch.cuda.Event()
        cls.event.record()


def pack_rows_host(arrs: Sequence[np.ndarray], out: np.ndarray, threads: int = 0, on_group=None) -> None:
    """Concatenate (len_i, 512) arrays into the preallocated `out` with several host threads (numpy's
    memcpy releases the GIL): packing the 4.7 GB of config 4 with one thread took longer than everything
    the GPU does with it.  `on_group(r0, r1)` is called from the calling thread, in row order, as soon as
    the rows [r0, r1) are in place, so the caller can start their H2D copy while the rest is being packed."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    n = len(arrs)
    lengths = np.fromiter((a.shape[0] for a in arrs), dtype=np.int64, count=n)
    cu = np.concatenate([[0], np.cumsum(lengths)])
    total_bytes = int(cu[-1]) * out.shape[1] * out.itemsize
    threads = threads or min(16, os.cpu_count() or 1)
    if n == 0:
        return
    if total_bytes < (32 << 20) or threads <= 1 or n < 2 * threads:
        np.concatenate(arrs, axis=0, out=out[: cu[-1]], casting="same_kind")
        if on_group is not None:
            on_group(0, int(cu[-1]))
        return
    n_groups = min(n, 4 * threads)
    targets = (np.arange(1, n_groups) * cu[-1]) // n_groups
    cuts = np.unique(np.concatenate([[0], np.searchsorted(cu, targets), [n]]))

    def work(g):
        lo, hi = int(cuts[g]), int(cuts[g + 1])
        np.concatenate(arrs[lo:hi], axis=0, out=out[cu[lo]:cu[hi]], casting="same_kind")
        return int(cu[lo]), int(cu[hi])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        for r0, r1 in ex.map(work, range(len(cuts) - 1)):  # results arrive in group order
            if on_group is not None and r1 > r0:
                on_group(r0, r1)


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise JegalError("jegal_b200.scoring needs a CUDA (sm_100) device; there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def layout_for(lengths: Sequence[int]) -> Layout:
    """Layouts are cached by (device, length vector): they own small device tables and a context bound to
    the device that was current when they were made.  Least recently used entries go first."""
    arr = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
    key = (torch.cuda.current_device() if torch.cuda.is_available() else -1, arr.tobytes())
    lay = _layout_cache.pop(key, None)
    if lay is None:
        while len(_layout_cache) >= 64:
            _layout_cache.pop(next(iter(_layout_cache)))
        lay = Layout.from_lengths(arr)
    _layout_cache[key] = lay  # (re)insert at the young end
    return lay


def _as_2d(x) -> np.ndarray:
    a = x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2 or a.shape[1] != 512:
        raise JegalError(f"expected a (rows, 512) embedding, got {a.shape}")
    if a.dtype not in (np.float16, np.float32):
        a = a.astype(np.float32)
    return a


class PackedClips:
    """A list of (len_i, 512) embeddings packed into one device matrix + ragged layout."""

    def __init__(self, rows: torch.Tensor, layout: Layout):
        self.rows = rows
        self.layout = layout

    @classmethod
    def from_list(cls, clips: Sequence[ArrayLike], pin: bool = True) -> "PackedClips":
        dev = _device()
        if isinstance(clips, torch.Tensor) and clips.dim() == 2:  # N single-row clips
            clips = clips.detach().cpu().numpy()
        if isinstance(clips, np.ndarray) and clips.ndim == 2:
            arrs, lengths = None, np.ones(clips.shape[0], dtype=np.int32)
            host = _as_2d(clips)
        else:
            arrs = [_as_2d(c) for c in clips]
            lengths = np.array([a.shape[0] for a in arrs], dtype=np.int32)
            dt = np.float16 if all(a.dtype == np.float16 for a in arrs) else np.float32
            host = None
        total = int(lengths.sum())
        tdt = torch.float16 if (dt if arrs is not None else host.dtype) == np.float16 else torch.float32
        if total == 0:
            return cls(torch.empty((0, 512), dtype=tdt, device=dev), layout_for(lengths))
        if pin:
            buf = _Staging.get(total * 512 * (2 if tdt == torch.float16 else 4)).view(tdt).view(total, 512)
        else:
            buf = torch.empty((total, 512), dtype=tdt)
        if arrs is not None and pin:
            # packed by several host threads; the H2D copy of a group starts as soon as it is in place
            rows = torch.empty((total, 512), dtype=tdt, device=dev)
            pack_rows_host(arrs, buf.numpy(), on_group=lambda r0, r1: rows[r0:r1].copy_(buf[r0:r1], non_blocking=True))
        else:
            if arrs is not None:
                pack_rows_host(arrs, buf.numpy())
            else:
                np.copyto(buf.numpy(), host, casting="same_kind")
            rows = buf.to(dev, non_blocking=True)
        if pin:
            _Staging.mark_copy()
        return cls(rows, layout_for(lengths))

    @classmethod
    def from_packed(cls, rows: torch.Tensor, cu_len: np.ndarray) -> "PackedClips":
        return cls(rows.to(_device()), layout_for(np.diff(np.asarray(cu_len, dtype=np.int64))))

    @property
    def n(self) -> int:
        return self.layout.n_clips


def _pack(x) -> PackedClips:
    return x if isinstance(x, PackedClips) else PackedClips.from_list(x)


_NATIVE16 = (torch.float16, torch.bfloat16)


def _stored_operands(x: PackedClips, normalize: bool, op_dtype: Optional[torch.dtype], fuse: bool = True):
    """Operand rows for the grouped kernels (K3 / K4) and whether the KERNEL has to normalise them.

    16-bit rows (what the reference's .pkl files hold) go to the kernel exactly as stored: the L2
    normalisation is fused into the operand load, so no normalised copy is written to and re-read from HBM.
    fp32 rows (or an explicit other ``op_dtype``, or ``fuse=False``) take one K0 pass: normalise + cast."""
    if fuse and x.rows.dtype in _NATIVE16 and op_dtype in (None, x.rows.dtype):
        return x.rows, bool(normalize)
    od = op_dtype or (x.rows.dtype if x.rows.dtype in _NATIVE16 else torch.bfloat16)
    if not normalize and x.rows.dtype == od:
        return x.rows, False  # nothing to do: rows are used as stored
    rows16, _ = ops.prep(x.rows, x.layout, normalize=normalize, out_dtype=od)
    return rows16, False


# ------------------------------------------------------------------------------ retrieval
def get_similarity_matrix(emb1, emb2) -> torch.Tensor:
    """Drop-in for evaluation/evaluate_retrieval.py:38-48.

    ``emb1`` / ``emb2``: list of (512,) arrays, (N, 512) ndarray or tensor (clip-level
    embeddings).  Returns the (N1, N2) cosine matrix as a CPU FloatTensor, row = query.
    Rows are normalised (F.normalize, eps 1e-12) by K0 and the contraction runs on the
    tensor cores (K1 with one-row clips).
    """
    a = _pack(np.stack([np.asarray(e).reshape(-1) for e in emb1]) if isinstance(emb1, (list, tuple)) else emb1)
    b = _pack(np.stack([np.asarray(e).reshape(-1) for e in emb2]) if isinstance(emb2, (list, tuple)) else emb2)
    a16, _ = ops.prep(a.rows, a.layout, normalize=True)
    b16, _ = ops.prep(b.rows, b.layout, normalize=True)
    s = ops.simpool_allpairs(a16, a.layout, b16, b.layout, "mean_mean")
    return s.cpu()


def _metrics_from_counts(n_greater: np.ndarray, n_equal: np.ndarray) -> dict:
    """compute_metrics' `ind` (evaluate_retrieval.py:52-57) lists, for every row, each sorted
    position whose value equals the diagonal: positions n_greater .. n_greater + n_equal - 1."""
    n_greater = n_greater.astype(np.int64)
    n_equal = n_equal.astype(np.int64)
    if n_equal.size == 0:
        raise JegalError("compute_metrics: empty similarity matrix")
    if np.any(n_equal < 1):  # x[i, i] != x[i, i]: a NaN diagonal (zero-length clip or NaN embedding)
        bad = np.flatnonzero(n_equal < 1)
        raise JegalError(f"compute_metrics: the ground-truth score of row(s) {bad[:8].tolist()} is NaN")
    total = int(n_equal.sum())
    m = {}
    for k in (1, 5, 10, 25, 50):
        hits = np.clip(k - n_greater, 0, n_equal)  # positions < k inside each row's run
        m[f"R{k}"] = float(hits.sum()) / total
    if np.all(n_equal == 1):
        ind = n_greater
    else:
        ind = np.repeat(n_greater, n_equal) + (np.arange(total) - np.repeat(np.cumsum(n_equal) - n_equal, n_equal))
    m["MR"] = float(np.median(ind) + 1)
    return m


def compute_metrics(x) -> dict:
    """Drop-in for evaluation/evaluate_retrieval.py:51-65 (R5, R10, R25, R50, MR; plus R1).

    ``x``: (N, N) similarity matrix (tensor or ndarray, CPU or CUDA), ground truth on the
    diagonal.  The O(N^2 log N) sort is replaced by a rank-of-positive count on the GPU;
    ties are counted the way the reference counts them.
    """
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.asarray(x, dtype=np.float32))
    t = t.to(device=_device(), dtype=torch.float32)
    ngt, neq = ops.rank_of_positive(t)
    return _metrics_from_counts(ngt.cpu().numpy(), neq.cpu().numpy())


def print_computed_metrics(metrics: dict) -> None:
    """evaluation/evaluate_retrieval.py:68-74 (same line format)."""
    print('R@5: {:.2f} - R@10: {:.2f} - R@25: {:.2f} - R@50: {:.2f} | Median R: {:.1f}'.format(
        metrics['R5'] * 100, metrics['R10'] * 100, metrics['R25'] * 100, metrics['R50'] * 100, metrics['MR']))


def get_metrics(emb1, emb2) -> dict:
    """evaluation/evaluate_retrieval.py:76-84 (also returns the metrics)."""
    metrics = compute_metrics(get_similarity_matrix(emb1, emb2))
    print_computed_metrics(metrics)
    return metrics


def score_allpairs(gestures, contents, mode: str = "mean_mean", refnorm: bool = False,
                   normalize_rows: bool = True, op_dtype: torch.dtype = torch.bfloat16,
                   content_major: bool = False, device_out: bool = False):
    """Pooled frame x word cosine score of every (gesture clip, content clip) pair.

    ``gestures`` / ``contents``: lists of (T_i, 512) / (W_j, 512) arrays or PackedClips.
    ``refnorm=True`` with mode "mean_mean" reproduces the reference's retrieval score
    (cosine of the mean-pooled clips, evaluate_retrieval.py:30-31,38-48): rows are used as
    stored and the pooled tile is scaled by 1/(||mean_g|| ||mean_c||).
    """
    g, c = _pack(gestures), _pack(contents)
    if refnorm:
        if mode != "mean_mean":
            raise JegalError("refnorm applies to mode='mean_mean' only")
        if g.rows.dtype == op_dtype and c.rows.dtype == op_dtype:  # stored rows are the operands: statistics only
            (g16, c16) = (g.rows, c.rows)
            gs, cs = ops.clip_means(g.rows, g.layout, want_rows=False, want_scale=True)[1], \
                ops.clip_means(c.rows, c.layout, want_rows=False, want_scale=True)[1]
        else:
            g16, gs = ops.prep(g.rows, g.layout, normalize=False, out_dtype=op_dtype, want_mean_scale=True)
            c16, cs = ops.prep(c.rows, c.layout, normalize=False, out_dtype=op_dtype, want_mean_scale=True)
    else:
        g16, gs = ops.prep(g.rows, g.layout, normalize=normalize_rows, out_dtype=op_dtype)
        c16, cs = ops.prep(c.rows, c.layout, normalize=normalize_rows, out_dtype=op_dtype)
    s = ops.simpool_allpairs(g16, g.layout, c16, c.layout, mode, gscale=gs, cscale=cs, content_major=content_major)
    return s if device_out else s.cpu().numpy()


def clip_similarity_matrix(gestures, contents, op_dtype: torch.dtype = torch.bfloat16, device_out: bool = False):
    """load_feats' temporal mean (evaluate_retrieval.py:30-31) + get_similarity_matrix in
    two kernels: K0 emits the unit-norm mean row of every clip, K1 contracts them.
    Returns (n_gest, n_cont): cos(mean_g, mean_c)."""
    g, c = _pack(gestures), _pack(contents)
    # read-only pass: the unit-norm mean rows are all that is written (n_clips x 1 KB)
    gm, _ = ops.clip_means(g.rows, g.layout, out_dtype=op_dtype)
    cm, _ = ops.clip_means(c.rows, c.layout, out_dtype=op_dtype)
    lg, lc = layout_for(np.ones(g.n, dtype=np.int32)), layout_for(np.ones(c.n, dtype=np.int32))
    s = ops.simpool_allpairs(gm, lg, cm, lc, "mean_mean")
    return s if device_out else s.cpu().numpy()


def retrieval_metrics(gestures, contents, mode: str = "reference") -> Tuple[dict, dict]:
    """Both directions of evaluate_retrieval.py:87-96 in one pass: (c2g, g2c) metric dicts.
    mode "reference" scores clips like the reference (cosine of mean-pooled embeddings);
    any pooling mode name scores the frame x word tiles instead."""
    if mode == "reference":
        s = clip_similarity_matrix(gestures, contents, device_out=True)
    else:
        s = score_allpairs(gestures, contents, mode, device_out=True)
    g2c = _metrics_from_counts(*[t.cpu().numpy() for t in ops.rank_of_positive(s)])
    c2g = _metrics_from_counts(*[t.cpu().numpy() for t in ops.rank_of_positive(s.t())])
    return c2g, g2c


def retrieve_topk(queries, gallery, k: int = 10, mode: str = "max_t_mean_w", queries_are: str = "gesture",
                  op_dtype: torch.dtype = torch.bfloat16) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k gallery clips for every query clip: (scores [Q, k], indices [Q, k])."""
    q, g = _pack(queries), _pack(gallery)
    q16, _ = ops.prep(q.rows, q.layout, out_dtype=op_dtype)
    g16, _ = ops.prep(g.rows, g.layout, out_dtype=op_dtype)
    if queries_are == "gesture":
        s = ops.simpool_allpairs(q16, q.layout, g16, g.layout, mode)
    else:
        s = ops.simpool_allpairs(g16, g.layout, q16, q.layout, mode, content_major=True)
    v, i = ops.topk(s, k)
    return v.cpu().numpy(), i.cpu().numpy()


# ------------------------------------------------------------------------------ spotting
def _parse_wb(wb):
    return ast.literal_eval(wb) if isinstance(wb, str) else wb


def spot_batch(gestures, contents, word_idx: Sequence[int], temp: float = TEMP, normalize: bool = True,
               windows: Optional[Tuple[Sequence[int], Sequence[int]]] = None, thresh: float = 0.5,
               want_full: bool = False, op_dtype: Optional[torch.dtype] = None, fuse: bool = True,
               want_heat: bool = True) -> dict:
    """All clips of a spotting set in one launch.  Returns numpy arrays:
    heat (list of (T_i,) target-word rows), full (list of (W_i, T_i) matrices) if asked,
    pred_frame, pred_score and correct (if windows=(lo, hi) given).

    fp16 (or bf16) embeddings are fed to K3 as stored and normalised inside the kernel; ``op_dtype`` /
    ``fuse=False`` force the K0 pass (normalise + cast to ``op_dtype``) in front of it instead."""
    g, c = _pack(gestures), _pack(contents)
    if g.n != c.n:
        raise JegalError("spot_batch: gestures and contents must list the same clips")
    dev = g.rows.device
    word_idx = np.asarray(word_idx, dtype=np.int32)
    lo_h = hi_h = None
    if windows is not None:
        lo_h, hi_h = np.asarray(windows[0], dtype=np.int32), np.asarray(windows[1], dtype=np.int32)
    lt, lw = g.layout.lengths, c.layout.lengths
    cu_t, cu_w = g.layout.cu_len, c.layout.cu_len
    wide = np.nonzero(lw > GROUPED_MAX_WORDS)[0]
    n = g.n
    heat_l: List[Optional[np.ndarray]] = [None] * n
    full_l: List[Optional[np.ndarray]] = [None] * n
    pred_frame = np.zeros(n, dtype=np.int32)
    pred_score = np.zeros(n, dtype=np.float32)
    correct = np.zeros(n, dtype=bool) if windows is not None else None

    narrow = np.nonzero(lw <= GROUPED_MAX_WORDS)[0]
    if len(narrow):
        if len(wide) == 0:
            gp, cp = g, c
        else:  # K3 takes clip i of both layouts: gather the rows of the narrow clips (one device gather each)
            keep = np.zeros(n, dtype=bool)
            keep[narrow] = True
            gp = PackedClips(g.rows[torch.from_numpy(np.repeat(keep, lt)).to(dev)], layout_for(lt[narrow]))
            cp = PackedClips(c.rows[torch.from_numpy(np.repeat(keep, lw)).to(dev)], layout_for(lw[narrow]))
        g_op, kn_g = _stored_operands(gp, normalize, op_dtype, fuse)
        c_op, kn_c = _stored_operands(cp, normalize, op_dtype, fuse)
        if kn_g != kn_c or g_op.dtype != c_op.dtype:  # mixed storage types: normalise + cast both with K0
            od = op_dtype or torch.bfloat16
            g_op = ops.prep(gp.rows, gp.layout, normalize=normalize, out_dtype=od)[0]
            c_op = ops.prep(cp.rows, cp.layout, normalize=normalize, out_dtype=od)[0]
            kn_g = False
        wi = torch.as_tensor(word_idx[narrow], device=dev)
        lo = hi = None
        if windows is not None:
            lo, hi = torch.as_tensor(lo_h[narrow], device=dev), torch.as_tensor(hi_h[narrow], device=dev)
        r = ops.spot(g_op, gp.layout, c_op, cp.layout, wi, tau=temp, want_heat=want_heat or not want_full,
                     want_full=want_full, win_lo=lo, win_hi=hi, thresh=thresh, normalize=kn_g)
        cu_n = gp.layout.cu_len
        pred_frame[narrow] = r["pred_frame"].cpu().numpy()
        pred_score[narrow] = r["pred_score"].cpu().numpy()
        if correct is not None:
            correct[narrow] = r["correct"].cpu().numpy().astype(bool)
        if r["heat"] is not None:
            heat = r["heat"].cpu().numpy()
            for k_, i in enumerate(narrow):
                heat_l[i] = heat[cu_n[k_]:cu_n[k_ + 1]]
        if want_full:
            full = r["full"].cpu().numpy()
            off = r["full_off"].cpu().numpy()
            for k_, i in enumerate(narrow):
                full_l[i] = full[off[k_]:off[k_ + 1]].reshape(int(lw[i]), int(lt[i]))
    if len(wide):
        # More words than the grouped kernel's 64 columns (long transcripts): the same arithmetic from K1's plain-GEMM
        # epilogue over the packed frames x words of a GROUP of such clips (one launch; the cross-clip blocks are
        # wasted tensor time, bounded by the group size), then the per-frame softmax over the clip's words + argmax
        # + decision in two small kernels (ops.spot_dense).  One launch set per group, not per clip.
        budget = 32 << 20  # floats of the group's dense cosine matrix (128 MB)
        groups, cur, st_, sw_ = [], [], 0, 0
        for i in wide:
            T, W = int(lt[i]), int(lw[i])
            if cur and (st_ + T) * (sw_ + W) > budget:
                groups.append(cur)
                cur, st_, sw_ = [], 0, 0
            cur.append(int(i))
            st_, sw_ = st_ + T, sw_ + W
        groups.append(cur)
        for grp in groups:
            keep = np.zeros(n, dtype=bool)
            keep[grp] = True
            g_rows = g.rows[torch.from_numpy(np.repeat(keep, lt)).to(dev)]
            c_rows = c.rows[torch.from_numpy(np.repeat(keep, lw)).to(dev)]
            gl_w, cl_w = layout_for(lt[grp]), layout_for(lw[grp])
            od = op_dtype or (g_rows.dtype if (g_rows.dtype in _NATIVE16 and c_rows.dtype == g_rows.dtype) else torch.bfloat16)
            g16 = g_rows if (not normalize and g_rows.dtype == od) else ops.prep(g_rows, gl_w, normalize=normalize, out_dtype=od)[0]
            c16 = c_rows if (not normalize and c_rows.dtype == od) else ops.prep(c_rows, cl_w, normalize=normalize, out_dtype=od)[0]
            cos = ops.simpool_allpairs(g16, layout_for(np.ones(gl_w.rows, dtype=np.int32)),
                                       c16, layout_for(np.ones(cl_w.rows, dtype=np.int32)), "mean_mean")
            lo = hi = None
            if windows is not None:
                lo, hi = torch.as_tensor(lo_h[grp], device=dev), torch.as_tensor(hi_h[grp], device=dev)
            r = ops.spot_dense(cos, gl_w, cl_w, torch.as_tensor(word_idx[grp], device=dev), tau=temp, want_full=want_full,
                               win_lo=lo, win_hi=hi, thresh=thresh)
            pred_frame[grp] = r["pred_frame"].cpu().numpy()
            pred_score[grp] = r["pred_score"].cpu().numpy()
            if correct is not None:
                correct[grp] = r["correct"].cpu().numpy().astype(bool)
            heat = r["heat"].cpu().numpy()
            cu_g = gl_w.cu_len
            if want_full:
                full = r["full"].cpu().numpy()
                off = r["full_off"].cpu().numpy()
            for k_, i in enumerate(grp):
                heat_l[i] = heat[cu_g[k_]:cu_g[k_ + 1]]
                if want_full:
                    full_l[i] = full[off[k_]:off[k_ + 1]].reshape(int(lw[i]), int(lt[i]))
    out = dict(heat=heat_l, pred_frame=pred_frame, pred_score=pred_score, correct=correct)
    if want_full:
        out["full"] = full_l
    return out


def get_attn_matrix(*args, temp: float = TEMP):
    """Drop-in for BOTH reference signatures:

      evaluate_spotting.py:39  get_attn_matrix(idx, gesture_emb, content_emb, word_boundaries, temp=0.07)
                               (lists indexed by idx; rows are re-normalised, :49-50)
      plot_heatmap.py:34       get_attn_matrix(gesture_emb, content_emb, word_boundaries, temp=0.07)
                               (one clip; rows used as stored)

    Returns (attn (W, T) float32 ndarray, all_words list) like the reference.
    """
    if len(args) >= 4 and isinstance(args[0], (int, np.integer)):
        idx, gesture_emb, content_emb, word_boundaries = args[:4]
        if len(args) > 4:
            temp = args[4]
        gesture, content, wb, normalize = gesture_emb[idx], content_emb[idx], word_boundaries[idx], True
    elif len(args) >= 3:
        gesture, content, wb = args[:3]
        if len(args) > 3:
            temp = args[3]
        normalize = False
    else:
        raise TypeError("get_attn_matrix: expected (idx, gesture_emb, content_emb, word_boundaries[, temp]) "
                        "or (gesture_emb, content_emb, word_boundaries[, temp])")
    wb = _parse_wb(wb)
    all_words = [wb[i][0] for i in range(len(wb))]
    r = spot_batch([gesture], [content], [0], temp=temp, normalize=normalize, want_full=True)
    return r["full"][0], all_words


def spot_targets(data_rows, word_boundaries, frame_thresh: int = 9):
    """Per clip: index of the target word (first match, like list.index at evaluate_spotting.py:70) and the
    accepted frame window [max(start - frame_thresh, 0), end + frame_thresh] (:75-78)."""
    word_idx, lo, hi = [], [], []
    for idx in range(len(word_boundaries)):
        row = data_rows[idx]
        twb = row.target_word_boundary if hasattr(row, "target_word_boundary") else row["target_word_boundary"]
        twb = _parse_wb(twb)
        allwb = _parse_wb(word_boundaries[idx])
        word_idx.append(allwb.index(twb))
        lo.append(max(twb[1] - frame_thresh, 0))
        hi.append(twb[2] + frame_thresh)
    return word_idx, lo, hi


def get_spotting_acc(data_rows, gesture_emb, content_emb, word_boundaries, thresh: float = 0.5,
                     frame_thresh: int = 9) -> float:
    """Drop-in for evaluation/evaluate_spotting.py:59-90 — same arguments, same printed line,
    same returned accuracy (%), but one kernel launch for the whole set."""
    n = len(gesture_emb)
    word_idx, lo, hi = spot_targets(data_rows, word_boundaries, frame_thresh)
    r = spot_batch(gesture_emb, content_emb, word_idx, windows=(lo, hi), thresh=thresh)
    correct = int(r["correct"].sum())
    accuracy = (correct / n) * 100
    print("Word Spotting Accuracy: {}".format(accuracy))
    return accuracy


# ------------------------------------------------------------------------------ ASD
def get_similarity_cos(query_emb, data_emb, temp: float = TEMP) -> np.ndarray:
    """Drop-in for evaluation/evaluate_asd.py:43-51: softmax over the P candidates of
    cos(query, candidate) / temp.  query_emb (1, 512), data_emb (P, 512) -> (P,) float32."""
    qa, da = _as_2d(query_emb), _as_2d(data_emb)
    if qa.shape[0] != 1:
        raise JegalError("get_similarity_cos: query_emb must be (1, 512)")
    if qa.dtype != da.dtype:
        qa, da = qa.astype(np.float32), da.astype(np.float32)
    dev = _device()
    q = torch.from_numpy(np.ascontiguousarray(qa)).to(dev)
    d = torch.from_numpy(np.ascontiguousarray(da)).to(dev)
    P = d.shape[0]
    # nn.CosineSimilarity clamps each norm at 1e-8 (evaluate_asd.py:45): one warp per candidate reads both rows
    cos = ops.pair_cosine(d, q, None, torch.zeros(P, dtype=torch.int32, device=dev), normalize=True, eps=1e-8)
    probs, _ = ops.group_softmax(cos, 1, P, tau=temp)
    return probs.view(-1).cpu().numpy()


def _pair_scores(g: PackedClips, c: PackedClips, pg: torch.Tensor, pc: torch.Tensor, pool: str,
                 op_dtype: Optional[torch.dtype] = None, fuse: bool = True) -> torch.Tensor:
    """Pooled frame x word cosine scores of the listed pairs: K4 (row normalisation fused into the load for
    16-bit stored rows) for content clips of <= 64 words, K1 on the single pair (any clip lengths) for the rest."""
    gl, cl = g.layout, c.layout
    dev = g.rows.device
    pc_h = pc.cpu().numpy()
    # K4 validates the WHOLE content layout (an unreferenced > 64-word clip would make it refuse the call),
    # so the filtered route below is taken whenever the layout holds such a clip at all
    if pc_h.size == 0 or cl.n_clips == 0 or int(cl.lengths.max()) <= GROUPED_MAX_WORDS:
        g_op, kn = _stored_operands(g, True, op_dtype, fuse)
        c_op, kn_c = _stored_operands(c, True, op_dtype, fuse)
        if kn != kn_c or g_op.dtype != c_op.dtype:
            od = op_dtype or torch.bfloat16
            g_op, c_op, kn = ops.prep(g.rows, gl, out_dtype=od)[0], ops.prep(c.rows, cl, out_dtype=od)[0], False
        return ops.simpool_pairs(g_op, gl, c_op, cl, pg, pc, pool, normalize=kn)["scores"]
    wide = np.nonzero(cl.lengths[pc_h] > GROUPED_MAX_WORDS)[0]
    pg_h = pg.cpu().numpy()
    scores = torch.empty((pg.numel(),), dtype=torch.float32, device=dev)
    narrow = np.setdiff1d(np.arange(pg.numel()), wide)
    if len(narrow):
        lw = cl.lengths
        keep = lw <= GROUPED_MAX_WORDS
        remap = np.cumsum(keep) - 1
        cn = PackedClips(c.rows[torch.from_numpy(np.repeat(keep, lw)).to(dev)], layout_for(lw[keep]))
        sel = torch.from_numpy(narrow).to(dev)
        pc_n = torch.from_numpy(remap[pc_h[narrow]].astype(np.int32)).to(dev)
        scores[sel] = _pair_scores(g, cn, pg[sel].contiguous(), pc_n, pool, op_dtype, fuse)
    # pairs whose content clip has more than 64 words: ONE all-pairs K1 launch over the gesture clips and the wide
    # content clips these pairs mention (K1 pools any clip lengths), then a gather of the listed entries
    g_ids, g_inv = np.unique(pg_h[wide], return_inverse=True)
    c_ids, c_inv = np.unique(pc_h[wide], return_inverse=True)
    lt = gl.lengths
    keep_g = np.zeros(gl.n_clips, dtype=bool)
    keep_g[g_ids] = True
    keep_c = np.zeros(cl.n_clips, dtype=bool)
    keep_c[c_ids] = True
    gw = PackedClips(g.rows[torch.from_numpy(np.repeat(keep_g, lt)).to(dev)], layout_for(lt[g_ids]))
    cw = PackedClips(c.rows[torch.from_numpy(np.repeat(keep_c, cl.lengths)).to(dev)], layout_for(cl.lengths[c_ids]))
    od = op_dtype or (gw.rows.dtype if (gw.rows.dtype in _NATIVE16 and cw.rows.dtype == gw.rows.dtype) else torch.bfloat16)
    g16 = ops.prep(gw.rows, gw.layout, normalize=True, out_dtype=od)[0]
    c16 = ops.prep(cw.rows, cw.layout, normalize=True, out_dtype=od)[0]
    sw = ops.simpool_allpairs(g16, gw.layout, c16, cw.layout, pool)
    scores[torch.from_numpy(wide).to(dev)] = sw[torch.from_numpy(g_inv).to(dev), torch.from_numpy(c_inv).to(dev)]
    return scores


def asd_batch(contents, gesture_tracks, pair_gest: Sequence[int], pair_cont: Sequence[int], tracks: int,
              prefixes: Sequence[int] = (2, 4, 6), temp: float = TEMP, mode: str = "reference",
              op_dtype: Optional[torch.dtype] = None, fuse: bool = True) -> dict:
    """Active-speaker scoring for many groups at once (evaluate_asd.py:54-127).

    ``gesture_tracks`` / ``contents``: clip lists (or PackedClips); candidate p of the flat
    pair list scores gesture clip pair_gest[p] against content clip pair_cont[p]; every
    `tracks` consecutive pairs form one group whose first entry is the true speaker.
    mode "reference": cosine of the mean-pooled embeddings, as the reference (one read-only pass over the
    stored rows yields the clip means, one warp per pair their cosine); otherwise a pooling mode name applied
    to the frame x word cosine tile (K4, row normalisation fused into the operand load).
    Returns dict(scores [n_groups, tracks], pred {P: int32 [n_groups]}, acc {P: float}).
    """
    g, c = _pack(gesture_tracks), _pack(contents)
    dev = g.rows.device
    pg = torch.as_tensor(np.asarray(pair_gest, dtype=np.int32), device=dev)
    pc = torch.as_tensor(np.asarray(pair_cont, dtype=np.int32), device=dev)
    if mode == "reference":
        # load_feats' emb.mean(axis=0) (evaluate_asd.py:31-36) + CosineSimilarity(eps=1e-8) (:45-47):
        # cos = m_g . m_c / (max(||m_g||, eps) max(||m_c||, eps)) = dot of the two clamped-unit mean rows
        gm, _ = ops.clip_means(g.rows, g.layout, mean_eps=1e-8)
        cm, _ = ops.clip_means(c.rows, c.layout, mean_eps=1e-8)
        scores = ops.pair_cosine(gm, cm, pg, pc, normalize=False)
    else:
        scores = _pair_scores(g, c, pg, pc, mode, op_dtype, fuse)
    n_groups = pg.numel() // tracks
    pred, acc = {}, {}
    for P in prefixes:
        if P > tracks:
            continue
        _, am = ops.group_softmax(scores, n_groups, P, stride=tracks, tau=temp, want_probs=False)
        pred[P] = am.cpu().numpy()
        acc[P] = float((pred[P] == 0).mean()) if n_groups else float("nan")
    return dict(scores=scores.cpu().numpy().reshape(n_groups, tracks), pred=pred, acc=acc)
