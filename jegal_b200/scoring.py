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

_layout_cache: Dict[bytes, Layout] = {}


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
    """Layouts are cached by their length vector (they own small device tables)."""
    arr = np.ascontiguousarray(np.asarray(lengths, dtype=np.int32))
    key = arr.tobytes()
    lay = _layout_cache.get(key)
    if lay is None:
        if len(_layout_cache) > 64:
            _layout_cache.clear()
        lay = Layout.from_lengths(arr)
        _layout_cache[key] = lay
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
    _, _, gm = ops.prep(g.rows, g.layout, normalize=False, out_dtype=op_dtype, want_mean_rows=True)
    _, _, cm = ops.prep(c.rows, c.layout, normalize=False, out_dtype=op_dtype, want_mean_rows=True)
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
               want_full: bool = False, op_dtype: torch.dtype = torch.bfloat16) -> dict:
    """All clips of a spotting set in one launch.  Returns numpy arrays:
    heat (list of (T_i,) target-word rows), full (list of (W_i, T_i) matrices) if asked,
    pred_frame, pred_score and correct (if windows=(lo, hi) given)."""
    g, c = _pack(gestures), _pack(contents)
    if g.n != c.n:
        raise JegalError("spot_batch: gestures and contents must list the same clips")
    dev = g.rows.device
    g16, _ = ops.prep(g.rows, g.layout, normalize=normalize, out_dtype=op_dtype)
    c16, _ = ops.prep(c.rows, c.layout, normalize=normalize, out_dtype=op_dtype)
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
            gn16, cn16, gl_n, cl_n = g16, c16, g.layout, c.layout
        else:  # K3 takes clip i of both layouts: gather the rows of the narrow clips (one device gather each)
            keep = np.zeros(n, dtype=bool)
            keep[narrow] = True
            gn16 = g16[torch.from_numpy(np.repeat(keep, lt)).to(dev)]
            cn16 = c16[torch.from_numpy(np.repeat(keep, lw)).to(dev)]
            gl_n, cl_n = layout_for(lt[narrow]), layout_for(lw[narrow])
        wi = torch.as_tensor(word_idx[narrow], device=dev)
        lo = hi = None
        if windows is not None:
            lo, hi = torch.as_tensor(lo_h[narrow], device=dev), torch.as_tensor(hi_h[narrow], device=dev)
        r = ops.spot(gn16, gl_n, cn16, cl_n, wi, tau=temp, want_heat=True, want_full=want_full,
                     win_lo=lo, win_hi=hi, thresh=thresh)
        cu_n = gl_n.cu_len
        heat = r["heat"].cpu().numpy()
        pred_frame[narrow] = r["pred_frame"].cpu().numpy()
        pred_score[narrow] = r["pred_score"].cpu().numpy()
        if correct is not None:
            correct[narrow] = r["correct"].cpu().numpy().astype(bool)
        if want_full:
            full = r["full"].cpu().numpy()
            off = r["full_off"].cpu().numpy()
        for k_, i in enumerate(narrow):
            heat_l[i] = heat[cu_n[k_]:cu_n[k_ + 1]]
            if want_full:
                full_l[i] = full[off[k_]:off[k_ + 1]].reshape(int(lw[i]), int(lt[i]))
    for i in wide:
        # more words than the grouped kernel's 64 columns (a long transcript): the same arithmetic from K1's
        # plain-GEMM epilogue (frames x words cosines), the per-frame softmax over words and K2's first argmax
        T, W = int(lt[i]), int(lw[i])
        cos = ops.simpool_allpairs(g16[cu_t[i]:cu_t[i + 1]], layout_for(np.ones(T, dtype=np.int32)),
                                   c16[cu_w[i]:cu_w[i + 1]], layout_for(np.ones(W, dtype=np.int32)), "mean_mean")
        probs, _ = ops.group_softmax(cos.view(-1), T, W, tau=temp)  # [T, W]: softmax over words for each frame
        row = probs[:, int(word_idx[i])].contiguous()
        v, f = ops.topk(row.view(1, T), 1)
        heat_l[i] = row.cpu().numpy()
        pred_frame[i] = int(f[0, 0])
        pred_score[i] = float(v[0, 0])
        if correct is not None:
            correct[i] = bool(lo_h[i] <= pred_frame[i] <= hi_h[i] and pred_score[i] >= thresh)
        if want_full:
            full_l[i] = probs.t().contiguous().cpu().numpy()
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


def get_spotting_acc(data_rows, gesture_emb, content_emb, word_boundaries, thresh: float = 0.5,
                     frame_thresh: int = 9) -> float:
    """Drop-in for evaluation/evaluate_spotting.py:59-90 — same arguments, same printed line,
    same returned accuracy (%), but one kernel launch for the whole set."""
    n = len(gesture_emb)
    word_idx, lo, hi = [], [], []
    for idx in range(n):
        row = data_rows[idx]
        twb = row.target_word_boundary if hasattr(row, "target_word_boundary") else row["target_word_boundary"]
        twb = _parse_wb(twb)
        allwb = _parse_wb(word_boundaries[idx])
        word_idx.append(allwb.index(twb))  # first match, like list.index at :70
        lo.append(max(twb[1] - frame_thresh, 0))
        hi.append(twb[2] + frame_thresh)
    r = spot_batch(gesture_emb, content_emb, word_idx, windows=(lo, hi), thresh=thresh)
    correct = int(r["correct"].sum())
    accuracy = (correct / n) * 100
    print("Word Spotting Accuracy: {}".format(accuracy))
    return accuracy


# ------------------------------------------------------------------------------ ASD
def get_similarity_cos(query_emb, data_emb, temp: float = TEMP) -> np.ndarray:
    """Drop-in for evaluation/evaluate_asd.py:43-51: softmax over the P candidates of
    cos(query, candidate) / temp.  query_emb (1, 512), data_emb (P, 512) -> (P,) float32."""
    q = _pack(_as_2d(query_emb))
    d = _pack(_as_2d(data_emb))
    if q.n != 1:
        raise JegalError("get_similarity_cos: query_emb must be (1, 512)")
    # nn.CosineSimilarity clamps each norm at 1e-8 (evaluate_asd.py:45)
    q16, _ = ops.prep(q.rows, q.layout, normalize=True, row_eps=1e-8)
    d16, _ = ops.prep(d.rows, d.layout, normalize=True, row_eps=1e-8)
    P = d.n
    dev = q.rows.device
    r = ops.simpool_pairs(d16, d.layout, q16, q.layout, torch.arange(P, dtype=torch.int32, device=dev),
                          torch.zeros(P, dtype=torch.int32, device=dev), "mean_mean", group_size=P, tau=temp,
                          want_probs=True)
    return r["probs"].cpu().numpy()


def _pair_scores(g16, gl: Layout, c16, cl: Layout, pg: torch.Tensor, pc: torch.Tensor, pool: str,
                 gs: Optional[torch.Tensor], cs: Optional[torch.Tensor]) -> torch.Tensor:
    """Pooled scores of the listed pairs: K4 for pairs whose content clip has <= 64 words, K1 on the
    single pair (any clip lengths) for the rest."""
    pc_h = pc.cpu().numpy()
    wide = np.nonzero(cl.lengths[pc_h] > GROUPED_MAX_WORDS)[0] if pc_h.size else np.zeros(0, dtype=np.int64)
    if len(wide) == 0:
        return ops.simpool_pairs(g16, gl, c16, cl, pg, pc, pool, gscale=gs, cscale=cs)["scores"]
    pg_h = pg.cpu().numpy()
    scores = torch.empty((pg.numel(),), dtype=torch.float32, device=g16.device)
    narrow = np.setdiff1d(np.arange(pg.numel()), wide)
    if len(narrow):
        # K4 validates the whole content layout: hand it the <= 64-word clips only (one device gather)
        lw = cl.lengths
        keep = lw <= GROUPED_MAX_WORDS
        remap = np.cumsum(keep) - 1
        cn16 = c16[torch.from_numpy(np.repeat(keep, lw)).to(g16.device)]
        cs_n = None if cs is None else cs[torch.from_numpy(np.nonzero(keep)[0]).to(g16.device)].contiguous()
        sel = torch.from_numpy(narrow).to(g16.device)
        pc_n = torch.from_numpy(remap[pc_h[narrow]].astype(np.int32)).to(g16.device)
        scores[sel] = ops.simpool_pairs(g16, gl, cn16, layout_for(lw[keep]), pg[sel].contiguous(), pc_n, pool,
                                        gscale=gs, cscale=cs_n)["scores"]
    cu_t, cu_w = gl.cu_len, cl.cu_len
    for p in wide:
        gi, ci = int(pg_h[p]), int(pc_h[p])
        one = ops.simpool_allpairs(g16[cu_t[gi]:cu_t[gi + 1]], layout_for([cu_t[gi + 1] - cu_t[gi]]),
                                   c16[cu_w[ci]:cu_w[ci + 1]], layout_for([cu_w[ci + 1] - cu_w[ci]]), pool,
                                   gscale=None if gs is None else gs[gi:gi + 1],
                                   cscale=None if cs is None else cs[ci:ci + 1])
        scores[p] = one[0, 0]
    return scores


def asd_batch(contents, gesture_tracks, pair_gest: Sequence[int], pair_cont: Sequence[int], tracks: int,
              prefixes: Sequence[int] = (2, 4, 6), temp: float = TEMP, mode: str = "reference",
              op_dtype: torch.dtype = torch.bfloat16) -> dict:
    """Active-speaker scoring for many groups at once (evaluate_asd.py:54-127).

    ``gesture_tracks`` / ``contents``: clip lists (or PackedClips); candidate p of the flat
    pair list scores gesture clip pair_gest[p] against content clip pair_cont[p]; every
    `tracks` consecutive pairs form one group whose first entry is the true speaker.
    mode "reference": cosine of mean-pooled embeddings, as the reference; otherwise a
    pooling mode name applied to the frame x word tile.
    Returns dict(scores [n_groups, tracks], pred {P: int32 [n_groups]}, acc {P: float}).
    """
    g, c = _pack(gesture_tracks), _pack(contents)
    dev = g.rows.device
    pg = torch.as_tensor(np.asarray(pair_gest, dtype=np.int32), device=dev)
    pc = torch.as_tensor(np.asarray(pair_cont, dtype=np.int32), device=dev)
    if mode == "reference":
        g16, gs = ops.prep(g.rows, g.layout, normalize=False, out_dtype=op_dtype, want_mean_scale=True, mean_eps=1e-8)
        c16, cs = ops.prep(c.rows, c.layout, normalize=False, out_dtype=op_dtype, want_mean_scale=True, mean_eps=1e-8)
        pool = "mean_mean"
    else:
        g16, gs = ops.prep(g.rows, g.layout, out_dtype=op_dtype)
        c16, cs = ops.prep(c.rows, c.layout, out_dtype=op_dtype)
        pool = mode
    scores = _pair_scores(g16, g.layout, c16, c.layout, pg, pc, pool, gs, cs)
    n_groups = pg.numel() // tracks
    pred, acc = {}, {}
    for P in prefixes:
        if P > tracks:
            continue
        _, am = ops.group_softmax(scores, n_groups, P, stride=tracks, tau=temp, want_probs=False)
        pred[P] = am.cpu().numpy()
        acc[P] = float((pred[P] == 0).mean()) if n_groups else float("nan")
    return dict(scores=scores.cpu().numpy().reshape(n_groups, tracks), pred=pred, acc=acc)
