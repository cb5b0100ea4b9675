"""CPU oracle for the JEGAL cross-modal scoring path — TEST INFRASTRUCTURE ONLY.

A plain fp32 torch/numpy restatement of the reference's scoring arithmetic
(Sindhu-Hegde/jegal, paths relative to the reference root).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module; nothing under ``jegal_b200/`` does, and
the product path has no CPU fallback.

Pinning status
  * get_similarity_matrix / compute_metrics / get_attn_matrix / spot decision /
    get_similarity_cos / the ASD argmax loop: PINNED against the reference's own
    functions executed in the build container (oracle/make_golden.py imports
    evaluation/evaluate_{retrieval,spotting,asd}.py and stores inputs + outputs in
    tests/golden/*.npz; tests/test_oracle_golden.py replays them here).
    utils/plot_heatmap.py cannot be imported (matplotlib is absent); its
    get_attn_matrix (:34-59) equals evaluate_spotting's minus the F.normalize and
    is pinned through that one.
  * simpool_* pooling modes: the T x W cosine tile is PINNED — tests/golden/simpool_tiles.npz
    holds tiles produced by the reference's own get_similarity_matrix
    (evaluate_retrieval.py:38-48: F.normalize + matmul) on per-frame / per-word rows of clip
    pairs, plus numpy max / mean of those tiles; cos_tile, pool_tile and both simpool_allpairs
    forms replay them (tests/test_oracle_golden.py).  Only the final amax / mean over a
    reference-made tile and the top-k order (stable argsort) are restated: the released code
    has no max-pool scoring and no top-k (training loss unreleased, README.md:163-165), so
    those two one-liners stay PARITY UNPINNED.  The mean/mean mode is also pinned through the
    identity with get_similarity_matrix on the mean-pooled clips.
  * word_level_* (models/jegal.py:131-252): PINNED — oracle/make_golden.py extracts the two
    methods' source from models/jegal.py with `ast` (the module itself cannot be imported
    offline: it downloads XLM-R at import, models/jegal.py:13-14) and executes them on synthetic
    inputs; tests/golden/wordlevel.npz holds inputs and outputs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

TEMP = 0.07  # evaluate_spotting.py:39, evaluate_asd.py:43, plot_heatmap.py:34
POOL_MODES = ("mean_mean", "max_t_mean_w", "max_w_mean_t", "max_max")


def _f32(x) -> torch.Tensor:
    """torch.FloatTensor(np.asarray(x).astype(np.float32)) as at evaluate_spotting.py:46-47."""
    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float32).cpu()
    if isinstance(x, (list, tuple)):
        x = np.stack([np.asarray(v) for v in x])
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x).astype(np.float32)))


def normalize_rows(x, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, p=2, dim=-1): x / max(||x||, eps)  (inference_embs.py:630-636)."""
    x = _f32(x)
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


# --------------------------------------------------------------------------- retrieval
def mean_pool(emb) -> np.ndarray:
    """feats[...].mean(axis=0).squeeze() (evaluate_retrieval.py:30-31): numpy mean in the
    stored dtype (fp16 arrays accumulate in fp32 and round the result back to fp16)."""
    return np.asarray(emb).mean(axis=0).squeeze()


def get_similarity_matrix(emb1, emb2) -> torch.Tensor:
    """evaluate_retrieval.py:38-48."""
    a = normalize_rows(_f32(emb1))
    b = normalize_rows(_f32(emb2))
    return a @ b.t()


def compute_metrics(x) -> dict:
    """evaluate_retrieval.py:51-65, including its tie behaviour (every sorted position
    whose value equals the diagonal counts as a hit, so len(ind) can exceed N)."""
    x = np.asarray(x, dtype=np.float32)
    sx = np.sort(-x, axis=1)
    d = np.diag(-x)[:, np.newaxis]
    ind = np.where((sx - d) == 0)[1]
    m = {}
    for k in (1, 5, 10, 25, 50):  # R1 is an addition (BASELINE.json asks for recall@1/5/10)
        m[f"R{k}"] = float(np.sum(ind < k)) / len(ind)
    m["MR"] = float(np.median(ind) + 1)
    return m


def rank_counts(x) -> Tuple[np.ndarray, np.ndarray]:
    """(#{j: x_ij > x_ii}, #{j: x_ij == x_ii}) — the two integers compute_metrics' `ind`
    is made of: row i contributes positions n_greater, ..., n_greater + n_equal - 1."""
    x = np.asarray(x, dtype=np.float32)
    d = np.diag(x)[:, None]
    return (x > d).sum(1).astype(np.int32), (x == d).sum(1).astype(np.int32)


def metrics_from_counts(n_greater: np.ndarray, n_equal: np.ndarray) -> dict:
    ind = np.concatenate([g + np.arange(e) for g, e in zip(n_greater.tolist(), n_equal.tolist())])
    m = {}
    for k in (1, 5, 10, 25, 50):
        m[f"R{k}"] = float(np.sum(ind < k)) / len(ind)
    m["MR"] = float(np.median(ind) + 1)
    return m


# --------------------------------------------------------------------------- spotting
def get_attn_matrix(gesture, content, temp: float = TEMP, normalize: bool = True) -> np.ndarray:
    """evaluate_spotting.py:39-57 (normalize=True) / utils/plot_heatmap.py:34-59 (False):
    softmax over WORDS of (G C^T)/temp for every frame, returned transposed (W, T)."""
    g = _f32(gesture)
    c = _f32(content)
    if normalize:
        g = F.normalize(g, p=2, dim=-1)
        c = F.normalize(c, p=2, dim=-1)
    a = torch.mm(g, c.t()) / temp
    a = F.softmax(a, dim=1)
    return a.numpy().T  # == np.array(attn_mat).T of the reference (a transposed view)


def spot_decision(attn: np.ndarray, word_idx: int, start: int, end: int, thresh: float = 0.5,
                  frame_thresh: int = 9) -> Tuple[int, float, bool]:
    """evaluate_spotting.py:70-82: argmax frame of the target word's row, window + threshold."""
    pred = int(np.argmax(attn[word_idx]))
    score = float(attn[word_idx][pred])
    lo = max(start - frame_thresh, 0)
    hi = end + frame_thresh
    return pred, score, bool(lo <= pred <= hi and score >= thresh)


# --------------------------------------------------------------------------- ASD
def get_similarity_cos(query_emb, data_emb, temp: float = TEMP) -> np.ndarray:
    """evaluate_asd.py:43-51: nn.CosineSimilarity(dim=1) (each norm clamped at 1e-8),
    /temp, softmax over the candidates."""
    q = _f32(query_emb)
    d = _f32(data_emb)
    qn = q.norm(dim=1, keepdim=True).clamp_min(1e-8)
    dn = d.norm(dim=1, keepdim=True).clamp_min(1e-8)
    sim = ((q / qn) * (d / dn)).sum(dim=1)
    return F.softmax(sim / temp, dim=0).numpy()


def asd_mean_emb(emb) -> torch.Tensor:
    """torch.FloatTensor(emb.mean(axis=0)).unsqueeze(0) (evaluate_asd.py:31-36)."""
    return _f32(np.asarray(emb).mean(axis=0)).unsqueeze(0)


def asd_predict(content_emb, gesture_embs: Sequence, n_candidates: Sequence[int] = (2, 4, 6)) -> List[int]:
    """evaluate_asd.py:91-100: candidates = [positive, negatives...]; argmax over each prefix."""
    q = asd_mean_emb(content_emb)
    g = torch.cat([asd_mean_emb(e) for e in gesture_embs])
    return [int(np.argmax(get_similarity_cos(q, g[:n]))) for n in n_candidates if n <= len(gesture_embs)]


# --------------------------------------------------------------------------- sim-pool (restated)
def cos_tile(gesture, content, normalize: bool = True) -> torch.Tensor:
    """(T, W) cosine tile: F.normalize + torch.mm exactly as evaluate_spotting.py:49-52."""
    g = _f32(gesture)
    c = _f32(content)
    if normalize:
        g = F.normalize(g, p=2, dim=-1)
        c = F.normalize(c, p=2, dim=-1)
    return torch.mm(g, c.t())


def pool_tile(s: torch.Tensor, mode: str) -> float:
    if mode == "mean_mean":
        return float(s.mean())
    if mode == "max_t_mean_w":
        return float(s.amax(dim=0).mean())
    if mode == "max_w_mean_t":
        return float(s.amax(dim=1).mean())
    if mode == "max_max":
        return float(s.amax())
    raise ValueError(mode)


def simpool_allpairs_loop(gestures: Sequence, contents: Sequence, mode: str, normalize: bool = True) -> np.ndarray:
    """Literal double loop over clip pairs (small cases only)."""
    out = np.empty((len(gestures), len(contents)), dtype=np.float32)
    for i, g in enumerate(gestures):
        for j, c in enumerate(contents):
            out[i, j] = pool_tile(cos_tile(g, c, normalize), mode)
    return out


def _segment(x: torch.Tensor, lengths: torch.Tensor, op: str) -> torch.Tensor:
    return torch.segment_reduce(x, "max" if op == "max" else "mean", lengths=lengths, axis=0)


def simpool_allpairs(gestures: Sequence, contents: Sequence, mode: str, normalize: bool = True,
                     device: str = "cpu", rows_g: Optional[torch.Tensor] = None,
                     rows_c: Optional[torch.Tensor] = None) -> np.ndarray:
    """Vectorised restatement of simpool_allpairs_loop (same fp32 arithmetic: one big
    matmul of the normalised rows, then segment max/mean).  ``rows_g`` / ``rows_c`` let a
    test pass pre-rounded (e.g. bf16-valued) rows to isolate the kernel's own arithmetic."""
    lt = torch.tensor([len(g) for g in gestures], dtype=torch.int64, device=device)
    lw = torch.tensor([len(c) for c in contents], dtype=torch.int64, device=device)
    if rows_g is None:
        rows_g = torch.cat([normalize_rows(g) if normalize else _f32(g) for g in gestures])
    if rows_c is None:
        rows_c = torch.cat([normalize_rows(c) if normalize else _f32(c) for c in contents])
    s = rows_g.to(device=device, dtype=torch.float32) @ rows_c.to(device=device, dtype=torch.float32).t()
    op_t, op_w = {"mean_mean": ("mean", "mean"), "max_t_mean_w": ("max", "mean"),
                  "max_w_mean_t": ("mean", "max"), "max_max": ("max", "max")}[mode]
    if mode == "max_w_mean_t":  # words first
        s = _segment(s.t().contiguous(), lw, op_w).t().contiguous()  # (sumT, nC)
        s = _segment(s, lt, op_t)
    else:  # frames first (order is irrelevant when both ops agree)
        s = _segment(s, lt, op_t)  # (nG, sumW)
        s = _segment(s.t().contiguous(), lw, op_w).t()
    return s.contiguous().cpu().numpy()


def refnorm_scales(clips: Sequence, eps: float = 1e-12) -> np.ndarray:
    """1 / max(||mean row||, eps) per clip: the factor that turns the mean/mean pooled tile
    into the reference's cosine of mean-pooled embeddings."""
    return np.array([1.0 / max(float(_f32(mean_pool(c)).norm()), eps) for c in clips], dtype=np.float32)


def topk(scores, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row top-k, descending, ties towards the lower index (stable argsort of -x)."""
    x = np.asarray(scores, dtype=np.float32)
    idx = np.argsort(-x, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(x, idx, axis=1), idx.astype(np.int32)


# ------------------------------------------------------------------ word-level pooling (producer side)
def word_level_embs(text_emb, text, input_ids, offset_mapping, audio_emb=None, word_boundaries=None,
                    special_token_ids=(0, 2, 1)):
    """Restatement of JEGAL.get_word_level_embs (models/jegal.py:131-211), loop for loop."""
    word_text, word_audio, invalid = [], [], []
    for b in range(input_ids.shape[0]):
        starts = [i for i, off in enumerate(offset_mapping[b])
                  if int(off[0]) == 0 and int(input_ids[b][i]) not in special_token_ids]  # :146-149
        et, ea, valid = [], [], True
        if audio_emb is not None:
            actual_start = int(word_boundaries[b][0][1])  # :155
        for idx, _ in enumerate(text[b]):
            if idx >= len(starts):  # :162-166
                valid = False
                invalid.append(b)
                break
            hi = starts[idx + 1] if idx < len(starts) - 1 else input_ids.shape[1]  # :168-171
            sub = text_emb[b, starts[idx]:hi]
            et.append(sub.mean(dim=0) if len(sub) > 1 else sub[0])  # :176-179
            if audio_emb is not None:
                s0 = int(word_boundaries[b][idx][1]) - actual_start
                e0 = int(word_boundaries[b][idx][2]) - actual_start
                fr = audio_emb[b, s0:e0 + 1]  # :188-191
                ea.append(fr.mean(dim=0) if len(fr) > 1 else fr[0])  # :193-196
        if valid:
            if len(et) <= 0:
                invalid.append(b)
            else:
                word_text.append(torch.stack(et))
                if audio_emb is not None:
                    word_audio.append(torch.stack(ea))
    return word_text, word_audio, invalid


def audio_word_level_embs(audio_emb, word_boundaries, invalid_sample_idx=None):
    """Restatement of JEGAL.get_audio_word_level_embs (models/jegal.py:213-252)."""
    out = []
    for b in range(audio_emb.shape[0]):
        if invalid_sample_idx is not None and b in invalid_sample_idx:
            continue
        actual_start = int(word_boundaries[b][0][1])
        ea = []
        for idx in range(len(word_boundaries[b])):
            s0 = int(word_boundaries[b][idx][1]) - actual_start
            e0 = int(word_boundaries[b][idx][2]) - actual_start
            fr = audio_emb[b, s0:e0 + 1]
            ea.append(fr.mean(dim=0) if len(fr) > 1 else fr[0])
        if len(ea) > 0:
            out.append(torch.stack(ea))
        elif invalid_sample_idx is not None:
            invalid_sample_idx.append(b)
        else:
            invalid_sample_idx = [b]
    return out, invalid_sample_idx
