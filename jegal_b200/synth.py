"""Seeded synthetic embeddings shaped like the reference's benchmarks.

There is no network and the JEGAL checkpoints are external downloads
(README.md:52-60 of the reference), so every workload uses structured random
embeddings: clip topic z_i, word factor u_{i,w};
    content  c_{i,w} = norm(a z_i + b u_{i,w} + s eps)
    gesture  g_{i,t} = norm(a z_i + b u_{i,w(t)} + s eps')   (frames outside any word drop the u term)
so diagonal pairs score highest, heatmaps peak inside the right word and the ASD
positive usually wins.  Rows are unit-norm and stored as fp16, which is what the
reference's .pkl files hold (inference_embs.py:614,629-646 run under autocast).

Ragged structure (lengths, word boundaries) comes from a numpy Generator, values
from a torch Generator on the requested device.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

D = 512

# word boundaries of the reference's two sample clips in frame units @25 fps
# (samples/sample1.txt, samples/sample2.txt through inference_embs.py:367-369); T = 56 / 68.
SAMPLE_CLIPS = [
    (56, [["amount", 1, 6], ["of", 7, 9], ["numbers", 9, 16], ["inside", 17, 24], ["the", 24, 27],
          ["hat", 27, 32], ["mixing", 43, 51], ["the", 52, 54]]),
    (68, [["merging", 0, 7], ["the", 8, 10], ["two", 11, 14], ["diseases", 16, 27], ["arthritis", 36, 45],
          ["deformans", 45, 60], ["and", 64, 66]]),
]


@dataclass
class ClipSet:
    """Packed ragged clips: rows of clip i are [cu[i], cu[i+1])."""
    gest: torch.Tensor            # [sum T, 512] fp16
    cont: torch.Tensor            # [sum W, 512] fp16
    cu_t: np.ndarray              # int32 [n+1]
    cu_w: np.ndarray              # int32 [n+1]
    boundaries: List[List[list]] = field(default_factory=list)   # per clip [[word, start, end], ...]
    target_word: Optional[np.ndarray] = None                     # int32 [n] (spotting)

    @property
    def n(self) -> int:
        return len(self.cu_t) - 1

    def gesture(self, i: int) -> torch.Tensor:
        return self.gest[self.cu_t[i]:self.cu_t[i + 1]]

    def content(self, i: int) -> torch.Tensor:
        return self.cont[self.cu_w[i]:self.cu_w[i + 1]]

    def gesture_list(self) -> List[np.ndarray]:
        g = self.gest.cpu().numpy()
        return [g[self.cu_t[i]:self.cu_t[i + 1]] for i in range(self.n)]

    def content_list(self) -> List[np.ndarray]:
        c = self.cont.cpu().numpy()
        return [c[self.cu_w[i]:self.cu_w[i + 1]] for i in range(self.n)]


def _cu(lengths: np.ndarray) -> np.ndarray:
    cu = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=cu[1:])
    assert cu[-1] < 2**31
    return cu.astype(np.int32)


def random_boundaries(rng: np.random.Generator, T: int, W: int) -> List[list]:
    """W words covering disjoint inclusive frame ranges inside [0, T), in order, gaps allowed."""
    assert T >= W >= 1
    cuts = np.sort(rng.choice(np.arange(1, T), size=W - 1, replace=False)) if W > 1 else np.array([], dtype=int)
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [T]])  # exclusive
    out = []
    for w in range(W):
        s, e = int(starts[w]), int(ends[w]) - 1
        if e - s >= 3 and rng.random() < 0.3:  # leave a silent gap after some words
            e -= int(rng.integers(1, max(2, (e - s) // 2)))
        out.append([f"w{w}", s, e])
    return out


def _unit_rows(x: torch.Tensor) -> torch.Tensor:
    return (x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)).to(torch.float16)


def make_clipset(len_t: Sequence[int], len_w: Sequence[int], seed: int, device="cpu",
                 boundaries: Optional[List[List[list]]] = None, a: float = 0.5, b: float = 1.0,
                 sigma: float = 1.0, with_targets: bool = False) -> ClipSet:
    len_t = np.asarray(len_t, dtype=np.int64)
    len_w = np.asarray(len_w, dtype=np.int64)
    n = len(len_t)
    rng = np.random.default_rng(seed)
    if boundaries is None:
        boundaries = [random_boundaries(rng, int(len_t[i]), int(len_w[i])) for i in range(n)]
    cu_t, cu_w = _cu(len_t), _cu(len_w)
    # frame -> global word row (or -1)
    frame_word = np.full(int(cu_t[-1]), -1, dtype=np.int64)
    for i in range(n):
        for w, (_, s, e) in enumerate(boundaries[i]):
            frame_word[cu_t[i] + s: cu_t[i] + min(e, len_t[i] - 1) + 1] = cu_w[i] + w
    frame_clip = np.repeat(np.arange(n), len_t)
    word_clip = np.repeat(np.arange(n), len_w)
    gen = torch.Generator(device=device).manual_seed(int(seed) + 7919)
    dev = torch.device(device)
    Z = torch.randn((n, D), generator=gen, device=dev)
    U = torch.randn((int(cu_w[-1]), D), generator=gen, device=dev)
    wc = torch.from_numpy(word_clip).to(dev)
    cont = a * Z[wc] + b * U + sigma * torch.randn((int(cu_w[-1]), D), generator=gen, device=dev)
    fc = torch.from_numpy(frame_clip).to(dev)
    fw = torch.from_numpy(frame_word).to(dev)
    gest = a * Z[fc] + sigma * torch.randn((int(cu_t[-1]), D), generator=gen, device=dev)
    has = fw >= 0
    gest[has] += b * U[fw[has]]
    cs = ClipSet(_unit_rows(gest), _unit_rows(cont), cu_t, cu_w, boundaries)
    if with_targets:
        cs.target_word = np.array([rng.integers(0, len_w[i]) for i in range(n)], dtype=np.int32)
    return cs


# ------------------------------------------------------------------ the five BASELINE.json configs
def cfg1_samples(seed: int = 1235, device="cpu") -> ClipSet:
    """Two clips at the shapes of samples/sample1, sample2 (T, W) = (56, 8), (68, 7)."""
    lt = [c[0] for c in SAMPLE_CLIPS]
    lw = [len(c[1]) for c in SAMPLE_CLIPS]
    return make_clipset(lt, lw, seed, device, boundaries=[c[1] for c in SAMPLE_CLIPS], with_targets=True)


def cfg2_retrieval(n: int = 1000, seed: int = 1236, device="cpu", **kw) -> ClipSet:
    """AVS-Ret-shaped: T ~ U{25..200}, W ~ U{4..40}."""
    rng = np.random.default_rng(seed)
    lt = rng.integers(25, 201, size=n)
    lw = np.minimum(rng.integers(4, 41, size=n), lt)
    return make_clipset(lt, lw, seed, device, **kw)


def cfg3_spotting(n: int = 20000, seed: int = 1237, device="cpu", **kw) -> ClipSet:
    """AVS-Spot-shaped: T in 25..220 (mean ~69), W in 4..12, one target word per clip."""
    rng = np.random.default_rng(seed)
    lt = np.clip(25 + rng.gamma(shape=2.2, scale=20.0, size=n), 25, 220).astype(np.int64)
    lw = rng.integers(4, 13, size=n)
    return make_clipset(lt, lw, seed, device, with_targets=True, **kw)


@dataclass
class AsdSet:
    clips: ClipSet              # n_groups * tracks gesture clips + n_groups content clips (see below)
    cont_cu: np.ndarray         # content layout: one content clip per group
    pair_gest: np.ndarray       # int32 [n_groups * tracks] gesture clip of every candidate (track 0 = positive)
    pair_cont: np.ndarray       # int32 [n_groups * tracks] content clip of every candidate
    tracks: int


def cfg4_asd(n_groups: int = 10000, tracks: int = 4, seed: int = 1238, device="cpu",
             t_range: Tuple[int, int] = (39, 191), **kw) -> AsdSet:
    """AVS-Asd-shaped: per group one content track (W 5..17) and `tracks` gesture tracks
    (T 39..191); track 0 is the true speaker, the others are unrelated clips."""
    rng = np.random.default_rng(seed)
    n = n_groups * tracks
    lt = rng.integers(t_range[0], t_range[1] + 1, size=n)
    lw = np.minimum(rng.integers(5, 18, size=n), lt)
    cs = make_clipset(lt, lw, seed, device, **kw)  # every gesture clip comes with its own (matching) content
    pos = np.arange(n_groups) * tracks
    pair_gest = np.arange(n, dtype=np.int32)
    pair_cont = np.repeat(pos, tracks).astype(np.int32)  # all candidates of a group face the positive's content
    return AsdSet(cs, cs.cu_w, pair_gest, pair_cont, tracks)


def cfg5_gallery(n_query: int = 1000, n_gallery: int = 65536, T: int = 64, W: int = 16, seed: int = 1239,
                 device="cpu", n_related: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
    """Large-gallery retrieval: queries [n_query*T, 512] fp16 (gesture clips), gallery
    [n_gallery*W, 512] fp16 (content clips).  Query q's true match is gallery clip gt[q]."""
    dev = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(seed)
    rng = np.random.default_rng(seed)
    gt = rng.choice(n_gallery, size=n_query, replace=n_query > n_gallery).astype(np.int32)
    gal = torch.randn((n_gallery * W, D), generator=gen, device=dev)
    q = 1.0 * torch.randn((n_query * T, D), generator=gen, device=dev)
    # frames of query q echo the words of its match: frame t speaks word t * W // T
    word_of_frame = (torch.arange(T, device=dev) * W) // T
    rows = (torch.from_numpy(gt.astype(np.int64)).to(dev)[:, None] * W + word_of_frame[None, :]).reshape(-1)
    q += 1.0 * gal[rows]
    return _unit_rows(q), _unit_rows(gal), gt


CFG5_BLOCK = 4096  # gallery clips per generator block


def cfg5_sharded(n_query: int = 1000, n_gallery: int = 65536, T: int = 64, W: int = 16, seed: int = 1239, device="cpu",
                 lo: int = 0, hi: Optional[int] = None, want_queries: bool = True):
    """The planted large-gallery workload, generated so that ANY rank can materialise ANY clip range bit-identically:
    the gallery is drawn in blocks of CFG5_BLOCK clips, block b from its own generator (seed, b), and the queries
    (frames of query q echo the words of gallery clip gt[q], frame t speaks word t * W // T) from a third stream.
    Returns (queries [n_query*T, 512] fp16 | None, gallery rows of clips [lo, hi) fp16, gt int32 [n_query]);
    the shard of an N-GPU run is exactly rows [lo*W, hi*W) of the 1-GPU gallery."""
    dev = torch.device(device)
    hi = n_gallery if hi is None else hi
    rng = np.random.default_rng(seed)
    gt = rng.choice(n_gallery, size=n_query, replace=n_query > n_gallery).astype(np.int32)
    shard = torch.empty((max(hi - lo, 0) * W, D), dtype=torch.float16, device=dev)
    q = None
    if want_queries:
        gen_q = torch.Generator(device=device).manual_seed(seed * 1000003 + 999983)
        q = torch.randn((n_query * T, D), generator=gen_q, device=dev)
        word_of_frame = (torch.arange(T, device=dev) * W) // T
    n_blocks = (n_gallery + CFG5_BLOCK - 1) // CFG5_BLOCK
    for b in range(n_blocks):
        b0, b1 = b * CFG5_BLOCK, min((b + 1) * CFG5_BLOCK, n_gallery)
        in_shard = b0 < hi and b1 > lo
        mine = np.nonzero((gt >= b0) & (gt < b1))[0] if want_queries else np.zeros(0, dtype=np.int64)
        if not in_shard and mine.size == 0:
            continue
        gen = torch.Generator(device=device).manual_seed(seed * 1000003 + b)
        blk = torch.randn(((b1 - b0) * W, D), generator=gen, device=dev)
        if mine.size:
            qi = torch.from_numpy(mine).to(dev)
            rows = ((torch.from_numpy(gt[mine].astype(np.int64)).to(dev) - b0)[:, None] * W + word_of_frame[None, :])  # [m, T]
            dst = (qi[:, None] * T + torch.arange(T, device=dev)[None, :]).reshape(-1)
            q[dst] += blk[rows.reshape(-1)]
        if in_shard:
            s0, s1 = max(lo, b0), min(hi, b1)
            shard[(s0 - lo) * W:(s1 - lo) * W] = _unit_rows(blk[(s0 - b0) * W:(s1 - b0) * W])
        del blk
    return (None if q is None else _unit_rows(q)), shard, gt
