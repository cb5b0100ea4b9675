"""Packed clip index on disk.

The reference re-reads and un-pickles every clip for every run (evaluate_retrieval.py:25-33,
evaluate_spotting.py:18-36; evaluate_asd.py:26-39 even re-reads the negatives of every row).  An index stores a
clip set ONCE, in the layout the kernels consume:

    <prefix>.gesture.rows.npy   [sum T, 512]  rows exactly as in the .pkl files (fp16 or fp32: nothing is lost)
    <prefix>.gesture.cu.npy     [n + 1] int32 ragged offsets (clip i = rows cu[i]:cu[i+1])
    <prefix>.gesture.mean.npy   [n, 512]      the clip's temporal mean, numpy semantics (load_feats'
                                              `.mean(axis=0)`, evaluate_retrieval.py:30-31, evaluate_asd.py:32,36)
    <prefix>.content.*          the same for the content side
    <prefix>.meta.json          clip names (= .pkl basenames) and the `info` fields the scripts read
                                (phrase, word_boundaries, target_word_boundary, filename / fname, text)

Everything is a plain .npy that np.load memory-maps; ``pinned()`` turns the rows into page-locked memory once, so
a scoring run is one sequential host->device stream (jegal_b200.streaming) instead of N pickle loads + N small
copies.  Inverse row norms are NOT stored: K3 / K4 derive them from the operand bytes they stage anyway
(normalisation fused into the load), a stored copy would only add bytes to read.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

INFO_KEYS = ("phrase", "word_boundaries", "target_word_boundary", "filename", "fname", "text")


class ClipIndex:
    def __init__(self, rows: np.ndarray, cu_len: np.ndarray, names: Optional[List[str]] = None,
                 mean: Optional[np.ndarray] = None):
        assert rows.ndim == 2 and rows.shape[1] == 512 and cu_len[-1] == rows.shape[0]
        self.rows = rows
        self.cu_len = np.asarray(cu_len, dtype=np.int32)
        self.names = names or [str(i) for i in range(len(cu_len) - 1)]
        self.mean = mean
        self._pinned: Optional[torch.Tensor] = None

    @classmethod
    def from_clips(cls, clips: Sequence[np.ndarray], names: Optional[List[str]] = None, with_mean: bool = True) -> "ClipIndex":
        dt = np.float16 if all(np.asarray(c).dtype == np.float16 for c in clips) else np.float32
        lengths = np.array([len(c) for c in clips], dtype=np.int64)
        cu = np.concatenate([[0], np.cumsum(lengths)])
        assert cu[-1] < 2**31
        rows = np.concatenate([np.asarray(c, dtype=dt).reshape(-1, 512) for c in clips]) if len(clips) else np.zeros((0, 512), dt)
        mean = None
        if with_mean:  # exactly what the reference's load_feats keeps: ndarray.mean(axis=0) in the stored dtype
            mean = np.stack([np.asarray(c, dtype=dt).reshape(-1, 512).mean(axis=0) for c in clips]) if len(clips) \
                else np.zeros((0, 512), dt)
        return cls(rows, cu.astype(np.int32), names, mean)

    def save(self, prefix: str) -> None:
        os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
        np.save(prefix + ".rows.npy", self.rows)
        np.save(prefix + ".cu.npy", self.cu_len)
        if self.mean is not None:
            np.save(prefix + ".mean.npy", self.mean)
        with open(prefix + ".names.json", "w") as f:
            json.dump(self.names, f)

    @classmethod
    def load(cls, prefix: str, mmap: bool = True) -> "ClipIndex":
        rows = np.load(prefix + ".rows.npy", mmap_mode="r" if mmap else None)
        cu = np.load(prefix + ".cu.npy")
        mean = np.load(prefix + ".mean.npy") if os.path.exists(prefix + ".mean.npy") else None
        with open(prefix + ".names.json") as f:
            names = json.load(f)
        return cls(rows, cu, names, mean)

    @property
    def n(self) -> int:
        return len(self.cu_len) - 1

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.cu_len)

    def clip(self, i: int) -> np.ndarray:
        return self.rows[self.cu_len[i]:self.cu_len[i + 1]]

    def pinned(self) -> torch.Tensor:
        """The rows as ONE page-locked host tensor (one sequential read of the memory-mapped file, done once)."""
        if self._pinned is None:
            t = torch.empty(self.rows.shape, dtype=torch.float16 if self.rows.dtype == np.float16 else torch.float32,
                            pin_memory=torch.cuda.is_available())
            np.copyto(t.numpy(), self.rows)
            self._pinned = t
        return self._pinned

    def to_packed(self):
        """Upload to the current CUDA device as a scoring.PackedClips."""
        from . import scoring

        host = self.pinned()
        return scoring.PackedClips.from_packed(host, self.cu_len)


def _info_dict(info: Any) -> Dict[str, Any]:
    from . import pkl_io

    out = {}
    for k in INFO_KEYS:
        v = pkl_io.info_get(info, k)
        if v is not None:
            out[k] = v if isinstance(v, (str, int, float)) else json.loads(json.dumps(v, default=str))
    return out


def build_from_pkl_dir(path: str, out_prefix: str, files: Optional[Sequence[str]] = None) -> dict:
    """One pass over a directory of reference .pkl files -> `<out_prefix>.gesture.*`, `<out_prefix>.content.*`,
    `<out_prefix>.meta.json`."""
    from . import pkl_io

    d = pkl_io.load_dir(path, files=files)
    names = [os.path.basename(f)[:-4] for f in d["files"]]
    g = ClipIndex.from_clips(d["gesture"], names)
    c = ClipIndex.from_clips(d["content"], names)
    g.save(out_prefix + ".gesture")
    c.save(out_prefix + ".content")
    with open(out_prefix + ".meta.json", "w") as f:
        json.dump({"names": names, "info": [_info_dict(i) for i in d["info"]]}, f)
    return dict(n=len(names), gesture_rows=int(g.rows.shape[0]), content_rows=int(c.rows.shape[0]))


class ClipSetIndex:
    """Both sides + metadata of one indexed .pkl directory."""

    def __init__(self, gesture: ClipIndex, content: ClipIndex, info: List[dict]):
        self.gesture, self.content, self.info = gesture, content, info
        self.names = gesture.names

    @classmethod
    def load(cls, prefix: str) -> "ClipSetIndex":
        with open(prefix + ".meta.json") as f:
            meta = json.load(f)
        return cls(ClipIndex.load(prefix + ".gesture"), ClipIndex.load(prefix + ".content"), meta["info"])

    @property
    def n(self) -> int:
        return self.gesture.n


def load_or_build(path: str, index_prefix: str) -> ClipSetIndex:
    """The `--index PREFIX` option of the scripts: use the index if it exists, build it from `path` otherwise."""
    if not os.path.exists(index_prefix + ".meta.json"):
        build_from_pkl_dir(path, index_prefix)
    return ClipSetIndex.load(index_prefix)
