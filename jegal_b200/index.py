"""Packed clip index on disk.

The reference re-reads and un-pickles every clip for every run (evaluate_retrieval.py:25-33,
evaluate_asd.py:26-39 even re-reads the negatives per row).  A packed index stores one side of a
clip set once — raw rows exactly as in the .pkl files (fp16 or fp32, so nothing is lost), the ragged
offsets and clip names — as plain .npy files that np.load can memory-map, so a scoring run starts
with one sequential read instead of N pickle loads.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Sequence

import numpy as np
import torch


class ClipIndex:
    def __init__(self, rows: np.ndarray, cu_len: np.ndarray, names: Optional[List[str]] = None):
        assert rows.ndim == 2 and rows.shape[1] == 512 and cu_len[-1] == rows.shape[0]
        self.rows = rows
        self.cu_len = np.asarray(cu_len, dtype=np.int32)
        self.names = names or [str(i) for i in range(len(cu_len) - 1)]

    @classmethod
    def from_clips(cls, clips: Sequence[np.ndarray], names: Optional[List[str]] = None) -> "ClipIndex":
        dt = np.float16 if all(np.asarray(c).dtype == np.float16 for c in clips) else np.float32
        lengths = np.array([len(c) for c in clips], dtype=np.int64)
        cu = np.concatenate([[0], np.cumsum(lengths)])
        rows = np.concatenate([np.asarray(c, dtype=dt).reshape(-1, 512) for c in clips]) if len(clips) else np.zeros((0, 512), dt)
        return cls(rows, cu.astype(np.int32), names)

    def save(self, prefix: str) -> None:
        os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
        np.save(prefix + ".rows.npy", self.rows)
        np.save(prefix + ".cu.npy", self.cu_len)
        with open(prefix + ".names.json", "w") as f:
            json.dump(self.names, f)

    @classmethod
    def load(cls, prefix: str, mmap: bool = True) -> "ClipIndex":
        rows = np.load(prefix + ".rows.npy", mmap_mode="r" if mmap else None)
        cu = np.load(prefix + ".cu.npy")
        with open(prefix + ".names.json") as f:
            names = json.load(f)
        return cls(rows, cu, names)

    @property
    def n(self) -> int:
        return len(self.cu_len) - 1

    def clip(self, i: int) -> np.ndarray:
        return self.rows[self.cu_len[i]:self.cu_len[i + 1]]

    def to_packed(self):
        """Upload to the current CUDA device as a scoring.PackedClips."""
        from . import scoring

        host = torch.from_numpy(np.ascontiguousarray(self.rows))
        return scoring.PackedClips.from_packed(host.pin_memory() if host.numel() else host, self.cu_len)


def build_from_pkl_dir(path: str, out_prefix: str) -> dict:
    """One pass over a directory of reference .pkl files -> `<out_prefix>.gesture.*`, `<out_prefix>.content.*`."""
    from . import pkl_io

    d = pkl_io.load_dir(path)
    names = [os.path.basename(f)[:-4] for f in d["files"]]
    g = ClipIndex.from_clips(d["gesture"], names)
    c = ClipIndex.from_clips(d["content"], names)
    g.save(out_prefix + ".gesture")
    c.save(out_prefix + ".content")
    return dict(n=len(names), gesture_rows=int(g.rows.shape[0]), content_rows=int(c.rows.shape[0]))
