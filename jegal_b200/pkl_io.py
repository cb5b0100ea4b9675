"""Reader / writer for the reference's embedding files (format unchanged).

One .pkl per clip: {"gesture_emb": (T, 512) ndarray, "content_emb": (W, 512) ndarray, "info": ...}
  * evaluation/extract_jegal_embs.py:107-123 writes `<video>__<clip>.pkl` with `info` = the pandas
    row of the benchmark CSV (phrase, word_boundaries (str), target_word_boundary (str), filename, ...);
  * inference_embs.py:629-646 writes `info` = {"fname", "word_boundaries": [[word, start, end], ...], "text"}.
Both flavours load here; `word_boundaries` may be a list or its string form.
"""
from __future__ import annotations

import ast
import os
import pickle
from concurrent.futures import ThreadPoolExecutor
from glob import glob
from typing import Any, Dict, List, Optional, Sequence

import numpy as np


def info_get(info: Any, key: str, default=None):
    """`info` is a dict (inference_embs.py) or a pandas Series (extract_jegal_embs.py)."""
    try:
        if hasattr(info, "get"):
            v = info.get(key, default)
        else:
            v = getattr(info, key, default)
    except Exception:
        v = default
    return default if v is None else v


def parse_boundaries(wb) -> list:
    return ast.literal_eval(wb) if isinstance(wb, str) else list(wb)


def load_pkl(path: str) -> Dict[str, Any]:
    with open(path, "rb") as f:
        return pickle.load(f)


def load_dir(path: str, threads: int = 8, files: Optional[Sequence[str]] = None) -> Dict[str, list]:
    """All clips of a directory, in sorted file order (the reference globs unsorted; order does not
    change any metric).  Returns dict(files, gesture, content, info)."""
    files = sorted(glob(os.path.join(path, "*.pkl"))) if files is None else list(files)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        feats = list(ex.map(load_pkl, files))
    return dict(
        files=files,
        gesture=[np.asarray(f["gesture_emb"]) for f in feats],
        content=[np.asarray(f["content_emb"]) for f in feats],
        info=[f["info"] for f in feats],
    )


def clip_pkl_name(filename: str) -> str:
    """`<video>/<clip>` -> `<video>__<clip>.pkl` (extract_jegal_embs.py:120, evaluate_asd.py:65)."""
    a, b = filename.split("/")[:2]
    return f"{a}__{b}.pkl"


def write_pkl(path: str, gesture_emb: np.ndarray, content_emb: np.ndarray, info: Any) -> None:
    with open(path, "wb") as f:
        pickle.dump({"gesture_emb": gesture_emb, "content_emb": content_emb, "info": info}, f)
