"""Heatmap rendering without matplotlib / cv2: the picture half of the
reference's utils/plot_heatmap.py::plot (:62-107) — jet colormap, the thresholded overlay blended
with cv2.addWeighted(alpha = 0.6), one square cell per (word, frame) — written as a PNG with the
standard library only.  Host-side image plumbing (the reference's is too); the attention matrix
itself comes from K3.  Tick labels are not drawn: the words go into the PNG's tEXt chunk and the
caller prints them.
"""
from __future__ import annotations

import struct
import zlib
from typing import Sequence

import numpy as np

# matplotlib's 'jet' segment data (x, y) per channel, sampled into a 256-entry LUT like LinearSegmentedColormap
_JET = {
    "r": [(0.0, 0.0), (0.35, 0.0), (0.66, 1.0), (0.89, 1.0), (1.0, 0.5)],
    "g": [(0.0, 0.0), (0.125, 0.0), (0.375, 1.0), (0.64, 1.0), (0.91, 0.0), (1.0, 0.0)],
    "b": [(0.0, 0.5), (0.11, 1.0), (0.34, 1.0), (0.65, 0.0), (1.0, 0.0)],
}
_LUT = np.stack([np.interp(np.linspace(0.0, 1.0, 256), *zip(*_JET[c])) for c in "rgb"], axis=1)


def jet(x: np.ndarray) -> np.ndarray:
    """cmap('jet')(x) for floats in [0, 1]: [..., 3] RGB in [0, 1] (index = int(x * 256), clipped)."""
    idx = np.clip((np.asarray(x, dtype=np.float64) * 256).astype(np.int64), 0, 255)
    return _LUT[idx]


def heatmap_rgb(attn_mtx: np.ndarray, thresh: float = 0.8, alpha: float = 0.6) -> np.ndarray:
    """The blended image of plot_heatmap.py:78-87 composited on white: [W, T, 3] floats in [0, 1]."""
    a = np.asarray(attn_mtx, dtype=np.float64)
    base = jet(a)
    th = a.copy()
    th[th < thresh] = 0.01
    over = jet(th)
    beta = 1.0 - alpha
    rgb = alpha * over + beta * base            # cv2.addWeighted on the colour channels
    opacity = alpha * alpha + beta * 1.0        # ... and on the alpha channel (overlay alpha = alpha, base = 1)
    return opacity * rgb + (1.0 - opacity) * 1.0


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path: str, rgb8: np.ndarray, text: str = "") -> None:
    """8-bit RGB PNG (no interlace, filter 0) with an optional tEXt 'words' chunk."""
    h, w, _ = rgb8.shape
    raw = b"".join(b"\x00" + rgb8[y].tobytes() for y in range(h))
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
    if text:
        png += _chunk(b"tEXt", b"words\x00" + text.encode("latin-1", "replace"))
    png += _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def render_heatmap(attn_mtx: np.ndarray, words: Sequence[str], fname: str = "heatmap", thresh: float = 0.8,
                   alpha: float = 0.6, cell: int = 24, colorbar: bool = True) -> str:
    """plot(attn_mtx, words, fname) of the reference, minus the text labels: writes <fname>.png
    ((W x T) cells of `cell` pixels, a jet colour bar on the right) and returns its path."""
    img = heatmap_rgb(attn_mtx, thresh, alpha)
    img = np.repeat(np.repeat(img, cell, axis=0), cell, axis=1)
    if colorbar:
        h = img.shape[0]
        bar = jet(np.linspace(1.0, 0.0, h))[:, None, :].repeat(max(cell // 2, 4), axis=1)
        gap = np.ones((h, max(cell // 2, 4), 3))
        img = np.concatenate([img, gap, bar], axis=1)
    rgb8 = np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)
    path = fname + ".png"
    write_png(path, rgb8, " ".join(words))
    return path
