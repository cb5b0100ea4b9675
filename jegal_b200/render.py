"""Heatmap rendering without matplotlib / cv2: the picture half of the
reference's utils/plot_heatmap.py::plot (:62-107) — jet colormap, the thresholded overlay blended
with cv2.addWeighted(alpha = 0.6), one square cell per (word, frame) — written as a PNG with the
standard library only.  Host-side image plumbing (the reference's is too); the attention matrix
itself comes from K3.  Tick labels (the words on the y axis as at plot_heatmap.py:89-90, frame numbers on the x
axis, 0.0 .. 1.0 on the colour bar) are drawn with a built-in 5 x 7 bitmap font; the words also go into the PNG's
tEXt chunk.
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


# classic 5 x 7 LCD font, ASCII 0x20 .. 0x7E: five column bytes per glyph, bit 0 = top row
_FONT = bytes.fromhex(
    "0000000000" "00005F0000" "0007000700" "147F147F14" "242A7F2A12" "2313086462" "3649552250" "0005030000"
    "001C224100" "0041221C00" "082A1C2A08" "08083E0808" "0050300000" "0808080808" "0060600000" "2010080402"
    "3E5149453E" "00427F4000" "4261514946" "2141454B31" "1814127F10" "2745454539" "3C4A494930" "0171090503"
    "3649494936" "064949291E" "0036360000" "0056360000" "0008142241" "1414141414" "4122140800" "0201510906"
    "324979413E" "7E1111117E" "7F49494936" "3E41414122" "7F4141221C" "7F49494941" "7F09090101" "3E41415132"
    "7F0808087F" "00417F4100" "2040413F01" "7F08142241" "7F40404040" "7F0204027F" "7F0408107F" "3E4141413E"
    "7F09090906" "3E4151215E" "7F09192946" "4649494931" "01017F0101" "3F4040403F" "1F2040201F" "7F2018207F"
    "6314081463" "0304780403" "6151494543" "00007F4141" "0204081020" "41417F0000" "0402010204" "4040404040"
    "0001020400" "2054545478" "7F48444438" "3844444420" "384444487F" "3854545418" "087E090102" "081454543C"
    "7F08040478" "00447D4000" "2040443D00" "007F102844" "00417F4000" "7C04180478" "7C08040478" "3844444438"
    "7C14141408" "081414187C" "7C08040408" "4854545420" "043F444020" "3C4040207C" "1C2040201C" "3C4030403C"
    "4428102844" "0C5050503C" "4464544C44" "0008364100" "00007F0000" "0041360800" "08082A1C08")


def text_width(text: str, scale: int = 2) -> int:
    return 6 * scale * len(text)


def draw_text(img: np.ndarray, x: int, y: int, text: str, scale: int = 2, color=(0.0, 0.0, 0.0)) -> None:
    """Draw ASCII `text` with its top-left corner at (x, y) into an [H, W, 3] float image (clipped at the borders)."""
    h, w, _ = img.shape
    for k, ch in enumerate(text):
        o = ord(ch)
        g = _FONT[(o - 32) * 5:(o - 32) * 5 + 5] if 32 <= o <= 126 else _FONT[(ord("?") - 32) * 5:(ord("?") - 32) * 5 + 5]
        for cx in range(5):
            for ry in range(7):
                if (g[cx] >> ry) & 1:
                    x0, y0 = x + (k * 6 + cx) * scale, y + ry * scale
                    if x0 < 0 or y0 < 0 or x0 >= w or y0 >= h:
                        continue
                    img[y0:min(y0 + scale, h), x0:min(x0 + scale, w)] = color


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
                   alpha: float = 0.6, cell: int = 24, colorbar: bool = True, labels: bool = True) -> str:
    """plot(attn_mtx, words, fname) of the reference: writes <fname>.png ((W x T) cells of `cell` pixels, the words as
    y tick labels, frame numbers as x tick labels, a jet colour bar with its scale on the right) and returns its path."""
    img = heatmap_rgb(attn_mtx, thresh, alpha)
    n_w, n_t = img.shape[0], img.shape[1]
    img = np.repeat(np.repeat(img, cell, axis=0), cell, axis=1)
    scale = 2 if cell >= 16 else 1
    if colorbar:
        h = img.shape[0]
        bw = max(cell // 2, 4)
        bar = jet(np.linspace(1.0, 0.0, h))[:, None, :].repeat(bw, axis=1)
        gap = np.ones((h, bw, 3))
        img = np.concatenate([img, gap, bar], axis=1)
    if labels:
        words = [str(w) for w in words][:n_w]
        left = max([text_width(w, scale) for w in words] + [0]) + 3 * scale
        right = (text_width("0.0", scale) + 3 * scale) if colorbar else 0
        bottom = 7 * scale + 4 * scale
        canvas = np.ones((img.shape[0] + bottom, left + img.shape[1] + right, 3))
        canvas[: img.shape[0], left:left + img.shape[1]] = img
        for i, w in enumerate(words):  # ax.set_yticklabels(words), plot_heatmap.py:89-90
            draw_text(canvas, left - text_width(w, scale) - 2 * scale, i * cell + (cell - 7 * scale) // 2, w, scale)
        step = 5 if n_t * cell // max(1, text_width("000", scale) + scale) >= (n_t + 4) // 5 else 10
        for t in range(0, n_t, step):  # frame numbers under every `step`-th column
            lab = str(t)
            draw_text(canvas, left + t * cell + (cell - text_width(lab, scale)) // 2 + scale // 2, img.shape[0] + 2 * scale, lab, scale)
        if colorbar:
            for v in (0.0, 0.2, 0.4, 0.6, 0.8, 1.0):
                y = int(round((1.0 - v) * (img.shape[0] - 1))) - (7 * scale) // 2
                draw_text(canvas, left + img.shape[1] + 2 * scale, min(max(y, 0), img.shape[0] - 7 * scale), f"{v:.1f}", scale)
        img = canvas
    rgb8 = np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)
    path = fname + ".png"
    write_png(path, rgb8, " ".join(str(w) for w in words))
    return path
