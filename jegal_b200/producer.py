"""Producer side of the scoring path: what inference_embs.py:628-646 and
evaluation/extract_jegal_embs.py:107-123 do with the model's outputs — F.normalize(p=2, dim=-1),
`[0].cpu().numpy()`, one pickle per clip — kept as the archival format, plus a device-side sink
that hands the SAME embeddings to the scoring kernels without the fp16-numpy / pickle round trip:
the per-clip outputs are packed on the device and K0 normalises them in fp32 and rounds once to
the 16-bit operand type the tensor cores consume.

    sink = EmbeddingSink(res_dir="embs")            # res_dir=None: no .pkl files
    for clip in clips:
        g, c = jegal_model.forward_inference(...)    # (1, T, 512), (1, W, 512) on the GPU
        sink.add(g, c, {"fname": name, "word_boundaries": wb, "text": text})
    gest, cont, infos = sink.finish()                # scoring.PackedClips (already normalised 16-bit rows)
    scores = scoring.score_allpairs(gest, cont, "max_t_mean_w", normalize_rows=False)
"""
from __future__ import annotations

import os
from typing import Any, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import ops, pkl_io
from .ops import JegalError
from .scoring import PackedClips, layout_for


def _clip_2d(x: torch.Tensor, what: str) -> torch.Tensor:
    if x.dim() == 3:
        if x.shape[0] != 1:
            raise JegalError(f"{what}: expected a (1, L, 512) batch of one clip, got {tuple(x.shape)}")
        x = x[0]  # inference_embs.py:632,636
    if x.dim() != 2 or x.shape[1] != 512 or not x.is_cuda:
        raise JegalError(f"{what}: expected a CUDA (L, 512) embedding, got {tuple(x.shape)}")
    return x


class EmbeddingSink:
    """Collects the model's per-clip outputs on the device; ``finish`` packs and normalises them (K0)."""

    def __init__(self, res_dir: Optional[str] = None, op_dtype: torch.dtype = torch.bfloat16):
        self.res_dir = res_dir
        self.op_dtype = op_dtype
        self.gest: List[torch.Tensor] = []
        self.cont: List[torch.Tensor] = []
        self.infos: List[Any] = []
        if res_dir:
            os.makedirs(res_dir, exist_ok=True)

    def add(self, gesture_emb: torch.Tensor, content_emb: torch.Tensor, info: Any) -> None:
        g, c = _clip_2d(gesture_emb, "gesture_emb"), _clip_2d(content_emb, "content_emb")
        self.gest.append(g)
        self.cont.append(c)
        self.infos.append(info)
        if self.res_dir:  # the reference's file, bit for bit its recipe (inference_embs.py:629-646)
            fname = pkl_io.info_get(info, "fname") or pkl_io.info_get(info, "filename") or f"clip{len(self.infos) - 1}"
            pkl_io.write_pkl(os.path.join(self.res_dir, str(fname).replace("/", "__") + ".pkl"),
                             F.normalize(g, p=2, dim=-1).cpu().numpy(), F.normalize(c, p=2, dim=-1).cpu().numpy(), info)

    def _pack(self, clips: List[torch.Tensor]) -> PackedClips:
        lengths = np.array([int(x.shape[0]) for x in clips], dtype=np.int32)
        layout = layout_for(lengths)
        dt = torch.float16 if all(x.dtype == torch.float16 for x in clips) else torch.float32
        rows = torch.cat([x.to(dt) for x in clips], dim=0) if clips else torch.empty((0, 512), dtype=dt, device="cuda")
        rows16, _ = ops.prep(rows.contiguous(), layout, normalize=True, out_dtype=self.op_dtype)  # F.normalize eps 1e-12
        return PackedClips(rows16, layout)

    def finish(self) -> Tuple[PackedClips, PackedClips, List[Any]]:
        return self._pack(self.gest), self._pack(self.cont), self.infos
