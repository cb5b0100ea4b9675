"""Device-tensor wrappers over the C ABI (include/jegal_b200.h).

torch is used only as the owner of device memory and streams; every compute
step is a call into libjegal_b200.so.  All functions raise ``JegalError`` when
the library or an sm_100 device is missing — there is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import JegalError, POOL_MODES

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class Context:
    """One per device; owns nothing but the error string and launch counter."""

    _by_device = {}

    def __init__(self, device: int):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise JegalError("no CUDA device: jegal_b200 has no CPU path")
        h = C.c_void_p()
        rc = self.lib.jegal_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise JegalError(f"jegal_ctx_create(device={device}) failed with {rc} (sm_100 device required)")
        self.h = h
        self.device = int(device)

    @classmethod
    def get(cls, device: Optional[int] = None) -> "Context":
        if device is None:
            device = torch.cuda.current_device()
        if device not in cls._by_device:
            cls._by_device[device] = Context(device)
        return cls._by_device[device]

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.jegal_last_error(self.h).decode()
            raise JegalError(f"{what} failed ({rc}): {msg}")

    @property
    def launches(self) -> int:
        return int(self.lib.jegal_launch_count(self.h))


class Layout:
    """Ragged layout of one packed operand (clip i = rows cu_len[i]:cu_len[i+1])."""

    def __init__(self, cu_len: Sequence[int], ctx: Optional[Context] = None):
        self.ctx = ctx or Context.get()
        cu = np.ascontiguousarray(np.asarray(cu_len, dtype=np.int32))
        if cu.ndim != 1 or cu.size < 1:
            raise JegalError("cu_len must be a 1-D array of n_clips + 1 offsets")
        self.cu_len = cu
        self.n_clips = int(cu.size - 1)
        self.rows = int(cu[-1])
        h = C.c_void_p()
        rc = self.ctx.lib.jegal_layout_create(
            self.ctx.h, cu.ctypes.data_as(C.POINTER(C.c_int32)), self.n_clips, _stream(), C.byref(h)
        )
        self.ctx.check(rc, "jegal_layout_create")
        self.h = h

    @classmethod
    def from_lengths(cls, lengths: Sequence[int], ctx: Optional[Context] = None) -> "Layout":
        cu = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lengths, dtype=np.int64), out=cu[1:])
        if cu[-1] >= 2**31:
            raise JegalError("more than 2^31 rows in one operand")
        return cls(cu.astype(np.int32), ctx)

    @property
    def lengths(self) -> np.ndarray:
        return np.diff(self.cu_len)

    def __del__(self):
        try:
            if getattr(self, "h", None) is not None and self.h.value:
                self.ctx.lib.jegal_layout_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _check_rows(x: torch.Tensor, layout: Layout, what: str) -> None:
    if not x.is_cuda or x.dim() != 2 or x.shape[1] != 512 or not x.is_contiguous():
        raise JegalError(f"{what}: expected a contiguous CUDA tensor of shape [rows, 512]")
    if x.shape[0] != layout.rows:
        raise JegalError(f"{what}: {x.shape[0]} rows but the layout has {layout.rows}")


def prep(
    emb: torch.Tensor,
    layout: Layout,
    normalize: bool = True,
    out_dtype: torch.dtype = torch.bfloat16,
    want_mean_scale: bool = False,
    row_eps: float = 1e-12,
    mean_eps: float = 1e-12,
    out: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """K0: normalise + cast rows; optionally 1/||mean row|| per clip."""
    _check_rows(emb, layout, "prep")
    if emb.dtype not in _DT or out_dtype not in (torch.bfloat16, torch.float16):
        raise JegalError("prep: unsupported dtype")
    ctx = layout.ctx
    if out is None:
        out = torch.empty((layout.rows, 512), dtype=out_dtype, device=emb.device)
    scale = torch.empty((layout.n_clips,), dtype=torch.float32, device=emb.device) if want_mean_scale else None
    rc = ctx.lib.jegal_prep(
        ctx.h, layout.h, _ptr(emb), _DT[emb.dtype], int(normalize), row_eps, mean_eps, _DT[out_dtype],
        _ptr(out), _ptr(scale), _stream(),
    )
    ctx.check(rc, "jegal_prep")
    return out, scale


def simpool_allpairs(
    gest_rows: torch.Tensor,
    gest_layout: Layout,
    cont_rows: torch.Tensor,
    cont_layout: Layout,
    mode: str = "mean_mean",
    gscale: Optional[torch.Tensor] = None,
    cscale: Optional[torch.Tensor] = None,
    content_major: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """K1: pooled score of every (gesture clip, content clip) pair.

    Returns [n_gest, n_cont] fp32, or [n_cont, n_gest] when ``content_major``.
    """
    _check_rows(gest_rows, gest_layout, "simpool gest")
    _check_rows(cont_rows, cont_layout, "simpool cont")
    if gest_rows.dtype != cont_rows.dtype or gest_rows.dtype not in (torch.bfloat16, torch.float16):
        raise JegalError("simpool: operands must both be bf16 or both fp16 (outputs of prep)")
    ctx = gest_layout.ctx
    nG, nC = gest_layout.n_clips, cont_layout.n_clips
    shape = (nC, nG) if content_major else (nG, nC)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=gest_rows.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise JegalError("simpool: bad out tensor")
    ld_g, ld_c = (1, nG) if content_major else (nC, 1)
    rc = ctx.lib.jegal_simpool_allpairs(
        ctx.h, gest_layout.h, _ptr(gest_rows), cont_layout.h, _ptr(cont_rows), _DT[gest_rows.dtype],
        POOL_MODES[mode], _ptr(gscale), _ptr(cscale), _ptr(out), ld_g, ld_c, _stream(),
    )
    ctx.check(rc, "jegal_simpool_allpairs")
    return out


def topk(scores: torch.Tensor, k: int, idx_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """K2: per-row top-k (descending, ties -> lower index)."""
    if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 2 or scores.stride(1) != 1:
        raise JegalError("topk: expected a CUDA fp32 matrix with unit column stride")
    ctx = Context.get(scores.device.index)
    nq, ng = scores.shape
    val = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    idx = torch.empty((nq, k), dtype=torch.int32, device=scores.device)
    rc = ctx.lib.jegal_topk(ctx.h, _ptr(scores), nq, ng, scores.stride(0), k, idx_offset, _ptr(val), _ptr(idx), _stream())
    ctx.check(rc, "jegal_topk")
    return val, idx


def topk_merge(vals: torch.Tensor, idxs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge [n_lists, n_q, k] sorted lists into [n_q, k]."""
    if vals.shape != idxs.shape or vals.dim() != 3 or not vals.is_contiguous() or not idxs.is_contiguous():
        raise JegalError("topk_merge: expected contiguous [n_lists, n_q, k] tensors")
    if vals.dtype != torch.float32 or idxs.dtype != torch.int32:
        raise JegalError("topk_merge: fp32 values and int32 indices")
    ctx = Context.get(vals.device.index)
    nl, nq, k = vals.shape
    oval = torch.empty((nq, k), dtype=torch.float32, device=vals.device)
    oidx = torch.empty((nq, k), dtype=torch.int32, device=vals.device)
    rc = ctx.lib.jegal_topk_merge(ctx.h, _ptr(vals), _ptr(idxs), nl, nq, k, _ptr(oval), _ptr(oidx), _stream())
    ctx.check(rc, "jegal_topk_merge")
    return oval, oidx


def rank_of_positive(
    scores: torch.Tensor, gt: Optional[torch.Tensor] = None
) -> Tuple[torch.Tensor, torch.Tensor]:
    """(#entries > positive, #entries == positive) per row; any 2-D strides."""
    if not scores.is_cuda or scores.dtype != torch.float32 or scores.dim() != 2:
        raise JegalError("rank_of_positive: expected a CUDA fp32 matrix")
    ctx = Context.get(scores.device.index)
    nq, ng = scores.shape
    ngt = torch.empty((nq,), dtype=torch.int32, device=scores.device)
    neq = torch.empty((nq,), dtype=torch.int32, device=scores.device)
    if gt is not None and (gt.dtype != torch.int32 or not gt.is_contiguous()):
        raise JegalError("rank_of_positive: gt must be contiguous int32")
    rc = ctx.lib.jegal_rank_of_positive(
        ctx.h, _ptr(scores), nq, ng, scores.stride(0), scores.stride(1), _ptr(gt), _ptr(ngt), _ptr(neq), _stream()
    )
    ctx.check(rc, "jegal_rank_of_positive")
    return ngt, neq
