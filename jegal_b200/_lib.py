"""ctypes binding of include/jegal_b200.h.

The shared library is built in-tree (jegal_b200/csrc/Makefile ->
jegal_b200/libjegal_b200.so).  There is no fallback: if the library is missing
or the device is not sm_100 every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# JEGAL_B200_LIB: developer override for A/B kernel builds (scripts/k1_variants.sh); never a fallback
LIB_PATH = os.environ.get("JEGAL_B200_LIB") or os.path.join(_HERE, "libjegal_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

# every symbol include/jegal_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "jegal_version",
    "jegal_ctx_create",
    "jegal_ctx_destroy",
    "jegal_last_error",
    "jegal_launch_count",
    "jegal_layout_create",
    "jegal_layout_destroy",
    "jegal_layout_rows",
    "jegal_layout_clips",
    "jegal_prep",
    "jegal_simpool_allpairs",
    "jegal_topk",
    "jegal_topk_merge",
    "jegal_rank_of_positive",
    "jegal_spot",
    "jegal_simpool_pairs",
    "jegal_group_softmax",
    "jegal_plan_column_tiles",
    "jegal_exchange_create",
    "jegal_exchange_ipc_handle",
    "jegal_exchange_connect",
    "jegal_exchange_destroy",
    "jegal_topk_exchange",
    "jegal_segment_mean",
    "jegal_clip_means",
    "jegal_pair_cosine",
    "jegal_spot_dense",
    "jegal_qgather_create",
    "jegal_qgather_ipc_handle",
    "jegal_qgather_connect",
    "jegal_qgather_destroy",
    "jegal_prep_gather",
]

F32, F16, BF16 = 0, 1, 2
POOL_MEAN_MEAN, POOL_MAX_T_MEAN_W, POOL_MAX_W_MEAN_T, POOL_MAX_MAX = 0, 1, 2, 3
POOL_MODES = {
    "mean_mean": POOL_MEAN_MEAN,
    "max_t_mean_w": POOL_MAX_T_MEAN_W,
    "max_w_mean_t": POOL_MAX_W_MEAN_T,
    "max_max": POOL_MAX_MAX,
}

_lib = None


class JegalError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a with nvcc (no GPU needed)."""
    if force:
        subprocess.run(["make", "-C", CSRC_DIR, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise JegalError("building libjegal_b200.so failed")
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JegalError(
            f"{LIB_PATH} is missing: run `make -C {CSRC_DIR}` (or __graft_entry__.build()); "
            "there is no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.jegal_version.restype = C.c_char_p
    lib.jegal_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.jegal_ctx_destroy.argtypes = [vp]
    lib.jegal_ctx_destroy.restype = None
    lib.jegal_last_error.argtypes = [vp]
    lib.jegal_last_error.restype = C.c_char_p
    lib.jegal_launch_count.argtypes = [vp]
    lib.jegal_launch_count.restype = i64
    lib.jegal_layout_create.argtypes = [vp, C.POINTER(i32), i32, vp, C.POINTER(vp)]
    lib.jegal_layout_destroy.argtypes = [vp]
    lib.jegal_layout_destroy.restype = None
    lib.jegal_layout_rows.argtypes = [vp]
    lib.jegal_layout_rows.restype = i64
    lib.jegal_layout_clips.argtypes = [vp]
    lib.jegal_layout_clips.restype = i32
    lib.jegal_prep.argtypes = [vp, vp, vp, C.c_int, C.c_int, f32, f32, C.c_int, vp, vp, vp, vp]
    lib.jegal_simpool_allpairs.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, i64, i64, vp]
    lib.jegal_topk.argtypes = [vp, vp, i32, i32, i64, i32, i32, vp, vp, vp]
    lib.jegal_topk_merge.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.jegal_rank_of_positive.argtypes = [vp, vp, i32, i32, i64, i64, vp, vp, vp, vp]
    if hasattr(lib, "jegal_spot"):
        lib.jegal_spot.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, f32, vp, f32, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp]
    if hasattr(lib, "jegal_simpool_pairs"):
        lib.jegal_simpool_pairs.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, f32, C.c_int, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp]
    if hasattr(lib, "jegal_group_softmax"):
        lib.jegal_group_softmax.argtypes = [vp, vp, i32, i32, i64, f32, vp, vp, vp]
    if hasattr(lib, "jegal_topk_exchange"):
        lib.jegal_exchange_create.argtypes = [vp, i32, i32, i32, i32, C.POINTER(vp)]
        lib.jegal_exchange_ipc_handle.argtypes = [vp, vp]
        lib.jegal_exchange_connect.argtypes = [vp, vp]
        lib.jegal_exchange_destroy.argtypes = [vp]
        lib.jegal_exchange_destroy.restype = None
        lib.jegal_topk_exchange.argtypes = [vp, vp, vp, i32, i64, i32, vp, vp, vp]
    if hasattr(lib, "jegal_segment_mean"):
        lib.jegal_segment_mean.argtypes = [vp, vp, C.c_int, i64, i32, vp, vp, i32, vp, C.c_int, i64, i32, vp]
    if hasattr(lib, "jegal_prep_gather"):
        lib.jegal_qgather_create.argtypes = [vp, i32, i32, i64, C.POINTER(vp)]
        lib.jegal_qgather_ipc_handle.argtypes = [vp, vp]
        lib.jegal_qgather_connect.argtypes = [vp, vp]
        lib.jegal_qgather_destroy.argtypes = [vp]
        lib.jegal_qgather_destroy.restype = None
        lib.jegal_prep_gather.argtypes = [vp, vp, vp, C.c_int, i64, i32, C.c_int, f32, C.c_int, C.POINTER(vp), vp]
    if hasattr(lib, "jegal_spot_dense"):
        lib.jegal_spot_dense.argtypes = [vp, vp, i64, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, f32, vp, vp]
    if hasattr(lib, "jegal_clip_means"):
        lib.jegal_clip_means.argtypes = [vp, vp, vp, C.c_int, f32, C.c_int, vp, vp, vp]
    if hasattr(lib, "jegal_pair_cosine"):
        lib.jegal_pair_cosine.argtypes = [vp, vp, i64, vp, i64, C.c_int, vp, vp, i32, C.c_int, f32, vp, vp]
    if hasattr(lib, "jegal_plan_column_tiles"):
        lib.jegal_plan_column_tiles.argtypes = [C.POINTER(i32), i32, i32, i32, vp, i32, C.POINTER(i32)]
    for name in SYMBOLS:
        fn = getattr(lib, name, None)
        if fn is not None and fn.restype is C.c_int and name not in ("jegal_layout_clips",):
            fn.restype = C.c_int
    _lib = lib
    return lib
