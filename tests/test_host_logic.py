"""CPU-side checks: the C-ABI library loads and exports every symbol include/jegal_b200.h
declares, the host logic of the scoring mirror, the synthetic generators, and that the
product refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

from jegal_b200 import _lib, scoring, sharded, synth
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "jegal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jegal_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jegal_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names
    assert b"sm_100a" in lib.jegal_version()


def test_library_is_sm100a_only():
    """The shipped .so carries sm_100a SASS with tcgen05 / TMA / TMEM instructions."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    g = [np.zeros((3, 512), np.float16)]
    with pytest.raises(_lib.JegalError):
        scoring.score_allpairs(g, g, "mean_mean")
    with pytest.raises(_lib.JegalError):
        scoring.compute_metrics(np.eye(4, dtype=np.float32))
    with pytest.raises(_lib.JegalError):
        scoring.get_similarity_cos(np.ones((1, 512), np.float32), np.ones((2, 512), np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jegal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    assert "oracle" not in open(os.path.join(ROOT, "jegal_b200", "scoring.py")).read().replace("the oracle", "")


def test_metrics_from_counts_matches_reference_semantics():
    rng = np.random.default_rng(3)
    for x in (rng.standard_normal((50, 50)).astype(np.float32), rng.integers(0, 4, (30, 30)).astype(np.float32)):
        ng, ne = oracle.rank_counts(x)
        assert scoring._metrics_from_counts(ng, ne) == oracle.compute_metrics(x)


def test_shard_ranges_cover_gallery():
    for n, w in [(65536, 8), (10, 4), (3, 8), (0, 2), (1000, 3)]:
        prev = 0
        for r in range(w):
            lo, hi = sharded.shard_range(n, r, w)
            assert lo == prev and lo <= hi <= n
            prev = hi
        assert prev == n


def test_synth_shapes_and_norms():
    cs = synth.cfg1_samples()
    assert [len(g) for g in cs.gesture_list()] == [56, 68] and [len(c) for c in cs.content_list()] == [8, 7]
    n = np.linalg.norm(cs.gest.float().numpy(), axis=1)
    assert np.abs(n - 1).max() < 2e-3 and cs.gest.dtype == torch.float16
    cs = synth.cfg2_retrieval(n=50)
    lt, lw = np.diff(cs.cu_t), np.diff(cs.cu_w)
    assert lt.min() >= 25 and lt.max() <= 200 and lw.min() >= 4 and lw.max() <= 40
    for i in range(cs.n):
        for (_, s, e) in cs.boundaries[i]:
            assert 0 <= s <= e < lt[i]
    cs = synth.cfg3_spotting(n=40)
    assert (np.diff(cs.cu_t) >= 25).all() and (np.diff(cs.cu_t) <= 220).all() and (np.diff(cs.cu_w) <= 12).all()
    ds = synth.cfg4_asd(n_groups=5, tracks=4)
    assert ds.pair_gest.tolist() == list(range(20)) and ds.pair_cont[:8].tolist() == [0, 0, 0, 0, 4, 4, 4, 4]
    q, g, gt = synth.cfg5_gallery(4, 64, 8, 4)
    assert q.shape == (32, 512) and g.shape == (256, 512) and len(gt) == 4


def test_synth_structure_is_learnable():
    """Diagonal pairs win and heatmaps peak inside the target word: the synthetic data exercises
    real decisions (not degenerate ties)."""
    cs = synth.make_clipset([60] * 12, [8] * 12, seed=9, with_targets=True)
    s = oracle.simpool_allpairs(cs.gesture_list(), cs.content_list(), "max_t_mean_w")
    assert (s.argmax(1) == np.arange(12)).mean() > 0.9
    hit = 0
    for i in range(cs.n):
        a = oracle.get_attn_matrix(cs.gesture(i).numpy(), cs.content(i).numpy())
        w = int(cs.target_word[i])
        _, s0, e0 = cs.boundaries[i][w]
        hit += oracle.spot_decision(a, w, s0, e0)[2]
    assert hit >= 8


# ----------------------------------------------------------------------------- K1 tile planner (host-only ABI call)
def _plan(lengths, width, allow_split):
    import ctypes as C
    lib = _lib.load()
    cu = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)
    n = C.c_int32()
    rc = lib.jegal_plan_column_tiles(cu.ctypes.data_as(C.POINTER(C.c_int32)), len(lengths), width, int(allow_split), None, 0, C.byref(n))
    if rc != 0:
        return rc, n.value, None
    buf = np.zeros((n.value, 12), dtype=np.uint32)
    rc = lib.jegal_plan_column_tiles(cu.ctypes.data_as(C.POINTER(C.c_int32)), len(lengths), width, int(allow_split),
                                     buf.ctypes.data_as(C.c_void_p), n.value, C.byref(n))
    return rc, n.value, buf


def _check_plan(lengths, width, allow_split):
    rc, n, tiles = _plan(lengths, width, allow_split)
    assert rc == 0
    cu = np.concatenate([[0], np.cumsum(lengths)])
    covered = np.zeros(cu[-1], dtype=np.int32)
    next_clip = 0
    next_seg = 0
    for t in tiles:
        row0, n_valid, clip0, part = (int(x) for x in t[:4])
        partial, seg0 = part & 1, clip0 + (part >> 1)
        assert seg0 == next_seg  # segments (whole clips and pieces of split clips) are numbered in tile order
        mask = t[4:]
        assert 1 <= n_valid <= width
        covered[row0:row0 + n_valid] += 1
        ends = [c * 32 + j for c in range(8) for j in range(32) if (int(mask[c]) >> j) & 1]
        assert ends and ends[-1] == n_valid - 1
        next_seg += len(ends)
        if partial:
            assert allow_split and len(ends) == 1 and lengths[clip0] > width
            assert cu[clip0] <= row0 and row0 + n_valid <= cu[clip0 + 1]
            if row0 + n_valid == cu[clip0 + 1]:
                next_clip = clip0 + 1
        else:
            assert clip0 == next_clip and row0 == cu[clip0]
            assert [row0 + e + 1 for e in ends] == list(cu[clip0 + 1: clip0 + 1 + len(ends)])  # whole clips, in order
            next_clip = clip0 + len(ends)
            if next_clip < len(lengths) and lengths[next_clip] <= width:
                assert n_valid + lengths[next_clip] > width  # greedy: the next clip did not fit
    assert next_clip == len(lengths) and (covered == 1).all()


def test_column_tile_planner_half_cuts():
    """flags = 3 (allow split + cut halves, the two-pass plan): no segment crosses column width/2, segments are
    numbered in tile order, and every clip is covered exactly once by its pieces."""
    rng = np.random.default_rng(1)
    for width in (128, 256):
        h = width // 2
        for _ in range(40):
            n = int(rng.integers(1, 60))
            lengths = rng.integers(1, 90, n)
            lengths = np.where(rng.random(n) < 0.1, rng.integers(width + 1, 3 * width, n), lengths)
            rc, nt, tiles = _plan(lengths, width, 3)
            assert rc == 0
            cu = np.concatenate([[0], np.cumsum(lengths)])
            covered = np.zeros(cu[-1], dtype=np.int32)
            next_seg, pieces = 0, np.zeros(n, dtype=np.int32)
            for t in tiles:
                row0, n_valid, clip0, part = (int(x) for x in t[:4])
                assert clip0 + (part >> 1) == next_seg
                ends = [c * 32 + j for c in range(8) for j in range(32) if (int(t[4 + c]) >> j) & 1]
                assert ends[-1] == n_valid - 1 and (n_valid <= h or (h - 1) in ends)
                start = 0
                for e in ends:  # every segment lies inside one clip and one half of the tile
                    c = int(np.searchsorted(cu, row0 + start, side="right") - 1)
                    assert cu[c] <= row0 + start and row0 + e + 1 <= cu[c + 1]
                    assert (start < h) == (e < h)
                    pieces[c] += 1
                    covered[row0 + start: row0 + e + 1] += 1
                    start = e + 1
                next_seg += len(ends)
            assert (covered == 1).all() and (pieces >= 1).all() and next_seg == pieces.sum()


def test_column_tile_planner_properties():
    rng = np.random.default_rng(0)
    for width in (128, 256):
        for _ in range(40):
            n = int(rng.integers(1, 60))
            lengths = rng.integers(1, width + 1, n)
            _check_plan(lengths, width, False)
            lengths = np.where(rng.random(n) < 0.2, rng.integers(width + 1, 3 * width, n), lengths)
            _check_plan(lengths, width, True)
        _check_plan([width] * 5, width, False)
        _check_plan([1] * 700, width, False)
        _check_plan([16] * 65536, width, False)


def test_column_tile_planner_rejects_overlong_clip_without_split():
    rc, bad, _ = _plan([10, 300, 20], 256, False)
    assert rc == -4 and bad == 1  # JEGAL_ERR_UNSUPPORTED, offending clip index


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_ctx_create_fails_cleanly_without_gpu():
    import ctypes as C
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.jegal_ctx_create(0, C.byref(h))
    assert rc == -2 and not h.value  # JEGAL_ERR_DEVICE, no context, no crash
    assert lib.jegal_last_error(None) == b"null ctx"
    assert lib.jegal_launch_count(None) == 0


def test_heatmap_renderer_jet_blend_and_png(tmp_path):
    """jegal_b200.render restates matplotlib's jet LUT and the overlay blend of utils/plot_heatmap.py:78-87."""
    import struct
    import zlib
    from jegal_b200 import render
    assert np.allclose(render.jet(np.array([0.0, 1.0])), [[0, 0, 0.5], [0.5, 0, 0]])
    mid = render.jet(np.array([0.5]))[0]
    assert abs(mid[1] - 1.0) < 1e-9 and abs(mid[0] - mid[2]) < 0.03  # green plateau, red ~ blue at the centre
    attn = np.array([[0.05, 0.9, 0.3], [0.95, 0.0, 0.79]], dtype=np.float32)
    rgb = render.heatmap_rgb(attn, thresh=0.8, alpha=0.6)
    over, base = render.jet(np.array([0.01])), render.jet(np.array([0.05]))
    want = 0.76 * (0.6 * over[0] + 0.4 * base[0]) + 0.24
    assert np.allclose(rgb[0, 0], want)
    path = render.render_heatmap(attn, ["hello", "world"], fname=str(tmp_path / "hm"), cell=4, labels=False)
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", raw[16:24])
    assert (h, w) == (2 * 4, 3 * 4 + 4 + 4) and b"hello world" in raw
    idat = raw[raw.index(b"IDAT") + 4: raw.index(b"IEND") - 8]
    pix = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)
    assert np.array_equal(pix[0, 0], np.rint(want * 255).astype(np.uint8))
    # with tick labels (plot_heatmap.py:89-92): margins for the words, the frame numbers and the colour-bar scale; the
    # heat-map cells sit unchanged behind the left margin, and the margins carry dark glyph pixels
    path = render.render_heatmap(attn, ["hello", "world"], fname=str(tmp_path / "hm2"), cell=16)
    raw = open(path, "rb").read()
    w2, h2 = struct.unpack(">II", raw[16:24])
    left = render.text_width("hello", 2) + 6
    assert w2 == left + 3 * 16 + 8 + 8 + render.text_width("0.0", 2) + 6 and h2 == 2 * 16 + 22
    idat = raw[raw.index(b"IDAT") + 4: raw.index(b"IEND") - 8]
    pix = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h2, 1 + 3 * w2)[:, 1:].reshape(h2, w2, 3)
    assert np.array_equal(pix[0, left], np.rint(want * 255).astype(np.uint8))
    assert (pix[:32, :left].sum(axis=2) == 0).sum() > 40 and (pix[32:, left:].sum(axis=2) == 0).sum() > 10
    img = np.ones((9, 14, 3))
    render.draw_text(img, 1, 1, "1!", 1)   # glyph check: '1' = 00 42 7F 40 00, '!' = 00 00 5F 00 00 (column bytes)
    cols = ["".join("#" if img[y, x, 0] < 0.5 else "." for y in range(1, 8)) for x in range(1, 12)]
    assert cols[2] == "#######" and cols[1] == ".#....#" and cols[8] == "#####.#"


def test_pack_rows_host_threaded_equals_concatenate():
    from jegal_b200.scoring import pack_rows_host
    rng = np.random.default_rng(4)
    arrs = [rng.standard_normal((int(L), 512)).astype(np.float16) for L in rng.integers(1, 400, size=700)]
    want = np.concatenate(arrs, axis=0)
    out = np.zeros_like(want)
    seen = []
    pack_rows_host(arrs, out, threads=4, on_group=lambda r0, r1: seen.append((r0, r1)))
    assert np.array_equal(out, want)
    assert seen[0][0] == 0 and seen[-1][1] == len(want) and all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    small = arrs[:3]
    out2 = np.zeros((sum(len(a) for a in small), 512), dtype=np.float16)
    pack_rows_host(small, out2)
    assert np.array_equal(out2, np.concatenate(small))


def test_wordlevel_range_logic_against_reference_golden(golden, monkeypatch):
    """The host half of jegal_b200.wordlevel (which rows belong to which word, invalid-sample bookkeeping)
    against the outputs of the reference's own methods; the K5 launch is replaced by a torch mean HERE ONLY,
    so the integer logic is covered without a GPU (the kernel itself is covered by the -m gpu suite)."""
    import torch
    from jegal_b200 import wordlevel
    from jegal_testutil import wordlevel_case

    def cpu_pool(emb, ranges, counts):
        x = emb.reshape(-1, emb.shape[-1])
        rows = [x[lo:hi].mean(dim=0) if hi - lo > 1 else x[lo] for lo, hi in ranges]
        return list(torch.split(torch.stack(rows), counts)) if rows else []

    monkeypatch.setattr(wordlevel, "_pool", cpu_pool)
    g = golden("wordlevel")
    text_emb, audio_emb, input_ids, offsets, text, bounds = wordlevel_case(g)
    wt, wa, inv = wordlevel.get_word_level_embs(text_emb, text, input_ids, offsets, audio_emb=audio_emb, word_boundaries=bounds)
    assert inv == list(g["invalid"]) and [len(x) for x in wt] == list(g["counts"])
    assert np.allclose(torch.cat(wt).numpy(), g["word_text"], atol=1e-6) and np.allclose(torch.cat(wa).numpy(), g["word_audio"], atol=1e-6)
    au, inv_a = wordlevel.get_audio_word_level_embs(audio_emb, bounds, list(inv))
    assert inv_a == list(g["audio_only_invalid"]) and np.allclose(torch.cat(au).numpy(), g["audio_only"], atol=1e-6)
    with pytest.raises(IndexError):  # an empty frame range fails like the reference's `[0]` on an empty tensor
        wordlevel.get_audio_word_level_embs(torch.zeros(1, 10, 256), [[["a", 100, 104], ["b", 120, 125]]])


def test_streaming_chunk_schedule():
    from jegal_b200.streaming import chunk_schedule
    cases = [(8192, 2048), (65536, 8192), (301, 64), (100, 64), (5, 64), (4096, 2048), (10000, 2048), (0, 64), (130, 64),
             # chunk sizes that are not multiples of 8 / 64 (3 GPUs, odd galleries): the ramps must still sum to a chunk
             (1000, 1000), (65536, 100), (65536, 8191), (21846, 2730), (50000, 6250), (999, 77), (257, 65), (12345, 1543)]
    rng = np.random.default_rng(0)
    cases += [(int(rng.integers(1, 200000)), int(rng.integers(1, 20000))) for _ in range(300)]
    for n, c in cases:
        sch = chunk_schedule(n, c)
        assert sum(sch) == n and all(x > 0 for x in sch) and max(sch, default=0) <= max(c, n if n < 128 else c)
        if len(sch) >= 8:  # ramps up at the start, down at the end
            assert sch[0] <= sch[2] <= sch[3] and sch[-1] <= sch[-3] <= sch[-4]
    assert chunk_schedule(300, 64, ramp=False) == [64, 64, 64, 64, 44]


def test_packed_index_round_trip(tmp_path):
    """The packed clip index (rows as stored + offsets + numpy-semantics temporal means + the info fields the scripts
    read) survives save / memory-mapped load, for both .pkl info flavours."""
    import pandas as pd
    from jegal_b200 import index, pkl_io
    rng = np.random.default_rng(3)
    d = tmp_path / "pkls"
    d.mkdir()
    clips = []
    for i in range(7):
        g = rng.standard_normal((int(rng.integers(3, 30)), 512)).astype(np.float16)
        c = rng.standard_normal((int(rng.integers(1, 9)), 512)).astype(np.float16)
        wb = [[f"w{k}", 2 * k, 2 * k + 1] for k in range(len(c))]
        info = pd.Series({"phrase": "p", "word_boundaries": str(wb), "target_word_boundary": str(wb[0]), "filename": f"v{i}/00001"}) \
            if i % 2 == 0 else {"fname": f"v{i}", "word_boundaries": wb, "text": "t"}
        pkl_io.write_pkl(str(d / f"v{i}__00001.pkl"), g, c, info)
        clips.append((g, c))
    stats = index.build_from_pkl_dir(str(d), str(tmp_path / "idx"))
    assert stats["n"] == 7 and stats["gesture_rows"] == sum(len(g) for g, _ in clips)
    ds = index.load_or_build("/nonexistent", str(tmp_path / "idx"))
    assert ds.n == 7 and ds.names == [f"v{i}__00001" for i in range(7)]
    for i, (g, c) in enumerate(clips):
        assert np.array_equal(np.asarray(ds.gesture.clip(i)), g) and np.array_equal(np.asarray(ds.content.clip(i)), c)
        assert np.array_equal(ds.gesture.mean[i], g.mean(axis=0)) and ds.gesture.mean.dtype == np.float16
        wb = ds.info[i]["word_boundaries"]
        assert pkl_io.parse_boundaries(wb)[0][0] == "w0"
    assert "target_word_boundary" in ds.info[0] and "text" in ds.info[1]


def test_clip_chunks_cover_every_clip_once():
    from types import SimpleNamespace
    from jegal_b200.streaming import clip_chunks
    rng = np.random.default_rng(5)
    for n, k in [(1, 8), (5, 8), (100, 8), (20000, 8), (999, 3), (64, 1)]:
        a = SimpleNamespace(n=n, lengths=rng.integers(1, 200, n).astype(np.int64), rows=SimpleNamespace(element_size=lambda: 2))
        b = SimpleNamespace(n=n, lengths=rng.integers(1, 20, n).astype(np.int64), rows=SimpleNamespace(element_size=lambda: 2))
        ch = clip_chunks([a, b], k)
        assert ch[0][0] == 0 and ch[-1][1] == n and all(x[1] == y[0] for x, y in zip(ch[:-1], ch[1:]))
        assert all(hi > lo for lo, hi in ch) and len(ch) <= k + 1
    assert clip_chunks([SimpleNamespace(n=0, lengths=np.zeros(0, np.int64), rows=SimpleNamespace(element_size=lambda: 2))], 4) == []


def test_balanced_schedule_covers_every_clip():
    from jegal_b200.streaming import balanced_schedule
    rng = np.random.default_rng(11)
    for n in [0, 1, 2, 127, 128, 300, 2047, 8192, 65536, 21846] + [int(x) for x in rng.integers(1, 300000, 200)]:
        sch = balanced_schedule(n)
        assert sum(sch) == n and all(x > 0 for x in sch) and len(sch) <= 16 and (n > 0 or sch == [])
        if sch:
            assert max(sch) - min(sch) <= 1
    assert balanced_schedule(8192) == [512] * 16 and balanced_schedule(300) == [150, 150]


def test_metrics_from_counts_rejects_nan_diagonal_and_empty():
    """A NaN ground-truth score (zero-length clip, NaN embedding) makes n_equal = 0 for its row: the reference would
    silently drop the row from every denominator; here it is an error, and so is an empty matrix."""
    from jegal_b200.scoring import _metrics_from_counts
    from jegal_b200._lib import JegalError
    m = _metrics_from_counts(np.array([0, 3, 7]), np.array([1, 1, 1]))
    assert m["R1"] == pytest.approx(1 / 3) and m["R5"] == pytest.approx(2 / 3) and m["MR"] == 4.0
    with pytest.raises(JegalError, match="NaN"):
        _metrics_from_counts(np.array([0, 3, 7]), np.array([1, 0, 1]))
    with pytest.raises(JegalError, match="empty"):
        _metrics_from_counts(np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
