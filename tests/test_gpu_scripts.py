"""The drop-in evaluation scripts on synthetic .pkl directories in BOTH info
flavours, against the oracle run on the same files."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.fixture(scope="module")
def pkl_dir(tmp_path_factory):
    from jegal_b200 import pkl_io, synth
    d = tmp_path_factory.mktemp("pkls")
    cs = synth.make_clipset(np.random.default_rng(1).integers(25, 120, 48), np.random.default_rng(2).integers(4, 13, 48),
                            seed=77, a=0.05, b=0.08, with_targets=True)
    names = [f"vid{i:03d}/{i % 3:05d}" for i in range(cs.n)]
    for i in range(cs.n):
        wb = cs.boundaries[i]
        if i % 2 == 0:   # extract_jegal_embs.py flavour: pandas row, string fields
            info = pd.Series({"phrase": " ".join(w[0] for w in wb), "word_boundaries": str(wb),
                              "target_word_boundary": str(wb[int(cs.target_word[i])]), "filename": names[i]})
        else:            # same keys, list-valued boundaries (inference_embs.py style lists)
            info = pd.Series({"phrase": " ".join(w[0] for w in wb), "word_boundaries": wb,
                              "target_word_boundary": wb[int(cs.target_word[i])], "filename": names[i]})
        pkl_io.write_pkl(os.path.join(d, pkl_io.clip_pkl_name(names[i])), cs.gesture(i).numpy(), cs.content(i).numpy(), info)
    rows = []
    for g in range(8):
        rows.append({"filename": names[g * 6], "neg_files": str([names[g * 6 + k] for k in range(1, 6)])})
    csv = os.path.join(d, "asd.csv")
    pd.DataFrame(rows).to_csv(csv, index=False)
    return str(d), cs, names, csv


def test_evaluate_retrieval_script(pkl_dir):
    d, cs, names, _ = pkl_dir
    out = run("evaluate_retrieval.py", "--path", d)
    gest, cont = cs.gesture_list(), cs.content_list()
    s = oracle.get_similarity_matrix([oracle.mean_pool(g) for g in gest], [oracle.mean_pool(c) for c in cont]).numpy()
    lines = [l for l in out.splitlines() if l.startswith("R@5")]
    assert len(lines) == 2

    def fmt(m):
        return 'R@5: {:.2f} - R@10: {:.2f} - R@25: {:.2f} - R@50: {:.2f} | Median R: {:.1f}'.format(
            m['R5'] * 100, m['R10'] * 100, m['R25'] * 100, m['R50'] * 100, m['MR'])
    near = (np.abs(s - np.diag(s)[:, None]) < 4e-3).sum() - len(s)
    if near == 0:  # no decision inside the tolerance band: printed metrics must be identical
        assert lines[0] == fmt(oracle.compute_metrics(s.T)) and lines[1] == fmt(oracle.compute_metrics(s))


def test_evaluate_spotting_script(pkl_dir):
    d, cs, names, _ = pkl_dir
    out = run("evaluate_spotting.py", "--path", d)
    acc = float([l for l in out.splitlines() if l.startswith("Word Spotting Accuracy")][0].split(":")[1])
    hits, fragile = 0, 0
    for i in range(cs.n):
        a = oracle.get_attn_matrix(cs.gesture(i).numpy(), cs.content(i).numpy())
        w = int(cs.target_word[i])
        _, s0, e0 = cs.boundaries[i][w]
        w_first = cs.boundaries[i].index(cs.boundaries[i][w])
        pred, score, ok = oracle.spot_decision(a, w_first, s0, e0)
        hits += ok
        fragile += abs(score - 0.5) < 1.5e-2 or np.sort(a[w_first])[-1] - np.sort(a[w_first])[-2] < 1.5e-2
    assert abs(acc - 100.0 * hits / cs.n) <= 100.0 * fragile / cs.n + 1e-9


def test_evaluate_asd_script(pkl_dir):
    d, cs, names, csv = pkl_dir
    out = run("evaluate_asd.py", "--path", d, "--file", csv)
    accs = [float(l.split("Acc:")[1]) for l in out.splitlines() if "spk: Correct" in l]
    assert len(accs) == 3
    ref = np.zeros(3)
    fragile = np.zeros(3)
    for g in range(8):
        cos = [float(np.dot(oracle.normalize_rows(oracle.asd_mean_emb(cs.content(g * 6).numpy()))[0],
                            oracle.normalize_rows(oracle.asd_mean_emb(cs.gesture(g * 6 + k).numpy()))[0])) for k in range(6)]
        for pi, p in enumerate((2, 4, 6)):
            ref[pi] += int(np.argmax(cos[:p]) == 0)
            srt = np.sort(cos[:p])
            fragile[pi] += (srt[-1] - srt[-2]) < 4e-3
    for pi in range(3):
        assert abs(accs[pi] - ref[pi] / 8) <= fragile[pi] / 8 + 1e-3


def test_plot_heatmap_script_and_index(pkl_dir, tmp_path):
    d, cs, names, _ = pkl_dir
    from jegal_b200 import index, pkl_io
    f = os.path.join(d, pkl_io.clip_pkl_name(names[0]))
    out = run("plot_heatmap.py", "--path", f, "--fname", str(tmp_path / "hm"))
    a = np.load(str(tmp_path / "hm.npy"))
    assert open(str(tmp_path / "hm.png"), "rb").read(8) == b"\x89PNG\r\n\x1a\n"  # drawn without matplotlib
    ref = oracle.get_attn_matrix(cs.gesture(0).numpy(), cs.content(0).numpy(), normalize=False)
    assert a.shape == ref.shape and np.abs(a - ref).max() < 1.5e-2
    stats = index.build_from_pkl_dir(d, str(tmp_path / "idx"))
    assert stats["n"] == cs.n
    gi = index.ClipIndex.load(str(tmp_path / "idx.gesture"))
    j = gi.names.index(pkl_io.clip_pkl_name(names[5])[:-4])
    assert np.array_equal(np.asarray(gi.clip(j)), cs.gesture(5).numpy())
    from jegal_b200 import scoring
    ci = index.ClipIndex.load(str(tmp_path / "idx.content"))
    s = scoring.clip_similarity_matrix(gi.to_packed(), ci.to_packed())
    order = [gi.names.index(pkl_io.clip_pkl_name(n)[:-4]) for n in names]
    ref_s = oracle.get_similarity_matrix([oracle.mean_pool(g) for g in cs.gesture_list()],
                                         [oracle.mean_pool(c) for c in cs.content_list()]).numpy()
    assert np.abs(s[np.ix_(order, order)] - ref_s).max() < 2e-3


def test_scripts_with_packed_index(pkl_dir, tmp_path):
    """--index PREFIX: the first run builds the packed index from the .pkl directory, later runs read only the index
    (rows streamed host->device overlapped with K3; retrieval / ASD from the stored temporal means) -- and print
    exactly what the .pkl route prints."""
    d, cs, names, csv = pkl_dir
    prefix = str(tmp_path / "avs")
    for script, extra in (("evaluate_spotting.py", []), ("evaluate_retrieval.py", []), ("evaluate_asd.py", ["--file", csv])):
        plain = run(script, "--path", d, *extra)
        built = run(script, "--path", d, "--index", prefix, *extra)
        assert os.path.exists(prefix + ".meta.json") and os.path.exists(prefix + ".gesture.rows.npy")
        again = run(script, "--path", "/nonexistent", "--index", prefix, *extra)  # the .pkl files are not needed any more
        pick = lambda out: [l for l in out.splitlines() if l.startswith(("R@5", "Word Spotting", "2 spk", "4 spk", "6 spk"))]
        assert pick(built) == pick(again) and len(pick(built)) >= 1
        if script == "evaluate_spotting.py":  # same kernel, same operands: identical
            assert pick(built) == pick(plain)
        else:  # clip means from numpy (index) vs K0 (pkl route): equal up to summation order
            for a, b in zip(pick(built), pick(plain)):
                na, nb = [float(x) for x in __import__("re").findall(r"[-+]?\d+\.\d+", a)], [float(x) for x in __import__("re").findall(r"[-+]?\d+\.\d+", b)]
                assert len(na) == len(nb) and all(abs(x - y) <= 2.5 for x, y in zip(na, nb)), (a, b)


def test_streamed_spotting_and_asd_equal_resident(pkl_dir):
    d, cs, names, _ = pkl_dir
    from jegal_b200 import index, scoring, streaming
    gi, ci = index.ClipIndex.from_clips(cs.gesture_list()), index.ClipIndex.from_clips(cs.content_list())
    g, c = streaming.HostClips.from_index(gi), streaming.HostClips.from_index(ci)
    tw = cs.target_word
    st = np.array([cs.boundaries[i][int(tw[i])][1] for i in range(cs.n)])
    en = np.array([cs.boundaries[i][int(tw[i])][2] for i in range(cs.n)])
    win = (np.maximum(st - 9, 0), en + 9)
    for n_chunks in (1, 3, 8):
        r = streaming.spot_streamed(g, c, tw, windows=win, n_chunks=n_chunks, want_heat=True)
        ref = scoring.spot_batch(cs.gesture_list(), cs.content_list(), tw, windows=win)
        assert np.array_equal(r["pred_frame"], ref["pred_frame"]) and np.array_equal(r["correct"], ref["correct"])
        assert np.array_equal(r["pred_score"], ref["pred_score"]) and np.array_equal(r["heat"], np.concatenate(ref["heat"]))
    # fp32 storage: no fused path, every chunk takes one K0 pass (normalise + cast to bf16) in front of K3
    g32 = streaming.HostClips(torch.from_numpy(np.concatenate(cs.gesture_list()).astype(np.float32)), np.diff(cs.cu_t))
    c32 = streaming.HostClips(torch.from_numpy(np.concatenate(cs.content_list()).astype(np.float32)), np.diff(cs.cu_w))
    r32 = streaming.spot_streamed(g32, c32, tw, windows=win, n_chunks=3, want_heat=True)
    assert np.abs(r32["heat"] - np.concatenate(ref["heat"])).max() < 1.5e-2 and (r32["pred_frame"] == ref["pred_frame"]).mean() > 0.9
    pair_g = np.arange(cs.n, dtype=np.int32)
    pair_c = np.repeat(np.arange(0, cs.n, 6, dtype=np.int32), 6)
    # pairs listed in an order that does not follow the gesture clips: the per-chunk selection takes the gather route
    perm = np.random.default_rng(3).permutation(cs.n // 6)
    pg_s = pair_g.reshape(-1, 6)[perm].reshape(-1)
    pc_s = pair_c.reshape(-1, 6)[perm].reshape(-1)
    ref_s = scoring.asd_batch(cs.content_list(), cs.gesture_list(), pg_s, pc_s, 6, mode="max_t_mean_w")
    r_s = streaming.asd_streamed(g, c, pg_s, pc_s, 6, mode="max_t_mean_w", n_chunks=4)
    assert np.array_equal(r_s["scores"], ref_s["scores"])
    for mode in ("reference", "max_t_mean_w"):
        ref = scoring.asd_batch(cs.content_list(), cs.gesture_list(), pair_g, pair_c, 6, mode=mode)
        for n_chunks in (1, 5):
            r = streaming.asd_streamed(g, c, pair_g, pair_c, 6, mode=mode, n_chunks=n_chunks)
            assert np.array_equal(r["scores"], ref["scores"]) and all(np.array_equal(r["pred"][P], ref["pred"][P]) for P in (2, 4, 6))
