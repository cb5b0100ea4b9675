"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FUNCTIONS.

Runs only in the build container (needs /root/reference, read-only).  The fixtures
it writes are committed, so nothing on the GPU box reads the reference.

    python oracle/make_golden.py

What is pinned (reference file:line):
  retrieval.npz  get_similarity_matrix + compute_metrics   evaluation/evaluate_retrieval.py:38-65
  spotting.npz   get_attn_matrix + get_spotting_acc        evaluation/evaluate_spotting.py:39-90
  asd.npz        load_feats + get_similarity_cos + evaluate_asd   evaluation/evaluate_asd.py:26-127
  simpool_tiles.npz  the frame x word COSINE TILES of clip pairs, produced by the reference's own
                 get_similarity_matrix (evaluate_retrieval.py:38-48: F.normalize + matmul) applied to a clip's
                 frame rows and word rows, and the four poolings of each tile taken with numpy.  This pins the
                 max-pool modes (which the released code does not contain) down to the final amax / mean.
  wordlevel.npz  JEGAL.get_word_level_embs + get_audio_word_level_embs   models/jegal.py:131-252
                 (method sources extracted with ast and executed: the module downloads XLM-R at import)
"""
import contextlib
import importlib.util
import io
import os
import pickle
import re
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("JEGAL_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from jegal_b200 import synth  # noqa: E402  (synthetic inputs only; no scoring code)


def import_ref(name: str, argv):
    """The evaluation scripts call parse_args() at import time: pre-set sys.argv."""
    old = sys.argv
    sys.argv = [name] + list(argv)
    try:
        spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, "evaluation", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = old
    return mod


def golden_retrieval():
    ref = import_ref("evaluate_retrieval", ["--path", "/nonexistent"])
    rng = np.random.default_rng(11)
    n = 24
    cs = synth.make_clipset(rng.integers(25, 80, size=n), rng.integers(4, 13, size=n), seed=101, a=0.05, b=0.08, sigma=1.0)
    g_list, c_list = cs.gesture_list(), cs.content_list()
    # load_feats (:30-31): temporal mean in the stored dtype (fp16)
    g_mean = [g.mean(axis=0).squeeze() for g in g_list]
    c_mean = [c.mean(axis=0).squeeze() for c in c_list]
    sim_c2g = ref.get_similarity_matrix(c_mean, g_mean).numpy()
    sim_g2c = ref.get_similarity_matrix(g_mean, c_mean).numpy()
    m_c2g = ref.compute_metrics(torch.from_numpy(sim_c2g))
    m_g2c = ref.compute_metrics(torch.from_numpy(sim_g2c))
    # a matrix with ties (integers): pins the over-counting behaviour of compute_metrics
    ties = rng.integers(0, 6, size=(40, 40)).astype(np.float32)
    m_ties = ref.compute_metrics(torch.from_numpy(ties))
    # random fp32 matrix, larger, so R25/R50 are non-trivial
    big = rng.standard_normal((120, 120)).astype(np.float32)
    m_big = ref.compute_metrics(torch.from_numpy(big))
    keys = ["R5", "R10", "R25", "R50", "MR"]
    np.savez_compressed(
        os.path.join(OUT, "retrieval.npz"),
        gest=cs.gest.numpy(), cont=cs.cont.numpy(), cu_t=cs.cu_t, cu_w=cs.cu_w,
        g_mean=np.stack(g_mean), c_mean=np.stack(c_mean),
        sim_c2g=sim_c2g, sim_g2c=sim_g2c,
        m_c2g=np.array([m_c2g[k] for k in keys]), m_g2c=np.array([m_g2c[k] for k in keys]),
        ties=ties, m_ties=np.array([m_ties[k] for k in keys]),
        big=big, m_big=np.array([m_big[k] for k in keys]),
    )
    print("retrieval:", m_c2g, m_ties)


def golden_spotting():
    ref = import_ref("evaluate_spotting", ["--path", "/nonexistent"])
    rng = np.random.default_rng(12)
    # the two sample-shaped clips (config 1) + 30 AVS-Spot-shaped ones
    cs1 = synth.cfg1_samples()
    n = 30
    lt = rng.integers(25, 90, size=n)
    lw = rng.integers(4, 13, size=n)
    cs2 = synth.make_clipset(lt, lw, seed=202, with_targets=True, b=0.6, sigma=1.3)
    gest, cont, wbs, rows, attn, decisions = [], [], [], [], [], []
    for cs in (cs1, cs2):
        for i in range(cs.n):
            gest.append(cs.gesture(i).numpy())
            cont.append(cs.content(i).numpy())
            wb = cs.boundaries[i]
            wbs.append(str(wb))
            tw = wb[int(cs.target_word[i])]
            rows.append(pd.Series({"phrase": " ".join(w[0] for w in wb), "word_boundaries": str(wb),
                                   "target_word_boundary": str(tw), "filename": f"vid/{i:05d}"}))
    targets = []
    for idx in range(len(gest)):
        a, words = ref.get_attn_matrix(idx, gest, cont, wbs)
        attn.append(np.ascontiguousarray(a))
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            acc = ref.get_spotting_acc([rows[idx]], [gest[idx]], [cont[idx]], [wbs[idx]])
        decisions.append(acc == 100.0)
        import ast
        twb = ast.literal_eval(rows[idx].target_word_boundary)
        targets.append([ast.literal_eval(wbs[idx]).index(twb), twb[1], twb[2]])
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        acc_all = ref.get_spotting_acc(rows, gest, cont, wbs)
    cu_t = np.concatenate([[0], np.cumsum([len(g) for g in gest])]).astype(np.int32)
    cu_w = np.concatenate([[0], np.cumsum([len(c) for c in cont])]).astype(np.int32)
    np.savez_compressed(
        os.path.join(OUT, "spotting.npz"),
        gest=np.concatenate(gest), cont=np.concatenate(cont), cu_t=cu_t, cu_w=cu_w,
        word_boundaries=np.array(wbs), targets=np.array(targets, dtype=np.int32),
        attn=np.concatenate([a.reshape(-1) for a in attn]),  # clip i: (W_i, T_i) row-major
        decisions=np.array(decisions), accuracy=np.float64(acc_all),
    )
    print("spotting: accuracy", acc_all, "decisions", int(np.sum(decisions)), "/", len(decisions))


def golden_asd():
    tmp = tempfile.mkdtemp(prefix="jegal_golden_asd_")
    n_groups, tracks = 10, 6
    ds = synth.cfg4_asd(n_groups, tracks, seed=303, t_range=(39, 70), a=0.05, b=0.08, sigma=1.0)
    cs = ds.clips
    names = [f"vid{i:04d}/{0:05d}" for i in range(cs.n)]
    for i in range(cs.n):
        with open(os.path.join(tmp, names[i].replace("/", "__") + ".pkl"), "wb") as f:
            pickle.dump({"gesture_emb": cs.gesture(i).numpy(), "content_emb": cs.content(i).numpy(),
                         "info": {"fname": names[i]}}, f)
    rows = []
    for g in range(n_groups):
        pos = g * tracks
        rows.append({"filename": names[pos], "neg_files": str([names[pos + k] for k in range(1, tracks)])})
    csv = os.path.join(tmp, "asd.csv")
    pd.DataFrame(rows).to_csv(csv, index=False)
    ref = import_ref("evaluate_asd", ["--path", tmp, "--file", csv])
    probs, preds = [], []
    for g in range(n_groups):
        pos = g * tracks
        _, q_cont = ref.load_feats(os.path.join(tmp, names[pos].replace("/", "__") + ".pkl"), load_content=True)
        gs = [ref.load_feats(os.path.join(tmp, names[pos + k].replace("/", "__") + ".pkl"), load_content=False)
              for k in range(tracks)]
        allg = torch.cat(gs)
        pr, pd_ = [], []
        for p in (2, 4, 6):
            s = ref.get_similarity_cos(q_cont, allg[:p])
            pr.append(np.pad(s, (0, 6 - p)))
            pd_.append(int(np.argmax(s)))
        probs.append(pr)
        preds.append(pd_)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        ref.evaluate_asd(ref.read_data(csv))
    acc = [float(x) for x in re.findall(r"Acc: ([0-9.]+)", buf.getvalue())]
    np.savez_compressed(
        os.path.join(OUT, "asd.npz"),
        gest=cs.gest.numpy(), cont=cs.cont.numpy(), cu_t=cs.cu_t, cu_w=cs.cu_w,
        tracks=np.int32(tracks), probs=np.array(probs, dtype=np.float32), preds=np.array(preds, dtype=np.int32),
        accuracy=np.array(acc),
    )
    print("asd: accuracies", acc)


def golden_simpool_tiles():
    """Every T x W tile = ref.get_similarity_matrix(frames of clip i, words of clip j): the reference's own
    normalise + matmul, on per-frame / per-word rows instead of per-clip means.  The pooled scores are plain
    numpy reductions of those reference-produced tiles."""
    ref = import_ref("evaluate_retrieval", ["--path", "/nonexistent"])
    rng = np.random.default_rng(21)
    n = 10
    lt = np.concatenate([rng.integers(25, 140, size=n - 2), [1, 300]])   # incl. a one-frame clip and a > 256-frame clip
    lw = np.concatenate([rng.integers(4, 30, size=n - 2), [1, 70]])      # incl. a one-word clip and a > 64-word clip
    cs = synth.make_clipset(lt, lw, seed=202, a=0.3, b=1.0, sigma=1.0)
    g_list, c_list = cs.gesture_list(), cs.content_list()
    tiles, off = [], [0]
    pooled = {m: np.zeros((n, n), dtype=np.float32) for m in ("mean_mean", "max_t_mean_w", "max_w_mean_t", "max_max")}
    for i in range(n):
        for j in range(n):
            t = ref.get_similarity_matrix(g_list[i], c_list[j]).numpy()  # (T_i, W_j) cosines, fp32
            assert t.shape == (len(g_list[i]), len(c_list[j]))
            tiles.append(t.reshape(-1))
            off.append(off[-1] + t.size)
            pooled["mean_mean"][i, j] = t.mean(dtype=np.float32)
            pooled["max_t_mean_w"][i, j] = t.max(axis=0).mean(dtype=np.float32)
            pooled["max_w_mean_t"][i, j] = t.max(axis=1).mean(dtype=np.float32)
            pooled["max_max"][i, j] = t.max()
    np.savez_compressed(os.path.join(OUT, "simpool_tiles.npz"), gest=cs.gest.numpy(), cont=cs.cont.numpy(), cu_t=cs.cu_t,
                        cu_w=cs.cu_w, tiles=np.concatenate(tiles), tile_off=np.array(off, dtype=np.int64),
                        **{"pooled_" + k: v for k, v in pooled.items()})
    print("simpool_tiles:", n * n, "tiles,", off[-1], "cosines")


def ref_methods(names):
    """models/jegal.py cannot be imported offline (AutoTokenizer.from_pretrained at :13-14), so the wanted
    methods are cut out of its source with ast and compiled as plain functions (self is unused by them);
    the module-level `tokenizer` they read is replaced by XLM-R's three special-token ids."""
    import ast
    import types

    src = open(os.path.join(REF, "models", "jegal.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "tokenizer": types.SimpleNamespace(cls_token_id=0, sep_token_id=2, pad_token_id=1)}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, os.path.join(REF, "models", "jegal.py"), "exec"), ns)
    return [ns[n] for n in names]


def wordlevel_inputs(seed=7, B=7, L=26, Tf=64, D=256):
    """Tokenizer-shaped synthetic batch: <s> sub-words </s> <pad>..., offsets (0, n) for word starts."""
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    text_emb = torch.randn(B, L, D, generator=g)
    audio_emb = torch.randn(B, Tf, D, generator=g)
    input_ids = np.ones((B, L), dtype=np.int64)  # pad = 1
    offsets = np.zeros((B, L, 2), dtype=np.int64)
    text, bounds = [], []
    for b in range(B):
        n_words = int(rng.integers(1, 8))
        pos = 1
        input_ids[b, 0] = 0
        words = []
        for w in range(n_words):
            n_sub = int(rng.integers(1, 4))
            if pos + n_sub >= L - 1:
                break
            c = 0
            for k in range(n_sub):
                ln = int(rng.integers(1, 5))
                input_ids[b, pos] = int(rng.integers(5, 1000))
                offsets[b, pos] = (c, c + ln)
                c += ln
                pos += 1
            words.append(f"w{b}_{w}")
        input_ids[b, pos] = 2
        if b == 3:
            words = words + ["extra1", "extra2"]  # more words than word starts -> invalid sample (:162-166)
        text.append(words)
        t0 = int(rng.integers(0, 200))
        cuts = np.sort(rng.choice(np.arange(1, Tf - 1), size=len(words), replace=False)) if len(words) < Tf - 2 else None
        wb, s0 = [], 0
        for w, word in enumerate(words):
            e0 = int(cuts[w])
            wb.append([word, t0 + s0, t0 + e0])  # end-inclusive, next word starts on the same frame (overlap)
            s0 = e0
        bounds.append(wb)
    return text_emb, audio_emb, torch.from_numpy(input_ids), torch.from_numpy(offsets), text, bounds


def golden_wordlevel():
    get_word, get_audio = ref_methods(["get_word_level_embs", "get_audio_word_level_embs"])
    text_emb, audio_emb, input_ids, offsets, text, bounds = wordlevel_inputs()
    wt, wa, inv = get_word(None, text_emb, text, input_ids, offsets, audio_emb=audio_emb, word_boundaries=bounds)
    wt2, wa2, inv2 = get_word(None, text_emb, text, input_ids, offsets, audio_emb=None, word_boundaries=None)
    assert wa2 == [] and inv2 == inv and all(torch.equal(a, b) for a, b in zip(wt, wt2))
    au, inv_a = get_audio(None, audio_emb, bounds, list(inv))
    # fp16 inputs (the .pkl dtype): the reference's mean keeps the dtype
    wt_h, wa_h, _ = get_word(None, text_emb.half().float().half(), text, input_ids, offsets,
                             audio_emb=audio_emb.half(), word_boundaries=bounds)
    np.savez_compressed(
        os.path.join(OUT, "wordlevel.npz"),
        text_emb=text_emb.numpy(), audio_emb=audio_emb.numpy(), input_ids=input_ids.numpy(), offsets=offsets.numpy(),
        n_words=np.int32([len(t) for t in text]),
        bounds=np.int64([[wb[1], wb[2]] for b in bounds for wb in b]),
        counts=np.int32([len(x) for x in wt]), invalid=np.int32(inv),
        word_text=torch.cat(wt).numpy(), word_audio=torch.cat(wa).numpy(),
        audio_only=torch.cat(au).numpy(), audio_only_counts=np.int32([len(x) for x in au]), audio_only_invalid=np.int32(inv_a),
        word_text_f16=torch.cat(wt_h).numpy(), word_audio_f16=torch.cat(wa_h).numpy(),
    )
    print("wordlevel: words per valid clip", [len(x) for x in wt], "invalid", inv)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    golden_retrieval()
    golden_spotting()
    golden_asd()
    golden_simpool_tiles()
    golden_wordlevel()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
