"""The CPU oracle against fixtures produced by EXECUTING the reference's own functions
(oracle/make_golden.py, run in the build container).  Bit-exact for integers/decisions; fp32
values to 1e-6 (same torch ops, same order)."""
import ast

import numpy as np
import pytest
import torch

from oracle import oracle
from jegal_testutil import split

KEYS = ["R5", "R10", "R25", "R50", "MR"]


def test_retrieval_similarity_matrix(golden):
    g = golden("retrieval")
    s = oracle.get_similarity_matrix(list(g["c_mean"]), list(g["g_mean"])).numpy()
    np.testing.assert_allclose(s, g["sim_c2g"], atol=1e-6)
    s2 = oracle.get_similarity_matrix(g["g_mean"], g["c_mean"]).numpy()
    np.testing.assert_allclose(s2, g["sim_g2c"], atol=1e-6)


def test_retrieval_mean_pool_matches_load_feats(golden):
    g = golden("retrieval")
    gm = np.stack([oracle.mean_pool(c) for c in split(g["gest"], g["cu_t"])])
    assert gm.dtype == np.float16 and np.array_equal(gm, g["g_mean"])


@pytest.mark.parametrize("name,mat", [("m_c2g", "sim_c2g"), ("m_g2c", "sim_g2c"), ("m_ties", "ties"), ("m_big", "big")])
def test_compute_metrics(golden, name, mat):
    g = golden("retrieval")
    m = oracle.compute_metrics(g[mat])
    assert [m[k] for k in KEYS] == list(g[name])
    # the count formulation used on the GPU is the same multiset of ranks
    m2 = oracle.metrics_from_counts(*oracle.rank_counts(g[mat]))
    assert m2 == m


def test_meanmean_refnorm_identity(golden):
    """mean/mean pooling of the T x W tile x 1/(||mean_g|| ||mean_c||) == the reference's
    cosine of mean-pooled clips.  Pins the only sim-pool mode the reference contains."""
    g = golden("retrieval")
    gest, cont = split(g["gest"], g["cu_t"]), split(g["cont"], g["cu_w"])
    pooled = oracle.simpool_allpairs(gest, cont, "mean_mean", normalize=False)
    s = pooled * oracle.refnorm_scales(gest)[:, None] * oracle.refnorm_scales(cont)[None, :]
    # the reference rounds each mean vector to fp16 (numpy mean of an fp16 array): 5e-4 relative per component
    np.testing.assert_allclose(s, g["sim_g2c"], atol=2e-4)


def test_spotting_attn_and_decisions(golden):
    g = golden("spotting")
    gest, cont = split(g["gest"], g["cu_t"]), split(g["cont"], g["cu_w"])
    off, correct = 0, []
    for i, (ge, co) in enumerate(zip(gest, cont)):
        a = oracle.get_attn_matrix(ge, co)
        ref = g["attn"][off:off + a.size].reshape(a.shape)
        off += a.size
        np.testing.assert_allclose(a, ref, atol=1e-6)
        widx, start, end = [int(x) for x in g["targets"][i]]
        wb = ast.literal_eval(str(g["word_boundaries"][i]))
        assert wb[widx][1] == start and wb[widx][2] == end
        correct.append(oracle.spot_decision(a, widx, start, end)[2])
    assert np.array_equal(np.array(correct), g["decisions"])
    assert np.mean(correct) * 100 == pytest.approx(float(g["accuracy"]))


def test_plot_heatmap_variant_is_unnormalised(golden):
    """utils/plot_heatmap.py:51-57 == evaluate_spotting's matrix without F.normalize; on unit-norm
    stored rows both agree to the fp16 storage error."""
    g = golden("spotting")
    ge, co = split(g["gest"], g["cu_t"])[0], split(g["cont"], g["cu_w"])[0]
    a = oracle.get_attn_matrix(ge, co, normalize=False)
    b = oracle.get_attn_matrix(ge, co, normalize=True)
    assert np.abs(a - b).max() < 5e-3


def test_asd(golden):
    g = golden("asd")
    tracks = int(g["tracks"])
    gest, cont = split(g["gest"], g["cu_t"]), split(g["cont"], g["cu_w"])
    n_groups = len(gest) // tracks
    hits = np.zeros(3)
    for grp in range(n_groups):
        pos = grp * tracks
        q = oracle.asd_mean_emb(cont[pos])
        cands = [oracle.asd_mean_emb(gest[pos + k]) for k in range(tracks)]
        import torch
        allg = torch.cat(cands)
        for pi, p in enumerate((2, 4, 6)):
            s = oracle.get_similarity_cos(q, allg[:p])
            np.testing.assert_allclose(s, g["probs"][grp, pi, :p], atol=1e-6)
            assert int(np.argmax(s)) == int(g["preds"][grp, pi])
        pred = oracle.asd_predict(cont[pos], [gest[pos + k] for k in range(tracks)])
        assert pred == list(g["preds"][grp])
        hits += (np.array(pred) == 0)
    np.testing.assert_allclose(hits / n_groups, g["accuracy"], atol=1e-3)


def test_simpool_vectorised_equals_loop():
    rng = np.random.default_rng(0)
    gest = [rng.standard_normal((int(t), 512)).astype(np.float16) for t in rng.integers(1, 40, 6)]
    cont = [rng.standard_normal((int(w), 512)).astype(np.float16) for w in rng.integers(1, 9, 5)]
    for mode in oracle.POOL_MODES:
        a = oracle.simpool_allpairs_loop(gest, cont, mode)
        b = oracle.simpool_allpairs(gest, cont, mode)
        np.testing.assert_allclose(a, b, atol=1e-6)


def test_topk_ties_prefer_lower_index():
    x = np.array([[1.0, 3.0, 3.0, 2.0, 3.0]], dtype=np.float32)
    v, i = oracle.topk(x, 4)
    assert i.tolist() == [[1, 2, 4, 3]] and v.tolist() == [[3.0, 3.0, 3.0, 2.0]]


# ------------------------------------------------------------------ word-level pooling (models/jegal.py:131-252)
def test_wordlevel_oracle_equals_reference_methods(golden):
    import torch
    from jegal_testutil import wordlevel_case

    g = golden("wordlevel")
    text_emb, audio_emb, input_ids, offsets, text, bounds = wordlevel_case(g)
    wt, wa, inv = oracle.word_level_embs(text_emb, text, input_ids, offsets, audio_emb=audio_emb, word_boundaries=bounds)
    assert inv == list(g["invalid"]) and [len(x) for x in wt] == list(g["counts"])
    assert np.array_equal(torch.cat(wt).numpy(), g["word_text"]) and np.array_equal(torch.cat(wa).numpy(), g["word_audio"])
    au, inv_a = oracle.audio_word_level_embs(audio_emb, bounds, list(inv))
    assert inv_a == list(g["audio_only_invalid"]) and [len(x) for x in au] == list(g["audio_only_counts"])
    assert np.array_equal(torch.cat(au).numpy(), g["audio_only"])
    wt_h, wa_h, _ = oracle.word_level_embs(text_emb.half(), text, input_ids, offsets, audio_emb=audio_emb.half(),
                                           word_boundaries=bounds)
    assert np.array_equal(torch.cat(wt_h).numpy(), g["word_text_f16"])
    assert np.array_equal(torch.cat(wa_h).numpy(), g["word_audio_f16"])


def test_simpool_tiles_golden_pins_the_max_pool_modes(golden):
    """tests/golden/simpool_tiles.npz holds T x W cosine tiles produced by the REFERENCE's get_similarity_matrix
    (evaluate_retrieval.py:38-48) on per-frame / per-word rows, and numpy poolings of them: the oracle's tile and
    all four pooling modes (three of which the released code does not contain) must reproduce them."""
    g = golden("simpool_tiles")
    gest, cont = split(g["gest"], g["cu_t"]), split(g["cont"], g["cu_w"])
    n = len(gest)
    off = g["tile_off"]
    for i in range(n):
        for j in range(n):
            t = g["tiles"][off[i * n + j]:off[i * n + j + 1]].reshape(len(gest[i]), len(cont[j]))
            mine = oracle.cos_tile(gest[i], cont[j]).numpy()
            assert np.abs(mine - t).max() < 2e-6
            for mode in oracle.POOL_MODES:
                assert abs(oracle.pool_tile(torch.from_numpy(t), mode) - g["pooled_" + mode][i, j]) < 1e-6
    for mode in oracle.POOL_MODES:  # the vectorised restatement the GPU tests compare against
        assert np.abs(oracle.simpool_allpairs(gest, cont, mode) - g["pooled_" + mode]).max() < 5e-6
        assert np.abs(oracle.simpool_allpairs_loop(gest, cont, mode) - g["pooled_" + mode]).max() < 5e-6
