"""BASELINE.json configs 2-4 at FULL size (config 5 is in test_gpu_parity.py), plus randomised
ragged shapes.  Where the CPU oracle is too slow for the full tensor, the oracle arithmetic
(fp32 torch: normalize + mm + segment max/mean) is evaluated on the GPU box's device."""
import numpy as np
import pytest
import torch

from jegal_testutil import gap_aware_equal
from oracle import oracle

pytestmark = pytest.mark.gpu
TOL, PROB_TOL = 2e-3, 1.5e-2


@pytest.fixture(scope="module")
def dev():
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def test_cfg2_avs_ret_full(dev):
    """1000 ragged clips (T 25-200, W 4-40): full N x N both directions, recall@k / MedR."""
    from jegal_b200 import ops, scoring, synth
    cs = synth.cfg2_retrieval(1000, a=0.03, b=0.05)
    gest, cont = cs.gesture_list(), cs.content_list()
    # (i) the reference's own score (cosine of mean-pooled clips) on the CPU oracle, full size
    ref = oracle.get_similarity_matrix([oracle.mean_pool(g) for g in gest], [oracle.mean_pool(c) for c in cont]).numpy()
    got = scoring.clip_similarity_matrix(gest, cont)
    assert np.abs(got - ref).max() < TOL
    fused = scoring.score_allpairs(gest, cont, "mean_mean", refnorm=True)
    assert np.abs(fused - ref).max() < TOL
    c2g, g2c = scoring.retrieval_metrics(gest, cont)
    for got_m, mat in ((g2c, ref), (c2g, ref.T)):
        ref_m = oracle.compute_metrics(mat)
        d = np.diag(mat)[:, None]
        fragile = int(((np.abs(mat - d) < 2 * TOL).sum(1) > 1).sum())  # rows whose rank may move
        for key in ("R1", "R5", "R10", "R25", "R50"):
            assert abs(got_m[key] - ref_m[key]) <= fragile / len(mat) + 1e-12, key
        assert abs(got_m["MR"] - ref_m["MR"]) <= 1.0 + fragile
    # (ii) the irreducible pooling modes against the fp32 oracle arithmetic evaluated on the device
    for mode in ("max_t_mean_w", "max_w_mean_t", "max_max"):
        ref_p = oracle.simpool_allpairs(gest, cont, mode, device="cuda")
        got_p = scoring.score_allpairs(gest, cont, mode)
        assert np.abs(got_p - ref_p).max() < TOL, mode
        # top-1 of every row identical wherever the oracle's margin exceeds the tolerance
        bad = gap_aware_equal(got_p.argmax(1), ref_p.argmax(1), lambda n, j: ref_p[n, j], TOL)
        assert not bad, (mode, bad[:5])


def test_cfg3_avs_spot_full(dev):
    """20 000 ragged clips: target-word heatmap rows, argmax frame, spot decision."""
    from jegal_b200 import scoring, synth
    cs = synth.cfg3_spotting(20000, b=0.6, sigma=1.3)
    gest, cont = cs.gesture_list(), cs.content_list()
    starts = np.array([cs.boundaries[i][int(cs.target_word[i])][1] for i in range(cs.n)])
    ends = np.array([cs.boundaries[i][int(cs.target_word[i])][2] for i in range(cs.n)])
    lo, hi = np.maximum(starts - 9, 0), ends + 9
    r = scoring.spot_batch(scoring.PackedClips.from_packed(cs.gest, cs.cu_t), scoring.PackedClips.from_packed(cs.cont, cs.cu_w),
                           cs.target_word, windows=(lo, hi))
    n_ok = n_fragile = 0
    for i in range(cs.n):
        a = oracle.get_attn_matrix(gest[i], cont[i])
        row = a[int(cs.target_word[i])]
        assert np.abs(r["heat"][i] - row).max() < PROB_TOL, i
        pred, score, ok = oracle.spot_decision(a, int(cs.target_word[i]), int(starts[i]), int(ends[i]))
        srt = np.sort(row)
        fragile = (srt[-1] - srt[-2] < PROB_TOL) or abs(score - 0.5) < PROB_TOL
        if not fragile:
            assert r["pred_frame"][i] == pred and bool(r["correct"][i]) == ok, i
        n_fragile += fragile
        n_ok += ok
    assert abs(float(r["correct"].mean()) - n_ok / cs.n) <= n_fragile / cs.n + 1e-12
    assert 0.2 < n_ok / cs.n < 0.999  # the decisions are not degenerate


def test_cfg4_avs_asd_full(dev):
    """10 000 groups x 4 candidate tracks: cosine of mean-pooled clips, argmax track."""
    from jegal_b200 import scoring, synth
    ds = synth.cfg4_asd(10000, 4, a=0.03, b=0.05)
    cs = ds.clips
    r = scoring.asd_batch(scoring.PackedClips.from_packed(cs.cont, cs.cu_w), scoring.PackedClips.from_packed(cs.gest, cs.cu_t),
                          ds.pair_gest, ds.pair_cont, tracks=4, prefixes=(2, 4))
    gm = torch.stack([oracle.asd_mean_emb(g)[0] for g in cs.gesture_list()])
    cm = torch.stack([oracle.asd_mean_emb(c)[0] for c in cs.content_list()])
    cos = torch.nn.functional.cosine_similarity(cm[torch.from_numpy(ds.pair_cont.astype(np.int64))],
                                                gm[torch.from_numpy(ds.pair_gest.astype(np.int64))], dim=1).numpy().reshape(-1, 4)
    assert np.abs(r["scores"] - cos).max() < TOL
    for P in (2, 4):
        ref_pred = cos[:, :P].argmax(1)
        bad = gap_aware_equal(r["pred"][P], ref_pred, lambda n, j: cos[n, j], TOL)
        assert not bad
        srt = np.sort(cos[:, :P], axis=1)
        fragile = int((srt[:, -1] - srt[:, -2] < TOL).sum())
        assert abs(r["acc"][P] - float((ref_pred == 0).mean())) <= fragile / len(cos) + 1e-12
    assert 0.3 < r["acc"][4] < 0.999


def test_random_ragged_shapes(dev):
    """Randomised ragged layouts incl. length-1 clips, lengths around the 8/32/128/256 tile edges."""
    from jegal_b200 import scoring
    rng = np.random.default_rng(2024)
    edge = np.array([1, 2, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 200, 255, 256])
    for trial in range(12):
        ng, nc = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        lt = np.where(rng.random(ng) < 0.5, rng.choice(edge, ng), rng.integers(1, 257, ng))
        lw = np.where(rng.random(nc) < 0.5, rng.choice(edge[edge <= 64], nc), rng.integers(1, 65, nc))

        def clips(lengths):
            out = []
            for L in lengths:
                x = rng.standard_normal((int(L), 512)).astype(np.float32)
                out.append((x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float16))
            return out
        gest, cont = clips(lt), clips(lw)
        for mode in oracle.POOL_MODES:
            ref = oracle.simpool_allpairs(gest, cont, mode)
            got = scoring.score_allpairs(gest, cont, mode)
            assert np.abs(got - ref).max() < TOL, (trial, mode, lt.tolist(), lw.tolist())
        n = min(ng, nc)
        widx = [int(rng.integers(0, len(c))) for c in cont[:n]]
        r = scoring.spot_batch(gest[:n], cont[:n], widx, want_full=True)
        for i in range(n):
            assert np.abs(r["full"][i] - oracle.get_attn_matrix(gest[i], cont[i])).max() < PROB_TOL, (trial, i)
