"""Small shapes through every kernel, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
the smoke() set (all pooling modes fused + two-pass, dense epilogue, top-k, spotting with the normalisation fused
into the load, ASD, word-level pooling) plus a K1 case whose clips straddle tiles and a > 256-row column clip."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as entry
from jegal_b200 import scoring, synth

entry.smoke()
cs = synth.make_clipset([300, 17, 129, 64, 1, 255], [70, 5, 33, 64, 1, 12], seed=9, with_targets=True)
gest, cont = cs.gesture_list(), cs.content_list()
for mode in ("max_t_mean_w", "max_w_mean_t", "mean_mean", "max_max"):
    for rowmat in ("0", "1"):
        os.environ["JEGAL_ROWMAT"] = rowmat
        try:
            scoring.score_allpairs(gest, cont, mode)
        except scoring.JegalError as e:  # fused max-then-mean refuses the > 256-row column clip: expected
            assert "two-pass" in str(e) or "rows on the column side" in str(e), e
os.environ.pop("JEGAL_ROWMAT", None)
r = scoring.spot_batch(gest, cont, cs.target_word, want_full=True)
scoring.asd_batch(cont, gest, np.arange(6), np.arange(6), 2, mode="max_max")
v, i = scoring.retrieve_topk(gest, cont, k=3)
torch.cuda.synchronize()
print("sanitize target done")
