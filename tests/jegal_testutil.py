"""Helpers shared by the tests (module name kept unique: other `tests` packages exist on sys.path)."""
import numpy as np


def split(rows, cu):
    return [rows[cu[i]:cu[i + 1]] for i in range(len(cu) - 1)]


def gap_aware_equal(got_idx, ref_idx, ref_scores_of, tol=2e-3):
    """Integer decisions must match wherever the oracle's score gap exceeds `tol`:
    a mismatch got != ref is an error only if |score(ref) - score(got)| > tol."""
    got_idx = np.asarray(got_idx)
    ref_idx = np.asarray(ref_idx)
    bad = []
    for n in np.flatnonzero(got_idx != ref_idx):
        if abs(ref_scores_of(n, ref_idx[n]) - ref_scores_of(n, got_idx[n])) > tol:
            bad.append(int(n))
    return bad


def wordlevel_case(g):
    """Rebuild the python-side inputs of tests/golden/wordlevel.npz (oracle/make_golden.py::wordlevel_inputs)."""
    import torch

    text, bounds, pos = [], [], 0
    for b, n in enumerate(g["n_words"]):
        text.append([f"w{b}_{w}" for w in range(int(n))])
        bounds.append([[f"w{b}_{w}", int(g["bounds"][pos + w][0]), int(g["bounds"][pos + w][1])] for w in range(int(n))])
        pos += int(n)
    return (torch.from_numpy(g["text_emb"]), torch.from_numpy(g["audio_emb"]), torch.from_numpy(g["input_ids"]),
            torch.from_numpy(g["offsets"]), text, bounds)
