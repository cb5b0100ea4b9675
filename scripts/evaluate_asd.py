#!/usr/bin/env python
"""AVS-Asd active-speaker selection on the B200 kernels — drop-in for the reference's
evaluation/evaluate_asd.py (same --path / --file, same printed lines).  Every clip is loaded once
(the reference re-reads the five negatives of every row) and all groups are scored in one launch."""
import argparse
import ast
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from jegal_b200 import pkl_io, scoring  # noqa: E402


def main_index(args, df):
    import torch

    from jegal_b200 import index, ops

    ds = index.load_or_build(args.path, args.index)
    pos = {n: i for i, n in enumerate(ds.names)}
    groups = []
    for i in range(len(df)):
        row = df.iloc[i]
        q = pkl_io.clip_pkl_name(row.filename)[:-4]
        if q not in pos:  # evaluate_asd.py:66-67
            continue
        cand = [pos[q]] + [pos[n] for n in (pkl_io.clip_pkl_name(x)[:-4] for x in ast.literal_eval(row.neg_files)) if n in pos]
        groups.append(cand)
    dev = torch.device("cuda", torch.cuda.current_device())
    gm = torch.from_numpy(np.ascontiguousarray(ds.gesture.mean)).to(dev)
    cm = torch.from_numpy(np.ascontiguousarray(ds.content.mean)).to(dev)
    results = {}
    for P in (2, 4, 6):
        by_size = {}
        for cand in groups:
            by_size.setdefault(min(P, len(cand)), []).append(cand)
        correct = total = 0
        for size, gl in by_size.items():
            pg = torch.tensor([c for cand in gl for c in cand[:size]], dtype=torch.int32, device=dev)
            pc = torch.tensor([cand[0] for cand in gl for _ in range(size)], dtype=torch.int32, device=dev)
            cos = ops.pair_cosine(gm, cm, pg, pc, normalize=True, eps=1e-8)  # CosineSimilarity, evaluate_asd.py:45-47
            _, am = ops.group_softmax(cos, len(gl), size, want_probs=False)
            correct += int((am == 0).sum())
            total += len(gl)
        results[P] = (correct, total)
    print("Total videos evaluated: {}".format(results[6][1]))
    for P, name in ((2, "2 spk"), (4, "4 spk"), (6, "6 spk")):
        c, t = results[P]
        print("{}: Correct: {} | Total: {} | Acc: {:.3f}".format(name, c, t, c / t if t else float("nan")))
    return results


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--path', type=str, required=True, help="Path to the directory to load the extracted JEGAL features")
    parser.add_argument('--file', type=str, required=True, help="Path to the AVS-ASD csv file")
    parser.add_argument('--index', type=str, default=None,
                        help="Prefix of a packed clip index (jegal_b200.index); built from --path on first use.  The index "
                             "keeps every clip's temporal mean (load_feats, evaluate_asd.py:31-36), so no .pkl is re-read")
    args = parser.parse_args()
    df = pd.read_csv(args.file)
    print("Total files: {}".format(len(df)))
    if args.index:
        return main_index(args, df)
    # groups whose query or any negative is missing are handled like the reference: a missing query
    # skips the row (evaluate_asd.py:66-67), a missing negative is dropped from the list (:81-82)
    groups, needed = [], {}
    for i in range(len(df)):
        row = df.iloc[i]
        q = os.path.join(args.path, pkl_io.clip_pkl_name(row.filename))
        if not os.path.exists(q):
            continue
        cand = [q]
        for neg in ast.literal_eval(row.neg_files):
            pth = os.path.join(args.path, pkl_io.clip_pkl_name(neg))
            if os.path.exists(pth):
                cand.append(pth)
        groups.append(cand)
        for c in cand:
            needed.setdefault(c, len(needed))
    files = sorted(needed, key=needed.get)
    d = pkl_io.load_dir(args.path, files=files)
    results = {}
    for P in (2, 4, 6):
        # the reference slices all_gesture_embs[:P]: a shorter candidate list is scored as it is
        by_size = {}
        for cand in groups:
            by_size.setdefault(min(P, len(cand)), []).append(cand)
        correct = total = 0
        for size, gl in by_size.items():
            pg = np.array([[needed[c] for c in cand[:size]] for cand in gl], dtype=np.int32).reshape(-1)
            pc = np.repeat(np.array([needed[cand[0]] for cand in gl], dtype=np.int32), size)
            r = scoring.asd_batch(d["content"], d["gesture"], pg, pc, tracks=size, prefixes=(size,))
            correct += int((r["pred"][size] == 0).sum())
            total += len(gl)
        results[P] = (correct, total)
    print("Total videos evaluated: {}".format(results[6][1]))
    for P, name in ((2, "2 spk"), (4, "4 spk"), (6, "6 spk")):
        c, t = results[P]
        print("{}: Correct: {} | Total: {} | Acc: {:.3f}".format(name, c, t, c / t if t else float("nan")))
    return results


if __name__ == "__main__":
    main()
