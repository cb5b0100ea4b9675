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


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--path', type=str, required=True, help="Path to the directory to load the extracted JEGAL features")
    parser.add_argument('--file', type=str, required=True, help="Path to the AVS-ASD csv file")
    args = parser.parse_args()
    df = pd.read_csv(args.file)
    print("Total files: {}".format(len(df)))
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
