#!/usr/bin/env python
"""AVS-Spot word spotting on the B200 kernels — drop-in for the reference's
evaluation/evaluate_spotting.py (same --path / --threshold / --frame_threshold, same printed line).
All clips are scored in one launch instead of a Python loop with six tiny torch ops per clip."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from jegal_b200 import pkl_io, scoring  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--path', type=str, required=True, help="Path to the directory to load the extracted JEGAL features")
    parser.add_argument('--threshold', type=float, default=0.5, help="Threshold for spotting")
    parser.add_argument('--frame_threshold', type=int, default=9, help="Frame threshold for spotting")
    parser.add_argument('--index', type=str, default=None,
                        help="Prefix of a packed clip index (jegal_b200.index); built from --path on first use, then the "
                             ".pkl files are not touched again: the rows stream host->device overlapped with the kernel")
    args = parser.parse_args()
    if args.index:
        from jegal_b200 import index, streaming
        from jegal_b200.ops import JegalError

        ds = index.load_or_build(args.path, args.index)
        print("No of files = ", ds.n)
        wbs = [i["word_boundaries"] for i in ds.info]
        word_idx, lo, hi = scoring.spot_targets(ds.info, wbs, args.frame_threshold)
        try:
            r = streaming.spot_streamed(streaming.HostClips.from_index(ds.gesture), streaming.HostClips.from_index(ds.content),
                                        word_idx, windows=(lo, hi), thresh=args.threshold)
        except JegalError:  # e.g. a transcript of more than 64 words: the general route
            r = scoring.spot_batch(ds.gesture.to_packed(), ds.content.to_packed(), word_idx, windows=(lo, hi),
                                   thresh=args.threshold, want_heat=False)
        accuracy = (int(r["correct"].sum()) / ds.n) * 100
        print("Word Spotting Accuracy: {}".format(accuracy))
        return accuracy
    d = pkl_io.load_dir(args.path)
    print("No of files = ", len(d["files"]))
    word_boundaries = [pkl_io.info_get(i, "word_boundaries") for i in d["info"]]
    return scoring.get_spotting_acc(d["info"], d["gesture"], d["content"], word_boundaries,
                                    thresh=args.threshold, frame_thresh=args.frame_threshold)


if __name__ == "__main__":
    main()
