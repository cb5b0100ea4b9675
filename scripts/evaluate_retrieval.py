#!/usr/bin/env python
"""AVS-Ret retrieval on the B200 kernels — drop-in for the reference's
evaluation/evaluate_retrieval.py (same --path argument, same printed lines).

Differences a user sees: both directions come from ONE pass over the directory (the reference
re-reads every .pkl per direction), the reference's crash at evaluate_retrieval.py:89 (2-value
unpack of load_feats' 4-tuple) does not exist, and `--pool` can score the frame x word tiles
instead of the mean-pooled clips.  The .pkl format is unchanged.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from jegal_b200 import pkl_io, scoring  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--path', type=str, required=True, help="Path to the directory to load the extracted JEGAL features")
    parser.add_argument('--pool', type=str, default="reference",
                        choices=["reference", "mean_mean", "max_t_mean_w", "max_w_mean_t", "max_max"],
                        help="reference = cosine of mean-pooled clips (evaluate_retrieval.py:30-31,38-48)")
    parser.add_argument('--index', type=str, default=None,
                        help="Prefix of a packed clip index (jegal_b200.index); built from --path on first use")
    args = parser.parse_args()
    if args.index:
        from jegal_b200 import index

        ds = index.load_or_build(args.path, args.index)
        print("No of files = ", ds.n)
        if args.pool == "reference":  # the index keeps load_feats' temporal means (evaluate_retrieval.py:30-31)
            g2c = scoring.compute_metrics(scoring.get_similarity_matrix(ds.gesture.mean, ds.content.mean))
            c2g = scoring.compute_metrics(scoring.get_similarity_matrix(ds.content.mean, ds.gesture.mean))
        else:
            c2g, g2c = scoring.retrieval_metrics(ds.gesture.to_packed(), ds.content.to_packed(), mode=args.pool)
    else:
        d = pkl_io.load_dir(args.path)
        print("No of files = ", len(d["files"]))
        c2g, g2c = scoring.retrieval_metrics(d["gesture"], d["content"], mode=args.pool)
    print("Content to Gesture Retrieval scores:")
    scoring.print_computed_metrics(c2g)
    print("-" * 97)
    print("-" * 97)
    print("Gesture to Content Retrieval scores:")
    scoring.print_computed_metrics(g2c)
    return c2g, g2c


if __name__ == "__main__":
    main()
