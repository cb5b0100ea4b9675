#!/usr/bin/env python
"""Frame x word attention heatmap of one clip on the B200 kernels — drop-in for the scoring half
of the reference's utils/plot_heatmap.py (same --path / --fname).  The matrix is computed by K3;
rendering needs matplotlib + cv2 exactly as in the reference and is skipped (the matrix is saved
as <fname>.npy) when they are not installed."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from jegal_b200 import pkl_io, scoring  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--path", type=str, required=True, help="Path to the JEGAL feature file")
    parser.add_argument("--fname", default="heatmap", help="Name of the heatmap to be saved")
    args = parser.parse_args()
    feats = pkl_io.load_pkl(args.path)
    gesture_emb, content_emb = feats["gesture_emb"], feats["content_emb"]
    word_boundaries = pkl_io.info_get(feats["info"], "word_boundaries")
    print("Gesture emb: ", gesture_emb.shape, "Content emb: ", content_emb.shape)
    attn_mtx, words = scoring.get_attn_matrix(gesture_emb, content_emb, word_boundaries)
    print("Attn mtx: ", attn_mtx.shape)
    print("Words: ", words)
    np.save(args.fname + ".npy", attn_mtx)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots(1, 1, figsize=(16, 20))
        im = ax.imshow(attn_mtx, cmap="jet")
        ax.set_yticks(list(range(len(words))))
        ax.set_yticklabels(words, fontsize=14)
        fig.colorbar(im, ax=ax, fraction=0.02)
        fig.savefig(args.fname + ".png")
    except ImportError:
        print("matplotlib not installed: wrote {}.npy only".format(args.fname))
    return attn_mtx, words


if __name__ == "__main__":
    main()
