#!/usr/bin/env python
"""Frame x word attention heatmap of one clip on the B200 kernels — drop-in for the scoring half
of the reference's utils/plot_heatmap.py (same --path / --fname).  The matrix is computed by K3 and
saved as <fname>.npy; the picture is drawn by jegal_b200.render (jet colormap + the reference's
thresholded overlay, standard library PNG writer: neither matplotlib nor cv2 is needed)."""
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
    from jegal_b200 import render
    print("wrote", render.render_heatmap(attn_mtx, words, fname=args.fname))
    return attn_mtx, words


if __name__ == "__main__":
    main()
