"""Per-role cycle accounting of K1 (JEGAL_K1_TRACE=1) on config 5 and on the ragged config 2."""
import os, sys
os.environ["JEGAL_K1_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jegal_b200 import ops, synth
dev = torch.device("cuda:0")
cs = synth.cfg2_retrieval(1000, device=dev)
gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
g16, _ = ops.prep(cs.gest, gl)
c16, _ = ops.prep(cs.cont, cl)
for mode in ("max_t_mean_w", "max_w_mean_t", "mean_mean", "max_max"):
    for _ in range(2):
        print(mode, file=sys.stderr, flush=True)
        ops.simpool_allpairs(g16, gl, c16, cl, mode)
Q, G, T, W = 1000, 65536, 64, 16
q, g, _ = synth.cfg5_gallery(Q, G, T, W, seed=1239, device=dev)
ql, gl = ops.Layout.from_lengths([T] * Q), ops.Layout.from_lengths([W] * G)
q16, _ = ops.prep(q, ql)
g16, _ = ops.prep(g, gl)
for _ in range(2):
    print("cfg5 max_t_mean_w", file=sys.stderr, flush=True)
    ops.simpool_allpairs(q16, ql, g16, gl, "max_t_mean_w")
