import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jegal_b200 import ops, synth
dev = torch.device("cuda:0")
cs = synth.cfg2_retrieval(1000, device=dev)
gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
g16, _ = ops.prep(cs.gest, gl)
c16, _ = ops.prep(cs.cont, cl)
mode = sys.argv[1] if len(sys.argv) > 1 else "max_w_mean_t"
for _ in range(2):
    ops.simpool_allpairs(g16, gl, c16, cl, mode)
torch.cuda.synchronize()
