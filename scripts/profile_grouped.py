"""ncu target: K3 at the config-3 size, unfused (pre-normalised operands) and with the normalisation fused into the load."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from jegal_b200 import ops, synth

dev = torch.device("cuda:0")
cs = synth.cfg3_spotting(int(os.environ.get("CFG3_N", 20000)), device=dev)
gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
wi = torch.from_numpy(cs.target_word).to(dev)
g16, _ = ops.prep(cs.gest, gl, out_dtype=torch.float16)
c16, _ = ops.prep(cs.cont, cl, out_dtype=torch.float16)
for rep in range(2):
    ops.spot(g16, gl, c16, cl, wi)
    ops.spot(cs.gest, gl, cs.cont, cl, wi, normalize=True)
torch.cuda.synchronize()
