"""Per-role cycle accounting of the grouped kernel (JEGAL_GROUPED_TRACE=1) at the config-3 / config-4 sizes."""
import os
import sys

os.environ["JEGAL_GROUPED_TRACE"] = "1"
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
    print("unfused", file=sys.stderr)
    ops.spot(g16, gl, c16, cl, wi)
    print("fused", file=sys.stderr)
    ops.spot(cs.gest, gl, cs.cont, cl, wi, normalize=True)
torch.cuda.synchronize()
