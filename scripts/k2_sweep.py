"""K2 top-10 of 1000 x 65536 for several JEGAL_TOPK_BLOCKS_PER_SM values: CUDA-event time + check against torch.topk."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jegal_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(1000, 65536, device=dev)
want = torch.topk(x, 10, dim=1)
for cap in sys.argv[1:] or ["8", "16", "24", "32"]:
    os.environ["JEGAL_TOPK_BLOCKS_PER_SM"] = cap
    v, i = ops.topk(x, 10)
    ok = bool(torch.equal(v, want.values) and torch.equal(i.long(), want.indices))
    for _ in range(5):
        ops.topk(x, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        ops.topk(x, 10)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(json.dumps({"blocks_per_sm": int(cap), "ms": round(ms, 4), "GBps": round(x.numel() * 4 / ms / 1e6, 1), "exact": ok}))
