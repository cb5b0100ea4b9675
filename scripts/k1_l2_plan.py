"""K1 on config 5 under different L2 plans (JEGAL_CHUNK_MB = size of the row-operand phase, JEGAL_C_POLICY /
JEGAL_R_POLICY = eviction hints of the column / row operand tiles: 0 normal, 1 evict_last, 2 evict_first).
Plain run: CUDA-event time per plan.  Under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:simpool`
(PLAN_REPS=1): one launch per plan, in the order printed, for the DRAM bytes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from jegal_b200 import ops, synth

PLANS = [(24, 2, 1), (24, 1, 2), (24, 1, 0), (16, 1, 2), (48, 2, 1), (96, 2, 1), (24, 0, 0), (12, 1, 2)]


def main():
    dev = torch.device("cuda:0")
    reps = int(os.environ.get("PLAN_REPS", 10))
    Q, G, T, W = 1000, 65536, 64, 16
    q, g, _ = synth.cfg5_sharded(Q, G, T, W, seed=1239, device=dev)
    ql, gl = ops.Layout.from_lengths([T] * Q), ops.Layout.from_lengths([W] * G)
    q16, _ = ops.prep(q, ql)
    g16, _ = ops.prep(g, gl)
    out = torch.empty((Q, G), dtype=torch.float32, device=dev)
    del q, g
    for chunk, cpol, rpol in PLANS:
        os.environ.update(JEGAL_CHUNK_MB=str(chunk), JEGAL_C_POLICY=str(cpol), JEGAL_R_POLICY=str(rpol))
        if reps > 1:
            for _ in range(2):
                ops.simpool_allpairs(q16, ql, g16, gl, "max_t_mean_w", out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.simpool_allpairs(q16, ql, g16, gl, "max_t_mean_w", out=out)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"chunk_mb": chunk, "c_policy": cpol, "r_policy": rpol, "ms": round(e0.elapsed_time(e1) / reps, 3)}), flush=True)


if __name__ == "__main__":
    main()
