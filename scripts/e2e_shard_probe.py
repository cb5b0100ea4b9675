"""One rank's share of the N-GPU end-to-end step, on ONE GPU: a gallery shard of G/N clips streamed from pinned host
memory against the 1000 queries (which also start on the host).  Separates the streaming structure (chunk schedule,
launch overheads) from anything the other ranks or the host's memory system add at N = 8."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from jegal_b200 import ops, streaming, synth


def main():
    dev = torch.device("cuda:0")
    Q, T, W, k = 1000, 64, 16, 10
    for n_shard in [int(x) for x in os.environ.get("SHARDS", "8192,16384").split(",")]:
        q, g, gt = synth.cfg5_sharded(Q, 65536, T, W, seed=1239, device=dev, lo=0, hi=n_shard)
        ql, gl = ops.Layout.from_lengths([T] * Q), ops.Layout.from_lengths([W] * n_shard)
        q16 = torch.empty((Q * T, 512), dtype=torch.bfloat16, device=dev)
        g16 = torch.empty((n_shard * W, 512), dtype=torch.bfloat16, device=dev)
        sc = torch.empty((Q, n_shard), dtype=torch.float32, device=dev)

        def resident():
            ops.prep(q, ql, out=q16)
            ops.prep(g, gl, out=g16)
            ops.simpool_allpairs(q16, ql, g16, gl, "max_t_mean_w", out=sc)
            return ops.topk(sc, k)

        for _ in range(3):
            resident()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            v, i = resident()
        torch.cuda.synchronize()
        res_ms = (time.perf_counter() - t0) / 20 * 1e3
        q_host = q.cpu().pin_memory()
        g_host = g.cpu()
        out = {"shard_clips": n_shard, "resident_ms": round(res_ms, 3), "copy_floor_ms_at_55GBps": round((g_host.numel() * 2 + q_host.numel() * 2) / 55e6, 3)}
        for name, kw in (("ramp(default)", dict(chunk_clips=max(2048, n_shard // 8))), ("4 chunks", dict(chunk_clips=n_shard // 4, ramp=False)),
                         ("geometric", dict(chunk_clips=-1))):
            if kw.get("chunk_clips") == -1:
                gal = streaming.StreamedGallery(g_host, np.full(n_shard, W), device=dev, schedule=streaming.balanced_schedule(n_shard))
            else:
                gal = streaming.StreamedGallery(g_host, np.full(n_shard, W), device=dev, **kw)
            for qp in (1, 4):
                def e2e():
                    vv, ii = streaming.retrieve_topk_streamed(q_host, ql, gal, k=k, q_parts=qp)
                    return vv.cpu(), ii.cpu()
                for _ in range(3):
                    hv, hi = e2e()
                assert np.array_equal(hi.numpy(), i.cpu().numpy())
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(20):
                    e2e()
                torch.cuda.synchronize()
                out[f"e2e_ms[{name}, q_parts={qp}, chunks={len(gal.chunks)}]"] = round((time.perf_counter() - t0) / 20 * 1e3, 3)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
