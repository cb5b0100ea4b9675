"""Quick A/B of the grouped kernels with the row normalisation fused into the load against the K0 + kernel
pipeline, at the config-3 / config-4 sizes (CUDA-event timings, algorithmic bytes from the STORED fp16 rows)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from jegal_b200 import ops, synth
from scripts.bench_grouped import timeit


def main():
    dev = torch.device("cuda:0")
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    cs = synth.cfg3_spotting(int(os.environ.get("CFG3_N", 20000)), device=dev)
    gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
    rows = gl.rows + cl.rows
    wi = torch.from_numpy(cs.target_word).to(dev)
    lo = torch.zeros(cs.n, dtype=torch.int32, device=dev)
    hi = torch.full((cs.n,), 1000, dtype=torch.int32, device=dev)
    byts = rows * 1024 + gl.rows * 4 + cs.n * 9

    def unfused():
        g16, _ = ops.prep(cs.gest, gl, out_dtype=torch.float16)
        c16, _ = ops.prep(cs.cont, cl, out_dtype=torch.float16)
        return ops.spot(g16, gl, c16, cl, wi, win_lo=lo, win_hi=hi)

    def fused():
        return ops.spot(cs.gest, gl, cs.cont, cl, wi, win_lo=lo, win_hi=hi, normalize=True)

    a, b = unfused(), fused()
    same = {k: (float((a[k].float() - b[k].float()).abs().max()) if a[k] is not None else None) for k in ("heat", "pred_frame", "pred_score", "correct")}
    for name, fn in (("K0+K3 (unfused, fp16 operands)", unfused), ("K3 fused normalisation (stored fp16 rows)", fused)):
        ms = timeit(fn)
        print(json.dumps({"stage": f"cfg3 {name}", "ms": ms, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm,
                          "bytes": byts, "max_abs_diff_fused_vs_unfused": same}))
    del cs
    ds = synth.cfg4_asd(int(os.environ.get("CFG4_N", 10000)), 4, device=dev)
    cs = ds.clips
    gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
    pg, pc = torch.from_numpy(ds.pair_gest).to(dev), torch.from_numpy(ds.pair_cont).to(dev)
    lw = np.diff(cs.cu_w)
    byts_pairs = (gl.rows + int(lw[ds.pair_cont].sum())) * 1024 + len(pg) * 4 + len(pg) // 4 * 4
    byts_ref = (gl.rows + cl.rows) * 1024 + (gl.n_clips + cl.n_clips) * 2048 + len(pg) * 4

    def asd_ref():
        gm, _ = ops.clip_means(cs.gest, gl, mean_eps=1e-8)
        cm, _ = ops.clip_means(cs.cont, cl, mean_eps=1e-8)
        s = ops.pair_cosine(gm, cm, pg, pc, normalize=False)
        return ops.group_softmax(s, len(pg) // 4, 4, want_probs=False)

    def asd_pool_fused():
        return ops.simpool_pairs(cs.gest, gl, cs.cont, cl, pg, pc, "max_t_mean_w", group_size=4, normalize=True)

    def asd_pool_unfused():
        g16, _ = ops.prep(cs.gest, gl, out_dtype=torch.float16)
        c16, _ = ops.prep(cs.cont, cl, out_dtype=torch.float16)
        return ops.simpool_pairs(g16, gl, c16, cl, pg, pc, "max_t_mean_w", group_size=4)

    d = float((asd_pool_fused()["scores"] - asd_pool_unfused()["scores"]).abs().max())
    for name, fn, byts in (("reference mode: clip means x2 + pair cosine + argmax", asd_ref, byts_ref),
                           ("K4 max_t_mean_w fused normalisation", asd_pool_fused, byts_pairs),
                           ("K0+K4 max_t_mean_w (unfused)", asd_pool_unfused, byts_pairs)):
        ms = timeit(fn)
        print(json.dumps({"stage": f"cfg4 {name}", "ms": ms, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm,
                          "bytes": byts, "max_abs_diff_fused_vs_unfused": d}))


if __name__ == "__main__":
    main()
