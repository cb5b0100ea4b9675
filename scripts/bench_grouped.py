"""HBM-bound stages at BASELINE configs 2-4: K3 spotting (cfg3), K4 ASD (cfg4), K0 prep, K2 top-k,
and the AVS-Ret-shaped all-pairs pass (cfg2).  Prints one JSON line per stage with achieved GB/s
(algorithmic bytes) against the measured HBM peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from jegal_b200 import ops, scoring, synth


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm = peaks["hbm_gbs"]
    n3 = int(os.environ.get("CFG3_N", 20000))
    # ---- cfg3 spotting
    cs = synth.cfg3_spotting(n3, device=dev)
    gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
    ms_prep = timeit(lambda: (ops.prep(cs.gest, gl), ops.prep(cs.cont, cl)))
    g16, _ = ops.prep(cs.gest, gl)
    c16, _ = ops.prep(cs.cont, cl)
    rows = gl.rows + cl.rows
    print(json.dumps({"stage": "K0 prep (cfg3 operands)", "ms": ms_prep, "GBps": rows * 512 * 4 / ms_prep / 1e6,
                      "frac_hbm": rows * 512 * 4 / ms_prep / 1e6 / hbm, "bytes": rows * 512 * 4}))
    wi = torch.from_numpy(cs.target_word).to(dev)
    lo = torch.zeros(cs.n, dtype=torch.int32, device=dev)
    hi = torch.full((cs.n,), 1000, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.spot(g16, gl, c16, cl, wi, win_lo=lo, win_hi=hi))
    byts = rows * 1024 + gl.rows * 4 + cs.n * 9
    print(json.dumps({"stage": f"K3 spot (cfg3, {cs.n} clips)", "ms": ms, "clips_per_s": cs.n / ms * 1e3,
                      "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm, "bytes": byts}))
    # ---- K5 word-level pooling at the cfg3 shape: every word = mean of its frames' 256-d audio features (fp16)
    feats = torch.randn(gl.rows, 256, device=dev).half()
    sb = torch.from_numpy(np.concatenate([cs.cu_t[i] + np.asarray([b[1] for b in cs.boundaries[i]]) for i in range(cs.n)]).astype(np.int32)).to(dev)
    se = torch.from_numpy(np.concatenate([cs.cu_t[i] + np.asarray([b[2] + 1 for b in cs.boundaries[i]]) for i in range(cs.n)]).astype(np.int32)).to(dev)
    se = torch.minimum(se, torch.tensor(gl.rows, dtype=torch.int32, device=dev))
    outw = torch.empty(sb.numel(), 256, dtype=torch.float16, device=dev)
    ms = timeit(lambda: ops.segment_mean(feats, sb, se, out=outw))
    byts = int((se - sb).sum().item()) * 512 + sb.numel() * 512 + sb.numel() * 8
    print(json.dumps({"stage": f"K5 word-level mean pooling (cfg3 shape: {sb.numel()} words over {gl.rows} frames, D=256 fp16)",
                      "ms": ms, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm, "bytes": byts}))
    del cs, g16, c16, feats
    # ---- cfg4 ASD
    n4 = int(os.environ.get("CFG4_N", 10000))
    ds = synth.cfg4_asd(n4, 4, device=dev)
    cs = ds.clips
    gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
    g16, gs = ops.prep(cs.gest, gl, normalize=False, want_mean_scale=True, mean_eps=1e-8)
    c16, cs_ = ops.prep(cs.cont, cl, normalize=False, want_mean_scale=True, mean_eps=1e-8)
    pg, pc = torch.from_numpy(ds.pair_gest).to(dev), torch.from_numpy(ds.pair_cont).to(dev)
    ms = timeit(lambda: ops.simpool_pairs(g16, gl, c16, cl, pg, pc, "mean_mean", gscale=gs, cscale=cs_, group_size=4))
    lw = np.diff(cs.cu_w)
    byts = (gl.rows + int(lw[ds.pair_cont].sum())) * 1024 + len(pg) * 4 + n4 * 4
    print(json.dumps({"stage": f"K4 pairs (cfg4, {n4} groups x 4)", "ms": ms, "groups_per_s": n4 / ms * 1e3,
                      "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm, "bytes": byts}))
    del ds, cs, g16, c16
    # ---- cfg2 retrieval
    cs = synth.cfg2_retrieval(1000, device=dev)
    gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
    g16, _ = ops.prep(cs.gest, gl)
    c16, _ = ops.prep(cs.cont, cl)
    flops = 2.0 * 512 * gl.rows * cl.rows
    for mode in ("max_t_mean_w", "max_w_mean_t", "mean_mean", "max_max"):
        ms = timeit(lambda: ops.simpool_allpairs(g16, gl, c16, cl, mode))
        print(json.dumps({"stage": f"K1 all-pairs (cfg2 1000x1000 ragged, {mode})", "ms": ms, "TFLOPs": flops / ms / 1e9,
                          "frac_tensor": flops / ms / 1e9 / peaks["bf16_tflops"], "pairs_per_s": 1e6 / ms * 1e3}))
    ms = timeit(lambda: scoring.clip_similarity_matrix(scoring.PackedClips(cs.gest, gl), scoring.PackedClips(cs.cont, cl), device_out=True))
    print(json.dumps({"stage": "reference-parity retrieval scores (cfg2: K0 mean rows + K1 1000x1000)", "ms": ms}))
    s = ops.simpool_allpairs(g16, gl, c16, cl, "mean_mean")
    ms = timeit(lambda: (ops.rank_of_positive(s), ops.rank_of_positive(s.t())))
    print(json.dumps({"stage": "K2 rank_of_positive both directions (cfg2)", "ms": ms}))
    # ---- clip-level cosine matrix (one-row clips: dense GEMM epilogue), 65536 x 1000 clip vectors
    a = torch.nn.functional.normalize(torch.randn(65536, 512, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(1000, 512, device=dev), dim=-1).bfloat16()
    la, lb = ops.Layout.from_lengths([1] * 65536), ops.Layout.from_lengths([1] * 1000)
    ms = timeit(lambda: ops.simpool_allpairs(a, la, b, lb, "mean_mean"))
    byts = (65536 + 1000) * 1024 + 65536 * 1000 * 4
    print(json.dumps({"stage": "K1 dense epilogue: 65536 x 1000 clip-level cosine matrix", "ms": ms,
                      "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm, "TFLOPs": 2.0 * 512 * 65536 * 1000 / ms / 1e9}))
    # ---- K2 top-k at cfg5 size
    x = torch.randn(1000, 65536, device=dev)
    ms = timeit(lambda: ops.topk(x, 10))
    byts = x.numel() * 4 + 1000 * 10 * 8
    print(json.dumps({"stage": "K2 top-10 of 1000 x 65536", "ms": ms, "GBps": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / hbm}))


if __name__ == "__main__":
    main()
