"""A/B harness for K1 builds (JEGAL_B200_LIB selects the .so): parity of a small ragged case against a
torch fp32 computation on the GPU, then CUDA-event timings of config 2 (all modes) and config 5."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from jegal_b200 import ops, synth

dev = torch.device("cuda:0")


def ref_scores(g, cu_t, c, cu_w, mode):
    s = g.float() @ c.float().t()
    out = torch.empty(len(cu_t) - 1, len(cu_w) - 1, device=s.device)
    for i in range(len(cu_t) - 1):
        si = s[cu_t[i]:cu_t[i + 1]]
        for j in range(len(cu_w) - 1):
            x = si[:, cu_w[j]:cu_w[j + 1]]
            if mode == "mean_mean":
                out[i, j] = x.mean()
            elif mode == "max_t_mean_w":
                out[i, j] = x.amax(0).mean()
            elif mode == "max_w_mean_t":
                out[i, j] = x.amax(1).mean()
            else:
                out[i, j] = x.amax()
    return out


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
    tag = os.path.basename(os.environ.get("JEGAL_B200_LIB", "default"))
    rng = np.random.default_rng(3)
    nt, nw = 70, 90
    T = rng.integers(1, 200, nt).tolist()
    W = rng.integers(1, 41, nw).tolist()
    ct = np.concatenate([[0], np.cumsum(T)]).astype(np.int32)
    cw = np.concatenate([[0], np.cumsum(W)]).astype(np.int32)
    gest = torch.nn.functional.normalize(torch.randn(int(ct[-1]), 512, device=dev), dim=-1).half()
    cont = torch.nn.functional.normalize(torch.randn(int(cw[-1]), 512, device=dev), dim=-1).half()
    gl, cl = ops.Layout(ct), ops.Layout(cw)
    g16, _ = ops.prep(gest, gl)
    c16, _ = ops.prep(cont, cl)
    worst = {}
    for mode in ("max_t_mean_w", "max_w_mean_t", "mean_mean", "max_max"):
        got = ops.simpool_allpairs(g16, gl, c16, cl, mode)
        torch.cuda.synchronize()
        want = ref_scores(g16, ct, c16, cw, mode)
        worst[mode] = float((got - want).abs().max())
    print(json.dumps({"lib": tag, "parity_max_abs_err": worst}), flush=True)
    assert max(worst.values()) < 2e-3, worst
    if os.environ.get("K1_AB_SMALL"):
        return
    cs = synth.cfg2_retrieval(1000, device=dev)
    gl, cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
    g16, _ = ops.prep(cs.gest, gl)
    c16, _ = ops.prep(cs.cont, cl)
    flops = 2.0 * 512 * gl.rows * cl.rows
    res = {}
    for mode in ("max_t_mean_w", "max_w_mean_t", "mean_mean", "max_max"):
        ms = timeit(lambda: ops.simpool_allpairs(g16, gl, c16, cl, mode))
        res["cfg2_" + mode] = [round(ms, 4), round(flops / ms / 1e9, 1)]
    del cs, g16, c16
    # config 5 operands: 1000 x T=64 query frames, 65536 x W=16 gallery words
    q = torch.nn.functional.normalize(torch.randn(64000, 512, device=dev), dim=-1).bfloat16()
    g = torch.nn.functional.normalize(torch.randn(65536 * 16, 512, device=dev), dim=-1).bfloat16()
    ql, gal = ops.Layout.from_lengths([64] * 1000), ops.Layout.from_lengths([16] * 65536)
    out = torch.empty(1000, 65536, device=dev)
    ms = timeit(lambda: ops.simpool_allpairs(q, ql, g, gal, "max_t_mean_w", out=out), n=10, warm=3)
    res["cfg5_max_t_mean_w"] = [round(ms, 3), round(2.0 * 512 * 64000 * 65536 * 16 / ms / 1e9, 1)]
    print(json.dumps({"lib": tag, "ms_tflops": res}), flush=True)


if __name__ == "__main__":
    main()
