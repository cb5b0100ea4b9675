"""Configs 2-4 end to end (host arrays in, host results out): the oracle port of the reference's CPU
path on the box's host cores next to the same call through jegal_b200.scoring, in one run.
(the CPU figures are a baseline for context, not the target.)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from jegal_b200 import scoring, synth
from oracle import oracle


def best(fn, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), r


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count()
    out = []
    # ---- config 2: AVS-Ret-shaped retrieval, 1000 clips, reference score + metrics, both directions
    cs = synth.cfg2_retrieval(1000)
    gest, cont = cs.gesture_list(), cs.content_list()

    def cpu_cfg2():
        gm = [oracle.mean_pool(g) for g in gest]
        cm = [oracle.mean_pool(c) for c in cont]
        s = oracle.get_similarity_matrix(cm, gm).numpy()
        return oracle.compute_metrics(s), oracle.compute_metrics(s.T)
    t_cpu, m_cpu = best(cpu_cfg2)
    scoring.retrieval_metrics(gest, cont)  # warm-up (layouts, kernels)
    t_gpu, m_gpu = best(lambda: scoring.retrieval_metrics(gest, cont))
    out.append({"config": "cfg2 AVS-Ret 1000 clips: mean-pool + N x N cosine + R@k/MedR both directions",
                "cpu_s": t_cpu, "gpu_s": t_gpu, "speedup": t_cpu / t_gpu, "cores": cores,
                "same_metrics": m_cpu[0] == m_gpu[0] and m_cpu[1] == m_gpu[1]})
    # the irreducible T x W pooling on the CPU (fp32 torch), 200-clip subsample scaled to 1000 x 1000
    sub = 200
    t_cpu, _ = best(lambda: oracle.simpool_allpairs(gest[:sub], cont[:sub], "max_t_mean_w"), n=2)
    scoring.score_allpairs(gest, cont, "max_t_mean_w")
    t_gpu, _ = best(lambda: scoring.score_allpairs(gest, cont, "max_t_mean_w"))
    out.append({"config": "cfg2 max_t_mean_w pooling of every T x W tile (CPU: 200 x 200 subsample scaled x25)",
                "cpu_s": t_cpu * 25, "gpu_s": t_gpu, "speedup": t_cpu * 25 / t_gpu, "cores": cores})
    # ---- config 3: spotting, CPU loop on a 2000-clip subsample scaled to 20000
    cs = synth.cfg3_spotting(20000)
    gest, cont = cs.gesture_list(), cs.content_list()
    tw = cs.target_word
    st = np.array([cs.boundaries[i][int(tw[i])][1] for i in range(cs.n)])
    en = np.array([cs.boundaries[i][int(tw[i])][2] for i in range(cs.n)])

    def cpu_cfg3(n=2000):
        ok = 0
        for i in range(n):
            a = oracle.get_attn_matrix(gest[i], cont[i])
            ok += oracle.spot_decision(a, int(tw[i]), int(st[i]), int(en[i]))[2]
        return ok
    t_cpu, _ = best(cpu_cfg3, n=2)
    win = (np.maximum(st - 9, 0), en + 9)
    scoring.spot_batch(gest, cont, tw, windows=win)
    t_gpu, _ = best(lambda: scoring.spot_batch(gest, cont, tw, windows=win))
    out.append({"config": "cfg3 AVS-Spot 20000 clips: heatmap row + argmax + decision (CPU: 2000-clip loop scaled x10)",
                "cpu_s": t_cpu * 10, "gpu_s": t_gpu, "speedup": t_cpu * 10 / t_gpu, "cores": cores,
                "note": "GPU time includes packing 20000 numpy clips on the host (several threads) and the H2D copy"})
    # ---- config 4: ASD, 10000 groups x 4 tracks
    ds = synth.cfg4_asd(10000, 4)
    cs = ds.clips
    gest, cont = cs.gesture_list(), cs.content_list()

    def cpu_cfg4(n=2000):
        hit = 0
        for g in range(n):
            hit += oracle.asd_predict(cont[g * 4], [gest[g * 4 + k] for k in range(4)], (4,))[0] == 0
        return hit
    t_cpu, _ = best(cpu_cfg4, n=2)
    scoring.asd_batch(cont, gest, ds.pair_gest, ds.pair_cont, 4, prefixes=(4,))
    t_gpu, _ = best(lambda: scoring.asd_batch(cont, gest, ds.pair_gest, ds.pair_cont, 4, prefixes=(4,)))
    out.append({"config": "cfg4 AVS-Asd 10000 groups x 4 tracks: cosine of mean-pooled clips + argmax (CPU: 2000 groups scaled x5)",
                "cpu_s": t_cpu * 5, "gpu_s": t_gpu, "speedup": t_cpu * 5 / t_gpu, "cores": cores,
                "note": "GPU time includes packing 40000 numpy clips on the host (several threads) and the H2D copy"})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
