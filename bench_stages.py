"""Workloads of BASELINE.json configs 2-4 for bench.py: device-resident step, host-to-host step, roofline figures,
the reference's CPU path (oracle port) and a parity check, each behind the same small interface.

    cfg2  AVS-Ret-shaped retrieval: 1000 ragged clips, every T x W tile pooled (max over frames, mean over words),
          ranks of the diagonal in both directions                                        -> tensor-bound (K1)
    cfg3  AVS-Spot-shaped spotting: 20 000 ragged clips, heat-map row + argmax frame + decision  -> HBM-bound (K3)
    cfg4  AVS-Asd-shaped speaker selection: 10 000 groups x 4 gesture tracks vs one content track,
          cosine of the temporal means (the reference's score) + argmax                   -> HBM-bound (K0 means)

Algorithmic work per step follows SURVEY.md 8(d): flops = 2 * 512 * sum(T) * sum(W) (all pairs), bytes = stored
operand bytes read once + outputs written once.
"""
from __future__ import annotations

import os
import time
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from jegal_b200 import ops, scoring, streaming, synth


def time_cuda(fn: Callable, n: int = 10, warm: int = 3) -> float:
    """Mean milliseconds per call, CUDA events on the current stream."""
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


def _pin(t: torch.Tensor) -> torch.Tensor:
    return t.cpu().pin_memory()


class Workload:
    name = ""
    desc = ""
    dtype = "f16"
    bound = "hbm"
    kernel = ""
    units = 0            # clip-pair scores per step
    h2d_bytes = 0
    d2h_bytes = 0

    def step(self):
        raise NotImplementedError

    def e2e_step(self):
        raise NotImplementedError

    def roofline_work(self) -> float:
        """Algorithmic bytes (hbm) or flops (tensor) of ONE launch of the dominant kernel."""
        raise NotImplementedError

    def dominant(self):
        """Run only the dominant kernel (its CUDA-event time is the roofline's denominator)."""
        raise NotImplementedError

    def cpu_reference(self, budget_s: float = 15.0) -> dict:
        raise NotImplementedError

    def parity(self) -> dict:
        raise NotImplementedError


# ------------------------------------------------------------------------------------------------ cfg3
class Cfg3Spotting(Workload):
    name = "cfg3"
    kernel = "grouped_kernel<SPOT> (K3, L2 normalisation fused into the operand load)"

    def __init__(self, dev, n: int = 20000, rank: int = 0, cpu_only: bool = False):
        self.dev = dev
        self.cs = synth.cfg3_spotting(n, seed=1237 + 7919 * rank, device=dev)
        cs = self.cs
        self.n = cs.n
        tw = cs.target_word
        self.st = np.array([cs.boundaries[i][int(tw[i])][1] for i in range(cs.n)])
        self.en = np.array([cs.boundaries[i][int(tw[i])][2] for i in range(cs.n)])
        self.win = (np.maximum(self.st - 9, 0).astype(np.int32), (self.en + 9).astype(np.int32))
        self.units = cs.n
        rows = int(cs.cu_t[-1]) + int(cs.cu_w[-1])
        self.desc = (f"AVS-Spot-shaped word spotting: {cs.n} ragged clips (T 25-220, mean {int(cs.cu_t[-1]) / cs.n:.0f}; W 4-12), D=512, "
                     "softmax over words / 0.07, target-word heat-map row + argmax frame + window/threshold decision")
        if cpu_only:
            return
        self.gl, self.cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
        self.wi = torch.from_numpy(tw).to(dev)
        self.lo, self.hi = torch.from_numpy(self.win[0]).to(dev), torch.from_numpy(self.win[1]).to(dev)
        self.bytes = rows * 1024 + self.gl.rows * 4 + cs.n * 9
        self.host = None

    def step(self, ev=None):
        cs = self.cs
        if ev is not None:
            ev[0].record()
        r = ops.spot(cs.gest, self.gl, cs.cont, self.cl, self.wi, win_lo=self.lo, win_hi=self.hi, normalize=True)
        if ev is not None:
            ev[1].record()
        return r

    def dominant(self):
        return self.step()

    def roofline_work(self) -> float:
        return float(self.bytes)

    def e2e_setup(self):
        self.host = (streaming.HostClips(_pin(self.cs.gest), np.diff(self.cs.cu_t), self.dev),
                     streaming.HostClips(_pin(self.cs.cont), np.diff(self.cs.cu_w), self.dev))
        self.h2d_bytes = self.host[0].nbytes + self.host[1].nbytes + 12 * self.n
        self.d2h_bytes = 9 * self.n

    def e2e_step(self):
        return streaming.spot_streamed(self.host[0], self.host[1], self.cs.target_word, windows=self.win)

    def cpu_reference(self, budget_s: float = 15.0) -> dict:
        """evaluate_spotting.py:59-90 as restated by the oracle: a Python loop over clips, six small torch ops each."""
        from oracle import oracle

        gest, cont, tw = self.cs.gesture_list(), self.cs.content_list(), self.cs.target_word
        n = min(self.n, 20000)
        t0 = time.perf_counter()
        done = 0
        for i in range(n):
            a = oracle.get_attn_matrix(gest[i], cont[i])
            oracle.spot_decision(a, int(tw[i]), int(self.st[i]), int(self.en[i]))
            done += 1
            if (i & 255) == 255 and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return dict(value=done / dt, seconds=dt, sample=f"get_attn_matrix + spot decision loop over {done} of {self.n} clips "
                    f"(evaluate_spotting.py:59-90 restated, fp32 torch CPU, {torch.get_num_threads()} threads)")

    def parity(self, sample: int = 400) -> dict:
        from oracle import oracle

        r = self.step()
        pf, ps, ok = r["pred_frame"].cpu().numpy(), r["pred_score"].cpu().numpy(), r["correct"].cpu().numpy().astype(bool)
        heat = r["heat"].cpu().numpy()
        idx = np.linspace(0, self.n - 1, sample).astype(int)
        worst, bad = 0.0, 0
        g, c = self.cs.gest.cpu().numpy(), self.cs.cont.cpu().numpy()
        for i in idx:
            a = oracle.get_attn_matrix(g[self.cs.cu_t[i]:self.cs.cu_t[i + 1]], c[self.cs.cu_w[i]:self.cs.cu_w[i + 1]])
            row = a[int(self.cs.target_word[i])]
            worst = max(worst, float(np.abs(heat[self.cs.cu_t[i]:self.cs.cu_t[i + 1]] - row).max()))
            pred, score, dec = oracle.spot_decision(a, int(self.cs.target_word[i]), int(self.st[i]), int(self.en[i]))
            fragile = abs(score - 0.5) < 2e-4 or (len(row) > 1 and np.sort(row)[-1] - np.sort(row)[-2] < 2e-4)
            if not fragile and (pf[i] != pred or ok[i] != dec):
                bad += 1
        return dict(sample=int(sample), max_abs_prob_err=worst, decisions_differ=bad, accuracy=float(ok.mean()))


# ------------------------------------------------------------------------------------------------ cfg4
class Cfg4Asd(Workload):
    name = "cfg4"
    kernel = "prep_kernel (K0 clip means, read-only pass over the stored rows)"

    def __init__(self, dev, n_groups: int = 10000, tracks: int = 4, rank: int = 0, cpu_only: bool = False):
        self.dev = dev
        ds = synth.cfg4_asd(n_groups, tracks, seed=1238 + 7919 * rank, device=dev)
        cs = ds.clips
        self.tracks, self.n_groups = tracks, n_groups
        # one content track per group: keep the positives' content clips only
        pos = np.arange(n_groups) * tracks
        lw = np.diff(cs.cu_w)
        keep = np.zeros(cs.n, dtype=bool)
        keep[pos] = True
        self.cont = cs.cont[torch.from_numpy(np.repeat(keep, lw)).to(dev)].contiguous()
        self.len_w = lw[pos]
        self.gest, self.len_t = cs.gest, np.diff(cs.cu_t)
        self.pair_gest = np.arange(n_groups * tracks, dtype=np.int32)
        self.pair_cont = np.repeat(np.arange(n_groups, dtype=np.int32), tracks)
        self.units = n_groups * tracks
        self.desc = (f"AVS-Asd-shaped active speaker: {n_groups} groups x {tracks} candidate gesture tracks (T 39-191) vs one content "
                     "track (W 5-17), D=512, cosine of the temporal means (evaluate_asd.py:31-51) + argmax track")
        self.host = None
        if cpu_only:
            return
        self.gl, self.cl = ops.Layout(cs.cu_t), ops.Layout.from_lengths(self.len_w)
        self.pg, self.pc = torch.from_numpy(self.pair_gest).to(dev), torch.from_numpy(self.pair_cont).to(dev)
        self.bytes_means = self.gl.rows * 1024 + self.gl.n_clips * 2048
        self.bytes = (self.gl.rows + self.cl.rows) * 1024 + (self.gl.n_clips + self.cl.n_clips) * 4096 + self.units * 12 + n_groups * 4

    def step(self, ev=None):
        if ev is not None:
            ev[0].record()
        gm, _ = ops.clip_means(self.gest, self.gl, mean_eps=1e-8)
        if ev is not None:
            ev[1].record()
        cm, _ = ops.clip_means(self.cont, self.cl, mean_eps=1e-8)
        s = ops.pair_cosine(gm, cm, self.pg, self.pc, normalize=False)
        return s, ops.group_softmax(s, self.n_groups, self.tracks, want_probs=False)[1]

    def dominant(self):
        return ops.clip_means(self.gest, self.gl, mean_eps=1e-8)

    def roofline_work(self) -> float:
        return float(self.bytes_means)

    def pool_step(self, mode: str = "max_t_mean_w"):
        """The frame x word tile route (K4, normalisation fused into the load) for the same groups."""
        return ops.simpool_pairs(self.gest, self.gl, self.cont, self.cl, self.pg, self.pc, mode, group_size=self.tracks,
                                 normalize=True)

    @property
    def pool_bytes(self) -> int:
        return (self.gl.rows + int(self.len_w[self.pair_cont].sum())) * 1024 + self.units * 4 + self.n_groups * 4

    def e2e_setup(self):
        self.host = (streaming.HostClips(_pin(self.gest), self.len_t, self.dev),
                     streaming.HostClips(_pin(self.cont), self.len_w, self.dev))
        self.h2d_bytes = self.host[0].nbytes + self.host[1].nbytes + 8 * self.units
        self.d2h_bytes = 4 * self.units + 4 * self.n_groups

    def e2e_step(self):
        return streaming.asd_streamed(self.host[0], self.host[1], self.pair_gest, self.pair_cont, self.tracks,
                                      prefixes=(self.tracks,))

    def cpu_reference(self, budget_s: float = 15.0) -> dict:
        """evaluate_asd.py:91-100 as restated by the oracle: per group, mean-pool + CosineSimilarity + softmax + argmax."""
        from oracle import oracle

        g, c = self.gest.cpu().numpy(), self.cont.cpu().numpy()
        cu_t = np.concatenate([[0], np.cumsum(self.len_t)])
        cu_w = np.concatenate([[0], np.cumsum(self.len_w)])
        t0 = time.perf_counter()
        done = 0
        for grp in range(self.n_groups):
            tr = [g[cu_t[grp * self.tracks + k]:cu_t[grp * self.tracks + k + 1]] for k in range(self.tracks)]
            oracle.asd_predict(c[cu_w[grp]:cu_w[grp + 1]], tr, (self.tracks,))
            done += 1
            if (grp & 127) == 127 and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        return dict(value=done * self.tracks / dt, seconds=dt,
                    sample=f"mean-pool + get_similarity_cos + argmax loop over {done} of {self.n_groups} groups "
                           f"(evaluate_asd.py:31-51,91-100 restated, fp32 torch CPU, {torch.get_num_threads()} threads)")

    def parity(self, sample: int = 300) -> dict:
        from oracle import oracle

        s, am = self.step()
        s, am = s.cpu().numpy().reshape(self.n_groups, self.tracks), am.cpu().numpy()
        g, c = self.gest.cpu().numpy(), self.cont.cpu().numpy()
        cu_t = np.concatenate([[0], np.cumsum(self.len_t)])
        cu_w = np.concatenate([[0], np.cumsum(self.len_w)])
        worst, bad = 0.0, 0
        for grp in np.linspace(0, self.n_groups - 1, sample).astype(int):
            q = oracle.asd_mean_emb(c[cu_w[grp]:cu_w[grp + 1]])
            tr = torch.cat([oracle.asd_mean_emb(g[cu_t[grp * self.tracks + k]:cu_t[grp * self.tracks + k + 1]]) for k in range(self.tracks)])
            cos = torch.nn.functional.cosine_similarity(q, tr, dim=1, eps=1e-8).numpy()
            worst = max(worst, float(np.abs(cos - s[grp]).max()))
            srt = np.sort(cos)
            if srt[-1] - srt[-2] > 2e-3 and int(np.argmax(cos)) != int(am[grp]):
                bad += 1
        return dict(sample=int(sample), max_abs_score_err=worst, decisions_differ=bad, accuracy=float((am == 0).mean()))


# ------------------------------------------------------------------------------------------------ cfg2
class Cfg2Retrieval(Workload):
    name = "cfg2"
    bound = "tensor"
    dtype = "bf16"
    kernel = "simpool_kernel (K1)"

    def __init__(self, dev, n: int = 1000, mode: str = "max_t_mean_w", rank: int = 0, cpu_only: bool = False):
        self.dev, self.mode = dev, mode
        self.cs = synth.cfg2_retrieval(n, seed=1236 + 7919 * rank, device=dev)
        cs = self.cs
        self.n = n
        self.units = n * n
        rows_t, rows_w = int(cs.cu_t[-1]), int(cs.cu_w[-1])
        self.flops = 2.0 * 512 * rows_t * rows_w
        self.desc = (f"AVS-Ret-shaped retrieval: {n} ragged clips (T 25-200, W 4-40; {rows_t} frames x {rows_w} words), D=512, "
                     f"every T x W tile pooled ({mode}), full N x N, ranks of the diagonal in both directions (R@1/5/10/25/50, MedR)")
        self.host = None
        if cpu_only:
            return
        self.gl, self.cl = ops.Layout(cs.cu_t), ops.Layout(cs.cu_w)
        self.g16 = torch.empty((self.gl.rows, 512), dtype=torch.bfloat16, device=dev)
        self.c16 = torch.empty((self.cl.rows, 512), dtype=torch.bfloat16, device=dev)
        self.scores = torch.empty((n, n), dtype=torch.float32, device=dev)

    def step(self, ev=None):
        ops.prep(self.cs.gest, self.gl, out=self.g16)
        ops.prep(self.cs.cont, self.cl, out=self.c16)
        if ev is not None:
            ev[0].record()
        s = ops.simpool_allpairs(self.g16, self.gl, self.c16, self.cl, self.mode, out=self.scores)
        if ev is not None:
            ev[1].record()
        return ops.rank_of_positive(s), ops.rank_of_positive(s.t())

    def dominant(self):
        return ops.simpool_allpairs(self.g16, self.gl, self.c16, self.cl, self.mode, out=self.scores)

    def roofline_work(self) -> float:
        return self.flops

    def e2e_setup(self):
        self.host = (streaming.HostClips(_pin(self.cs.gest), np.diff(self.cs.cu_t), self.dev),
                     streaming.HostClips(_pin(self.cs.cont), np.diff(self.cs.cu_w), self.dev))
        self.h2d_bytes = self.host[0].nbytes + self.host[1].nbytes
        self.d2h_bytes = 4 * 4 * self.n

    def e2e_step(self):
        g, c = self.host
        main = torch.cuda.current_stream(self.dev)
        evs = streaming._copy_chunks([g, c], [(0, self.n)], streaming._copy_stream(self.dev), main)
        main.wait_event(evs[0])
        ops.prep(g.rows_dev, self.gl, out=self.g16)
        ops.prep(c.rows_dev, self.cl, out=self.c16)
        s = ops.simpool_allpairs(self.g16, self.gl, self.c16, self.cl, self.mode, out=self.scores)
        a, b = ops.rank_of_positive(s), ops.rank_of_positive(s.t())
        counts = torch.stack([a[0], a[1], b[0], b[1]]).cpu().numpy()
        return (scoring._metrics_from_counts(counts[0], counts[1]), scoring._metrics_from_counts(counts[2], counts[3]))

    def cpu_reference(self, budget_s: float = 15.0) -> dict:
        """Like-for-like: the oracle's fp32 T x W sim-pool (F.normalize + mm + amax/mean) on a clip subsample."""
        from oracle import oracle

        gest, cont = self.cs.gesture_list(), self.cs.content_list()
        sub = 200
        t0 = time.perf_counter()
        oracle.simpool_allpairs(gest[:sub], cont[:sub], self.mode)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        gm, cm = [oracle.mean_pool(x) for x in gest], [oracle.mean_pool(x) for x in cont]
        s = oracle.get_similarity_matrix(cm, gm).numpy()
        oracle.compute_metrics(s), oracle.compute_metrics(s.T)
        dt_lit = time.perf_counter() - t1
        return dict(value=sub * sub / dt, seconds=dt,
                    sample=f"fp32 torch sim-pool ({self.mode}) of {sub} x {sub} of the {self.n} x {self.n} clip pairs (oracle restatement, "
                           f"{torch.get_num_threads()} threads)",
                    reference_literal=dict(seconds=dt_lit, what="mean-pool + get_similarity_matrix + compute_metrics, both directions, all "
                                           f"{self.n} clips (evaluate_retrieval.py:30-31,38-65 restated): the reference's own (collapsed) score"))

    def parity(self, sample: int = 120) -> dict:
        from oracle import oracle

        self.step()
        torch.cuda.synchronize()
        idx = np.linspace(0, self.n - 1, sample).astype(int)
        gest, cont = self.cs.gesture_list(), self.cs.content_list()
        ref = oracle.simpool_allpairs([gest[i] for i in idx], [cont[i] for i in idx], self.mode)
        got = self.scores[torch.from_numpy(idx).to(self.dev)][:, torch.from_numpy(idx).to(self.dev)].cpu().numpy()
        (ng, ne), _ = ops.rank_of_positive(self.scores), None
        return dict(sample=int(sample), max_abs_score_err=float(np.abs(got - ref).max()),
                    recall_at_1=float((ng.cpu().numpy() == 0).mean()))


WORKLOADS = {"cfg2": Cfg2Retrieval, "cfg3": Cfg3Spotting, "cfg4": Cfg4Asd}


# ------------------------------------------------------------------------------------------------ stage table
def stage_table(dev, peaks: dict, reps: int = 10) -> List[dict]:
    """Every kernel of the path once at its BASELINE config size (CUDA events, `reps` launches after 3 warm-ups):
    the figures behind the one headline number, folded into bench.py's default line as `stages`."""
    hbm, tfl = peaks["hbm"], peaks["tflops"]
    out: List[dict] = []

    def add_hbm(stage, ms, byts, **kw):
        out.append(dict(stage=stage, ms=round(ms, 4), GBps=round(byts / ms / 1e6, 1), frac_hbm=round(byts / ms / 1e6 / hbm, 4),
                        bytes=int(byts), **kw))

    def add_tensor(stage, ms, flops, **kw):
        out.append(dict(stage=stage, ms=round(ms, 4), TFLOPs=round(flops / ms / 1e9, 1), frac_tensor=round(flops / ms / 1e9 / tfl, 4),
                        flops=flops, **kw))

    # ---- cfg3: spotting
    w3 = Cfg3Spotting(dev)
    ms = time_cuda(w3.step, reps)
    add_hbm(f"cfg3 K3 spotting, {w3.n} clips, stored fp16 rows -> decisions (normalisation fused into the load)",
            ms, w3.bytes, clips_per_s=round(w3.n / ms * 1e3))
    cs = w3.cs
    rows = w3.gl.rows + w3.cl.rows
    ms = time_cuda(lambda: (ops.prep(cs.gest, w3.gl), ops.prep(cs.cont, w3.cl)), reps)
    add_hbm("K0 prep (normalise + cast, cfg3 operands; the pass the fused K3 no longer needs)", ms, rows * 512 * 4)
    feats = torch.randn(w3.gl.rows, 256, device=dev).half()
    sb = torch.from_numpy(np.concatenate([cs.cu_t[i] + np.asarray([b[1] for b in cs.boundaries[i]]) for i in range(cs.n)]).astype(np.int32)).to(dev)
    se = torch.from_numpy(np.concatenate([cs.cu_t[i] + np.asarray([b[2] + 1 for b in cs.boundaries[i]]) for i in range(cs.n)]).astype(np.int32)).to(dev)
    se = torch.minimum(se, torch.tensor(w3.gl.rows, dtype=torch.int32, device=dev))
    outw = torch.empty(sb.numel(), 256, dtype=torch.float16, device=dev)
    ms = time_cuda(lambda: ops.segment_mean(feats, sb, se, out=outw), reps)
    add_hbm(f"K5 word-level mean pooling (cfg3 shape: {sb.numel()} words over {w3.gl.rows} frames, D=256 fp16)", ms,
            int((se - sb).sum().item()) * 512 + sb.numel() * 512 + sb.numel() * 8)
    del w3, cs, feats, outw
    # ---- cfg4: ASD
    w4 = Cfg4Asd(dev)
    add_hbm(f"cfg4 ASD reference score, {w4.n_groups} groups x {w4.tracks}: K0 clip means x2 + pair cosine + argmax",
            time_cuda(w4.step, reps), w4.bytes)
    add_hbm("cfg4 K0 clip means of the gesture tracks alone (dominant kernel)", time_cuda(w4.dominant, reps), w4.bytes_means)
    for mode in ("max_t_mean_w", "mean_mean"):
        add_hbm(f"cfg4 K4 pairs + group argmax, {mode} over the T x W tile (normalisation fused into the load)",
                time_cuda(lambda: w4.pool_step(mode), reps), w4.pool_bytes)
    del w4
    # ---- cfg2: ragged all-pairs, every pooling mode + the reference-parity route
    w2 = Cfg2Retrieval(dev)
    w2.step()
    for mode in ("max_t_mean_w", "max_w_mean_t", "mean_mean", "max_max"):
        w2.mode = mode
        add_tensor(f"cfg2 K1 all-pairs 1000 x 1000 ragged, {mode}", time_cuda(w2.dominant, reps), w2.flops)
    w2.mode = "max_t_mean_w"
    pg, pc = scoring.PackedClips(w2.cs.gest, w2.gl), scoring.PackedClips(w2.cs.cont, w2.cl)
    out.append(dict(stage="cfg2 reference-parity retrieval scores (K0 clip means + K1 1000 x 1000 clip vectors)",
                    ms=round(time_cuda(lambda: scoring.clip_similarity_matrix(pg, pc, device_out=True), reps), 4)))
    s = w2.scores
    out.append(dict(stage="cfg2 K2 rank_of_positive, both directions", ms=round(time_cuda(lambda: (ops.rank_of_positive(s), ops.rank_of_positive(s.t())), reps), 4)))
    del w2, pg, pc
    # ---- K1 plain-GEMM epilogue and K2 at config-5 size
    a = torch.nn.functional.normalize(torch.randn(65536, 512, device=dev), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(1000, 512, device=dev), dim=-1).bfloat16()
    la, lb = ops.Layout.from_lengths([1] * 65536), ops.Layout.from_lengths([1] * 1000)
    ms = time_cuda(lambda: ops.simpool_allpairs(a, la, b, lb, "mean_mean"), reps)
    add_hbm("K1 dense-store epilogue: 65536 x 1000 clip-level cosine matrix", ms, (65536 + 1000) * 1024 + 65536 * 1000 * 4)
    x = torch.randn(1000, 65536, device=dev)
    ms = time_cuda(lambda: ops.topk(x, 10), reps)
    add_hbm("K2 top-10 of 1000 x 65536", ms, x.numel() * 4 + 1000 * 10 * 8)
    return out
