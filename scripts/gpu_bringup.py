"""GPU bring-up checks, one stage per process (a trapped kernel poisons the CUDA context).

    python scripts/gpu_bringup.py            # run every stage under its own timeout
    python scripts/gpu_bringup.py --stage simpool_small
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _rand_clips(n, lo, hi, seed, dtype="float16"):
    import numpy as np

    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n)
    clips = []
    for L in lens:
        x = rng.standard_normal((int(L), 512)).astype("float32")
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        clips.append(x.astype(dtype))
    return clips


def _pack(clips, device):
    import numpy as np
    import torch

    from jegal_b200 import ops

    lay = ops.Layout.from_lengths([len(c) for c in clips])
    rows = torch.from_numpy(np.concatenate(clips)).to(device)
    return rows, lay


def stage_basic():
    import numpy as np
    import torch

    from jegal_b200 import ops
    from oracle import oracle

    dev = torch.device("cuda:0")
    clips = _rand_clips(37, 1, 50, 1)
    rows, lay = _pack(clips, dev)
    out, sc = ops.prep(rows * 3.0, lay, normalize=True, want_mean_scale=True)
    ref = torch.cat([oracle.normalize_rows(c.astype("float32") * 3.0) for c in clips])
    err = (out.float().cpu() - ref).abs().max().item()
    print("prep max err", err)
    assert err < 5e-3
    ref_sc = oracle.refnorm_scales([(c.astype("float32") * 3.0).astype("float16") for c in clips])
    e2 = np.abs(sc.cpu().numpy() / ref_sc - 1).max()
    print("meannorm rel err", e2)
    assert e2 < 2e-3
    x = torch.randn(50, 3001, device=dev)
    v, i = ops.topk(x, 10, idx_offset=7)
    rv, ri = oracle.topk(x.cpu().numpy(), 10)
    assert np.array_equal(i.cpu().numpy(), ri + 7) and np.array_equal(v.cpu().numpy(), rv)
    xt = torch.randint(0, 5, (20, 100), device=dev).float()  # heavy ties
    v, i = ops.topk(xt, 32)
    rv, ri = oracle.topk(xt.cpu().numpy(), 32)
    assert np.array_equal(i.cpu().numpy(), ri), "tie order"
    sq = torch.randn(64, 64, device=dev)
    g, e = ops.rank_of_positive(sq)
    rg, re_ = oracle.rank_counts(sq.cpu().numpy())
    assert np.array_equal(g.cpu().numpy(), rg) and np.array_equal(e.cpu().numpy(), re_)
    g, e = ops.rank_of_positive(sq.t())
    rg, re_ = oracle.rank_counts(sq.t().cpu().numpy())
    assert np.array_equal(g.cpu().numpy(), rg)
    print("basic OK")


def _check_simpool(gclips, cclips, modes, tol=2e-3, tag=""):
    import numpy as np
    import torch

    from jegal_b200 import ops
    from oracle import oracle

    dev = torch.device("cuda:0")
    g_rows, g_lay = _pack(gclips, dev)
    c_rows, c_lay = _pack(cclips, dev)
    g16, _ = ops.prep(g_rows, g_lay)
    c16, _ = ops.prep(c_rows, c_lay)
    ok = True
    for mode in modes:
        for cm in (False, True):
            s = ops.simpool_allpairs(g16, g_lay, c16, c_lay, mode, content_major=cm)
            torch.cuda.synchronize()
            s = s.t() if cm else s
            ref_k = oracle.simpool_allpairs(gclips, cclips, mode, device="cuda", rows_g=g16.float(), rows_c=c16.float())
            ref = oracle.simpool_allpairs(gclips, cclips, mode, device="cuda")
            e_k = np.abs(s.cpu().numpy() - ref_k).max()
            e_r = np.abs(s.cpu().numpy() - ref).max()
            flag = "OK" if (e_k < 1e-4 and e_r < tol) else "FAIL"
            ok &= flag == "OK"
            print(f"{tag} {mode:14s} cm={int(cm)} err_vs_same_rows={e_k:.2e} err_vs_fp32={e_r:.2e} {flag}")
    return ok


def stage_simpool_small():
    ok = True
    # uniform config-5-like shapes, tiny
    g = _rand_clips(8, 64, 64, 2)
    c = _rand_clips(40, 16, 16, 3)
    ok &= _check_simpool(g, c, ["max_t_mean_w", "mean_mean", "max_w_mean_t", "max_max"], tag="uniform")
    tmax = 120 if os.environ.get("JEGAL_CTA_GROUP") == "1" else 200
    g = _rand_clips(33, 25, tmax, 4)
    c = _rand_clips(45, 4, 40, 5)
    ok &= _check_simpool(g, c, ["max_t_mean_w", "mean_mean", "max_w_mean_t", "max_max"], tag="ragged")
    g = _rand_clips(5, 1, 3, 6)
    c = _rand_clips(3, 1, 2, 7)
    ok &= _check_simpool(g, c, ["max_t_mean_w", "mean_mean", "max_w_mean_t", "max_max"], tag="tiny")
    g = _rand_clips(3, 250, 400, 8)
    c = _rand_clips(4, 260, 300, 9)
    ok &= _check_simpool(g, c, ["mean_mean", "max_max"], tag="long")
    assert ok
    print("simpool_small OK")


def stage_simpool_mid():
    tmax = 120 if os.environ.get("JEGAL_CTA_GROUP") == "1" else 200
    g = _rand_clips(300, 25, tmax, 10)
    c = _rand_clips(300, 4, 40, 11)
    assert _check_simpool(g, c, ["max_t_mean_w", "mean_mean", "max_w_mean_t", "max_max"], tag="mid")
    print("simpool_mid OK")


def stage_time_cfg5():
    import torch

    from jegal_b200 import ops

    dev = torch.device("cuda:0")
    Q, G, T, W = 1000, int(os.environ.get("BRINGUP_G", 65536)), 64, 16
    gen = torch.Generator(device=dev).manual_seed(1)
    g = torch.nn.functional.normalize(torch.randn(Q * T, 512, device=dev, generator=gen), dim=-1).half()
    c = torch.nn.functional.normalize(torch.randn(G * W, 512, device=dev, generator=gen), dim=-1).half()
    gl = ops.Layout.from_lengths([T] * Q)
    cl = ops.Layout.from_lengths([W] * G)
    g16, _ = ops.prep(g, gl)
    c16, _ = ops.prep(c, cl)
    out = torch.empty((Q, G), dtype=torch.float32, device=dev)
    for mode in ("max_t_mean_w", "mean_mean"):
        for _ in range(2):
            ops.simpool_allpairs(g16, gl, c16, cl, mode, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            ops.simpool_allpairs(g16, gl, c16, cl, mode, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 2.0 * 512 * Q * T * G * W
        print(json.dumps({"stage": "time_cfg5", "mode": mode, "cta_group": os.environ.get("JEGAL_CTA_GROUP", "2"),
                          "ms": ms, "tflops": flops / ms / 1e9, "scores_per_s": Q * G / ms * 1e3}))
    # spot-check numerics on a slice
    from oracle import oracle
    import numpy as np

    s = ops.simpool_allpairs(g16, gl, c16, cl, "max_t_mean_w", out=out)
    torch.cuda.synchronize()
    sub_q, sub_g = 16, 512
    S = (g16[: sub_q * T].float() @ c16[-sub_g * W :].float().t()).view(sub_q, T, sub_g, W)
    ref = S.amax(1).mean(-1)
    err = (s[:sub_q, -sub_g:] - ref).abs().max().item()
    print("cfg5 slice err", err)
    assert err < 1e-4
    t0 = time.time()
    v, i = ops.topk(s, 10)
    torch.cuda.synchronize()
    rv, ri = torch.topk(s, 10, dim=1)
    print("topk match", bool((ri.int() == i).all()), "t", time.time() - t0)


STAGES = {
    "basic": (stage_basic, 300),
    "simpool_small": (stage_simpool_small, 300),
    "simpool_mid": (stage_simpool_mid, 300),
    "time_cfg5": (stage_time_cfg5, 400),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    ap.add_argument("--stages", default="basic,simpool_small,simpool_mid,time_cfg5")
    ap.add_argument("--cta-groups", default="1,2")
    a = ap.parse_args()
    if a.stage:
        STAGES[a.stage][0]()
        sys.exit(0)
    results = {}
    for cg in a.cta_groups.split(","):
        for name in a.stages.split(","):
            if name == "basic" and cg != a.cta_groups.split(",")[0]:
                continue
            env = dict(os.environ, JEGAL_CTA_GROUP=cg)
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "--stage", name], env=env, timeout=STAGES[name][1],
                                   capture_output=True, text=True)
                rc, out = r.returncode, r.stdout[-6000:] + r.stderr[-3000:]
            except subprocess.TimeoutExpired as e:
                rc, out = -999, f"TIMEOUT {e}"
            print(f"===== stage {name} cta_group={cg} rc={rc} ({time.time() - t0:.1f}s)\n{out}", flush=True)
            results[f"{name}/cg{cg}"] = rc
    print(json.dumps(results))
