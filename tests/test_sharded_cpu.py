"""World-size-2 gloo test of the gallery-sharded retrieval host logic (sharding, index
offsets, all-gather, merge order).  The CUDA kernels are replaced by oracle stand-ins here —
this covers the plumbing that runs around them on N GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jegal_b200 import sharded, synth
from oracle import oracle


class FakeLayout:
    def __init__(self, lengths):
        self.lengths = np.asarray(lengths)
        self.n_clips = len(lengths)
        self.cu_len = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int32)


def _split(rows, lay):
    r = rows.float().numpy()
    return [r[lay.cu_len[i]:lay.cu_len[i + 1]] for i in range(lay.n_clips)]


def oracle_local_topk(q, ql, g, gl, k, mode, idx_offset, queries_are):
    if gl.n_clips == 0:
        return torch.full((ql.n_clips, k), float("-inf")), torch.full((ql.n_clips, k), -1, dtype=torch.int32)
    s = oracle.simpool_allpairs(_split(q, ql), _split(g, gl), mode)
    v, i = oracle.topk(s, k)
    if v.shape[1] < k:  # shard smaller than k
        pad = k - v.shape[1]
        v = np.pad(v, ((0, 0), (0, pad)), constant_values=-np.inf)
        i = np.pad(i, ((0, 0), (0, pad)), constant_values=-1 - idx_offset)
    return torch.from_numpy(v.copy()), torch.from_numpy((i + idx_offset).astype(np.int32))


def oracle_merge(vals, idxs):
    w, nq, k = vals.shape
    v = vals.permute(1, 0, 2).reshape(nq, w * k).numpy()
    i = idxs.permute(1, 0, 2).reshape(nq, w * k).numpy()
    out_v, out_i = np.empty((nq, k), np.float32), np.empty((nq, k), np.int32)
    for q in range(nq):
        valid = i[q] >= 0
        order = sorted(np.flatnonzero(valid), key=lambda j: (-v[q, j], i[q, j]))[:k]
        out_v[q, :len(order)], out_i[q, :len(order)] = v[q, order], i[q, order]
        out_v[q, len(order):], out_i[q, len(order):] = -np.inf, -1
    return torch.from_numpy(out_v), torch.from_numpy(out_i)


def _worker(rank, world, port, n_gallery, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Q, T, W, k = 6, 8, 4, 5
        q, g, gt = synth.cfg5_gallery(Q, n_gallery, T, W, seed=77)
        ql, gl = FakeLayout([T] * Q), FakeLayout([W] * n_gallery)
        lo, hi = sharded.shard_range(n_gallery, rank, world)
        r0, r1 = sharded.shard_rows(gl.cu_len, lo, hi)
        q_local = q.clone() if rank == 0 else torch.zeros_like(q)
        q_local = sharded.broadcast_queries(q_local, Q * T, q.dtype, "cpu")
        assert torch.equal(q_local, q)
        v, i = sharded.retrieve_topk_sharded(q_local, ql, g[r0:r1], FakeLayout([W] * (hi - lo)), lo, k=k,
                                             local_topk_fn=oracle_local_topk, merge_fn=oracle_merge)
        rv, ri = oracle.topk(oracle.simpool_allpairs(_split(q, ql), _split(g, gl), "max_t_mean_w"), k)
        kk = min(k, n_gallery)
        assert np.array_equal(i.numpy()[:, :kk], ri[:, :kk]), (rank, i, ri)
        np.testing.assert_allclose(v.numpy()[:, :kk], rv[:, :kk], atol=1e-6)
        if n_gallery >= 16:
            assert (i.numpy()[:, 0] == gt).all()  # planted matches are found across shards
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_gallery", [37, 3])
def test_sharded_retrieval_world2_gloo(n_gallery):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_gallery, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
