"""Gallery-sharded retrieval on >= 2 real GPUs (NCCL): N-GPU result == 1-GPU result, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from jegal_b200 import ops, sharded, synth
        Q, G, T, W, k = 64, 4099, 64, 16, 10  # gallery size not divisible by the world size
        q, g, gt = synth.cfg5_gallery(Q, G, T, W, seed=5, device=dev)
        ql = ops.Layout.from_lengths([T] * Q)
        q_in = q.clone() if rank == 0 else torch.zeros_like(q)
        q_in = sharded.broadcast_queries(q_in, Q * T, q.dtype, dev)
        assert torch.equal(q_in, q)
        q16, _ = ops.prep(q_in, ql)
        lo, hi = sharded.shard_range(G, rank, world)
        sl = ops.Layout.from_lengths([W] * (hi - lo))
        s16, _ = ops.prep(g[lo * W:hi * W].contiguous(), sl)
        v, i = sharded.retrieve_topk_sharded(q16, ql, s16, sl, lo, k=k)
        gl = ops.Layout.from_lengths([W] * G)
        g16, _ = ops.prep(g, gl)
        rv, ri = ops.topk(ops.simpool_allpairs(q16, ql, g16, gl, "max_t_mean_w"), k)
        assert torch.equal(i, ri) and torch.equal(v, rv)
        assert np.array_equal(i[:, 0].cpu().numpy(), gt)
        # fused K2 + NVLink peer-memory exchange + merge (no NCCL on the data path), several steps in a row
        ex = ops.TopkExchange(Q, k)
        for step in range(5):
            pv, pi = sharded.retrieve_topk_sharded(q16, ql, s16, sl, lo, k=k, exchange=ex)
            assert torch.equal(pi, ri) and torch.equal(pv, rv), f"p2p exchange differs at step {step}"
        # C2: normalise + cast + all-gather of the query operand in one kernel over peer memory: every rank feeds only
        # its slice of the raw rows and ends with the operand K0 would have produced from all of them, several steps
        # in a row (two buffer parities), then the streamed host-to-host retrieval built on it
        qg = ops.QueryGather(Q * T)
        r0, r1 = qg.slice_rows()
        for step in range(5):
            got = qg.prep_gather(q[r0:r1].contiguous() * (2.0 ** step), r0)  # K0 renormalises (powers of two: bit-identical operand)
            torch.cuda.synchronize()
            assert torch.equal(got, q16), f"prep_gather differs at step {step}"
        from jegal_b200 import streaming
        path = f"/dev/shm/jegal_test_queries_{port}.bin"  # shared memory: every rank maps and page-locks the same bytes
        if rank == 0:
            qh = streaming.shared_host_tensor(path, (Q * T, 512), torch.float16, create=True)
            qh.copy_(q.cpu())
        dist.barrier()
        if rank != 0:
            qh = streaming.shared_host_tensor(path, (Q * T, 512), torch.float16, create=False)
        gal = streaming.StreamedGallery(g[lo * W:hi * W].cpu(), np.full(hi - lo, W), device=dev, idx_base=lo,
                                        schedule=streaming.balanced_schedule(hi - lo, min_clips=16))
        for gather in (qg, True):  # fused peer-memory gather, NCCL all-gather + K0
            sv, si = streaming.retrieve_topk_streamed(qh, ql, gal, k=k, q_gather=gather)
            mv, mi = sharded._cuda_merge(*[torch.stack(x) for x in zip(*_allgather_lists(sv, si, world))])
            assert torch.equal(mi, ri) and torch.equal(mv, rv), f"streamed + q_gather={type(gather).__name__} differs"
        streaming.release_shared_host_tensor(qh)
        dist.barrier()
        if rank == 0:
            os.unlink(path)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def _allgather_lists(v, i, world):
    vs = [torch.empty_like(v) for _ in range(world)]
    is_ = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(vs, v.contiguous())
    dist.all_gather(is_, i.contiguous())
    return list(zip(vs, is_))


def test_sharded_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
