"""Pinned host->device copy bandwidth of one box, per rank alone and with all ranks copying at once, with the host
buffers allocated before and after binding each rank to the CPUs next to its GPU (bench.bind_to_gpu_numa).
Run under torchrun (one rank per GPU); rank 0 prints one JSON line.  This is the floor of bench.py's e2e leg."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench

MB = int(os.environ.get("H2D_MB", 256))


def measure(host, dev_buf, reps=8):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dev_buf.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return host.numel() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def phase(tag, rank, world, dev):
    host = torch.empty(MB << 20, dtype=torch.uint8, pin_memory=True)
    host.fill_(rank)
    dbuf = torch.empty(MB << 20, dtype=torch.uint8, device=dev)
    measure(host, dbuf, 2)
    solo = torch.zeros(world, dtype=torch.float64, device=dev)
    for r in range(world):
        dist.barrier()
        if r == rank:
            solo[r] = measure(host, dbuf)
    dist.all_reduce(solo)
    dist.barrier()
    conc = torch.zeros(world, dtype=torch.float64, device=dev)
    conc[rank] = measure(host, dbuf)
    dist.all_reduce(conc)
    return {f"{tag}_alone_GBps": [round(x, 1) for x in solo.tolist()], f"{tag}_concurrent_GBps": [round(x, 1) for x in conc.tolist()]}


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world, "MB_per_copy": MB, "host_cpus": os.cpu_count()}
    out.update(phase("unbound", rank, world, dev))
    n = bench.bind_to_gpu_numa(local)
    cpus = [None] * world
    dist.all_gather_object(cpus, n)
    out["cpus_next_to_gpu"] = cpus
    out.update(phase("numa_bound", rank, world, dev))
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
