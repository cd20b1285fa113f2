#!/usr/bin/env python
"""What can the box feed?  Every rank copies a pinned host buffer to its GPU in a loop (cudaMemcpyAsync, 256 MB per
copy), all ranks at once: per-rank and aggregate host-to-device GB/s at this world size.  Names the limiter of the
multi-GPU e2e curve (PCIe link vs. what the host memory / PCIe fabric sustains for N concurrent links).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 tools/probe_h2d.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    world, rank, local = bench.init_dist(int(os.environ.get("WORLD_SIZE", "1")))
    dev = torch.device("cuda", local)
    n = 256 << 20
    out = {}
    for label, nbuf in (("one_buffer_reused", 1), ("eight_buffers_2GiB", 8)):
        host = [torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(rank + 1) for _ in range(nbuf)]
        devb = torch.empty(n, dtype=torch.uint8, device=dev)
        for h in host:
            devb.copy_(h, non_blocking=True)
        bench.barrier(world)
        reps = 24
        t0 = time.perf_counter()
        for i in range(reps):
            devb.copy_(host[i % nbuf], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        mine = n * reps / dt / 1e9
        if world > 1:
            import torch.distributed as td
            rates = [None] * world
            td.all_gather_object(rates, round(mine, 2))
        else:
            rates = [round(mine, 2)]
        out[label] = {"per_rank_gbs": rates, "aggregate_gbs": round(sum(rates), 1)}
        del host
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        cpus = os.cpu_count()
    if world > 1:
        import torch.distributed as td
        numa = [None] * world
        td.all_gather_object(numa, bench.NUMA_INFO)
    else:
        numa = [bench.NUMA_INFO]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "host_cores_visible": cpus, "bytes_per_copy": n, "numa": numa, **out}), flush=True)
    if world > 1:
        import torch.distributed as td
        td.destroy_process_group()


if __name__ == "__main__":
    main()
