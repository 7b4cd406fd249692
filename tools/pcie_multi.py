#!/usr/bin/env python
"""Raw host<->device copy ceiling of a multi-GPU box, one process per GPU (torchrun), no engine involved:
every rank copies 2 GiB host->device and 2 GiB device->host between pinned memory and its own GPU, first alone in
turn (ranks take turns), then all ranks at once.  Explains the end-to-end scaling curve of bench.py (VERDICT r01
item 5): the e2e leg moves exactly these bytes per rank and step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_multi.py

Rank 0 prints one JSON line.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 31
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def timed(reps=4):
        both()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            both()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    # (a) one rank at a time
    alone = torch.zeros(world, dtype=torch.float64, device=dev)
    for r in range(world):
        barrier()
        if r == rank:
            alone[r] = timed()
        barrier()
    if world > 1:
        dist.all_reduce(alone)
    # (b) all ranks at once
    barrier()
    t_all = torch.tensor([timed()], dtype=torch.float64, device=dev)
    barrier()
    t_list = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    if world > 1:
        dist.all_gather(t_list, t_all)
    else:
        t_list = [t_all]
    numa = None
    try:
        p = torch.cuda.get_device_properties(local)
        with open("/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)) as f:
            numa = int(f.read())
    except Exception:
        pass
    numas = [None] * world
    if world > 1:
        dist.all_gather_object(numas, numa)
    else:
        numas = [numa]
    if rank == 0:
        ts = [float(t.item()) for t in t_list]
        al = [float(x) for x in alone.tolist()]
        print(json.dumps({
            "n_gpus": world, "bytes_each_way_per_rank": n, "cpus_allowed": len(os.sched_getaffinity(0)),
            "alone_ms": [round(x * 1e3, 1) for x in al],
            "alone_GBps_per_direction": [round(n / x / 1e9, 1) for x in al],
            "together_ms": [round(x * 1e3, 1) for x in ts],
            "together_GBps_per_direction_per_rank": [round(n / x / 1e9, 1) for x in ts],
            "together_aggregate_GBps_both_directions": round(2 * n * world / max(ts) / 1e9, 1),
            "gpu_numa_nodes": numas}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
