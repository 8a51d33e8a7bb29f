#!/usr/bin/env python
"""Host<->device copy bandwidth of the box, to read the e2e (host-buffer) leg of bench.py against: pinned H2D, D2H and both
directions at once, on one rank alone and on all ranks simultaneously (contiguous 2 GB copies on two streams).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/pcie_probe.py
prints one JSON object on rank 0."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gb = 2
    n = gb * (1 << 30) // 8
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float64, device=dev)
    d_out = torch.ones(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(kind, everyone):
        active = everyone or rank == 0
        barrier()
        t0 = time.perf_counter()
        nbytes = 0
        if active:
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
                nbytes += n * 8
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
                nbytes += n * 8
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt, float(nbytes)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t[0] = mx[0]
        return float(t[1]) / float(t[0]) * 1e-9     # aggregate GB/s over the ranks that copied

    out = {"n_gpus": world, "gb_per_direction_per_rank": gb, "unit": "GB/s aggregate over the active ranks (wall clock, max over ranks)"}
    for kind in ("h2d", "d2h", "both"):
        for everyone in (False, True):
            run(kind, everyone)
            out[f"{kind}_{'all_ranks' if everyone else 'rank0_alone'}"] = round(max(run(kind, everyone) for _ in range(2)), 1)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
