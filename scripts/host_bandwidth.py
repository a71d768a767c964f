#!/usr/bin/env python
"""What the host side of the box delivers to N GPUs at once: every rank copies the bench's
per-step volumes (17.9 MB host->device, 26.5 MB device->host, pinned buffers, two streams)
back to back with no kernel in between.  The aggregate is the ceiling of bench.py's end-to-end
figure at N GPUs: solves/s <= aggregate bytes/s / 44.3 MB * 4096.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/host_bandwidth.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        import bench
        bench.pin_rank_to_cores(local, world)          # same placement as the bench
    except Exception:
        pass
    up, down, slots, steps = 17_858_560, 26_476_544, 4, 200
    h_in = [torch.empty(up, dtype=torch.uint8, pin_memory=True) for _ in range(slots)]
    h_out = [torch.empty(down, dtype=torch.uint8, pin_memory=True) for _ in range(slots)]
    d_in = [torch.empty(up, dtype=torch.uint8, device="cuda") for _ in range(slots)]
    d_out = [torch.empty(down, dtype=torch.uint8, device="cuda") for _ in range(slots)]
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def run(n):
        for i in range(n):
            with torch.cuda.stream(s_up):
                d_in[i % slots].copy_(h_in[i % slots], non_blocking=True)
            with torch.cuda.stream(s_down):
                h_out[i % slots].copy_(d_out[i % slots], non_blocking=True)

    run(20)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_up.wait_stream(torch.cuda.current_stream()); s_down.wait_stream(torch.cuda.current_stream())
    run(steps)
    torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_down)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        total = world * steps * (up + down)
        print(json.dumps({"n_gpus": world, "aggregate_gbs": total / sec / 1e9,
                          "per_gpu_gbs": total / sec / 1e9 / world,
                          "e2e_ceiling_solves_per_s": total / sec / (up + down) * 4096}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
