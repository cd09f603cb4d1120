#!/usr/bin/env python
"""Host-link bandwidth per GPU when 1, 2, 4 and all ranks of the box copy at the same time (pinned host
memory, copy engines, 64 MiB per copy): what bounds the end-to-end figure of N ranks that each move their
observations to the host every step.
    torchrun --standalone --nproc-per-node 8 tools/pcie_concurrency.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import _bind_to_gpu_numa  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
_bind_to_gpu_numa(local)
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n = 64 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if rank == 0:
    topo = os.popen("nvidia-smi topo -m 2>/dev/null").read()
    print(topo, flush=True)
for active in [1, 2, 4, world]:
    if active > world:
        continue
    for direction in ("d2h", "h2d", "both"):
        torch.cuda.synchronize()
        dist.barrier()
        gbs = 0.0
        if rank < active:
            s2 = torch.cuda.Stream()
            for it in range(3):
                e0.record()
                for _ in range(10):
                    if direction in ("d2h", "both"):
                        host.copy_(dev, non_blocking=True)
                    if direction == "both":
                        with torch.cuda.stream(s2):
                            dev2 = getattr(sys.modules[__name__], "_dev2", None)
                            if dev2 is None:
                                dev2 = torch.empty(n, dtype=torch.uint8, device="cuda")
                                host2 = torch.empty(n, dtype=torch.uint8).pin_memory()
                                sys.modules[__name__]._dev2, sys.modules[__name__]._host2 = dev2, host2
                            dev2.copy_(sys.modules[__name__]._host2, non_blocking=True)
                    if direction == "h2d":
                        dev.copy_(host, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s2)
                e1.record()
                torch.cuda.synchronize()
            gbs = 10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
        out = [None] * world
        dist.all_gather_object(out, gbs)
        if rank == 0:
            print(json.dumps({"active_ranks": active, "direction": direction,
                              "gb_per_s_per_rank": [round(x, 1) for x in out[:active]],
                              "note": "both = D2H and H2D streams at once, figure is per direction"}), flush=True)
dist.destroy_process_group()
