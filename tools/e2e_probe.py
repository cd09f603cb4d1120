"""GPU probe of the host-buffer (end-to-end) step: us per step for zero-copy / staged copies, chunk counts,
fused kernel on/off.  python tools/e2e_probe.py [envs] [workload]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powergridworld_b200 import _native as N                                   # noqa: E402
from powergridworld_b200.scenarios import bench as SB                          # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = sys.argv[2] if len(sys.argv) > 2 else "c1"
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for fused, zc, chunks in ((1, 1, 1), (1, 1, 2), (1, 1, 4), (1, 1, 8), (1, 0, 1), (0, 1, 1), (0, 0, 1), (0, 0, 4)):
    env = SB.make_env(wl, num_envs=E, **({"pf_kernel": "tc2"} if wl in ("c1", "c3") else {}))
    if env.pf_solver is not None:
        env.set_option(N.OPT_FUSED, fused)
    env.set_option(N.OPT_HOST_ZERO_COPY, zc)
    env.set_option(N.OPT_HOST_CHUNKS, chunks)
    soc = np.full((env.num_storage, E), 30.0)
    acts = [torch.rand((env.act_dim, E), dtype=torch.float64).mul_(2).sub_(1).pin_memory() for _ in range(4)]
    env.reset_host(soc)
    for i in range(20):
        env.step_host(acts[i % 4])
    best = 1e9
    for rep in range(3):
        env.reset_host(soc)
        t0 = time.perf_counter()
        for i in range(100):
            env.step_host(acts[i % 4])
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e4)
    print(json.dumps({"workload": wl, "E": E, "fused": fused, "zero_copy": zc, "chunks": chunks,
                      "us_per_step": best, "env_steps_per_s": E / best * 1e6}), flush=True)
    env.close()
