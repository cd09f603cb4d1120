"""GPU probe: zero-copy host step -- wall time per call against the GPU time of its kernel (events around
the call), for C1 at 4096 envs; and the same kernel on device buffers."""
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
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
env = SB.c1_env(num_envs=E, pf_kernel="tc2")
soc = np.full((env.num_storage, E), 30.0)
acts = [torch.rand((env.act_dim, E), dtype=torch.float64).mul_(2).sub_(1).pin_memory() for _ in range(4)]
env.reset_host(soc)
for i in range(20):
    env.step_host(acts[i % 4])
for same_buffer in (False, True):
    env.reset_host(soc)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    t0 = time.perf_counter()
    for i in range(100):
        ev[i][0].record()
        env.step_host(acts[0 if same_buffer else i % 4])
        ev[i][1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e4
    gpu = np.array([a.elapsed_time(b) * 1e3 for a, b in ev])
    print(json.dumps({"mode": "host zero-copy", "same_action_buffer": same_buffer, "wall_us_per_step": wall,
                      "gpu_us_median": float(np.median(gpu)), "gpu_us_min": float(gpu.min())}), flush=True)
# device buffers, same kernel
a = acts[0].cuda()
env.reset_batch(torch.as_tensor(soc).cuda())
for i in range(10):
    env.step_batch(a)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
for i in range(100):
    ev[i][0].record(); env.step_batch(a); ev[i][1].record()
torch.cuda.synchronize()
gpu = np.array([x.elapsed_time(y) * 1e3 for x, y in ev])
print(json.dumps({"mode": "device buffers", "gpu_us_median": float(np.median(gpu))}), flush=True)
# actions on the device, outputs in host memory; and the reverse
import ctypes as C
lib, h = env._lib, env._h
pin = env._pinned()
sp = torch.cuda.current_stream().cuda_stream
dev_obs, dev_rew, dev_done = env.obs, env.rew, env.done
for label, ap, op, rp, dp in (("actions device, outputs host", a.data_ptr(), pin["obs"].data_ptr(), pin["rew"].data_ptr(), pin["done"].data_ptr()),
                              ("actions host, outputs device", acts[0].data_ptr(), dev_obs.data_ptr(), dev_rew.data_ptr(), dev_done.data_ptr())):
    env.reset_batch(torch.as_tensor(soc).cuda())
    for i in range(10):
        lib.pgw_step(h, ap, op, rp, dp, sp)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    for i in range(100):
        ev[i][0].record(); lib.pgw_step(h, ap, op, rp, dp, sp); ev[i][1].record()
    torch.cuda.synchronize()
    gpu = np.array([x.elapsed_time(y) * 1e3 for x, y in ev])
    print(json.dumps({"mode": label, "gpu_us_median": float(np.median(gpu))}), flush=True)
