import sys, json, numpy as np, torch
sys.path.insert(0, "/root/repo")
from powergridworld_b200 import _native as N
from powergridworld_b200.scenarios import bench as SB
E = 262144
for polish in (1, 0, 2):
    env = SB.c1_env(num_envs=E, pf_kernel="tc2")
    env.set_option(N.OPT_PF_POLISH, polish)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    soc = torch.full((env.num_storage, E), 30.0, dtype=torch.float64, device="cuda")
    act = torch.rand((env.act_dim, E), dtype=torch.float64, device="cuda") * 2 - 1
    env.reset_batch(soc)
    for i in range(10): env.step_batch(act)
    env.set_kernel_timing(True)
    for i in range(30): env.step_batch(act)
    a, p, n = env.kernel_timing()
    print(json.dumps({"polish": polish, "comp_us": a / n * 1e3, "pf_us": p / n * 1e3}), flush=True)
    env.close()
