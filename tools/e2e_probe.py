import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from powergridworld_b200.scenarios import catalog as S
from powergridworld_b200.scenarios.namespace import PRODUCT_NS as NS
E=4096
env = NS.CoordinatedMultiBuildingControlEnv(**S.buildings_scenario(NS, NS.OpenDSSSolver, 1.2), num_envs=E, pf_kernel="tc2")
soc = np.full((env.num_storage,E),30.0)
acts=[torch.rand((env.act_dim,E),dtype=torch.float64).mul_(2).sub_(1).pin_memory() for _ in range(4)]
env.reset_host(soc)
for i in range(20): env.step_host(acts[i%4])
torch.cuda.synchronize()
for rep in range(3):
    env.reset_host(soc)
    t0=time.perf_counter()
    for i in range(100): env.step_host(acts[i%4])
    torch.cuda.synchronize()
    print("step_host us/step", (time.perf_counter()-t0)*1e4)
# raw C call only
import ctypes as C
hs=env._host_step; po,pr,pd=hs["ptrs"]; st=torch.cuda.current_stream().cuda_stream
env.reset_host(soc)
t0=time.perf_counter()
for i in range(100):
    env._lib.pgw_step_host(env._h, acts[i%4].data_ptr(), po,pr,pd, st)
print("raw C call us/step", (time.perf_counter()-t0)*1e4)
# device-only step
a=acts[0].cuda()
env.reset_host(soc)
torch.cuda.synchronize(); t0=time.perf_counter()
for i in range(40): env.step_batch(a)
torch.cuda.synchronize(); print("device step us", (time.perf_counter()-t0)/40*1e6)
# copies alone
d=torch.empty_like(a); o=torch.empty((env.obs_dim,E),dtype=torch.float64,device='cuda'); ho=torch.empty((env.obs_dim,E),dtype=torch.float64).pin_memory()
for name,fn in [("h2d act", lambda: d.copy_(acts[0],non_blocking=True)), ("d2h obs", lambda: ho.copy_(o,non_blocking=True))]:
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(100): fn(); torch.cuda.synchronize()
    print(name, "us incl sync", (time.perf_counter()-t0)*1e4)
