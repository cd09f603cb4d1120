"""Small end-to-end run of every kernel variant (written for compute-sanitizer memcheck / racecheck; the
tool is closed on the current GPU pool, the script still serves as a smoke run of all variants):
IEEE-13 with all three solvers (two-kernel path, the tcgen05 one with its float64 polish) and the
fused step kernel (device buffers and page-locked host buffers, ragged last tile), the
123-bus-class feeder with the tc2 solver (ragged tile), the stand-alone solve, a Home-Steward
house batch, charging stations on shared and on per-env rosters."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from powergridworld_b200 import _native as N
from powergridworld_b200.scenarios import catalog as S
from powergridworld_b200.scenarios import catalog_hs as SH
from powergridworld_b200.scenarios.namespace import PRODUCT_HS_NS as HNS
from powergridworld_b200.scenarios.namespace import PRODUCT_NS as PNS
from powergridworld_b200.base_hs import house_agent_config

rng = np.random.default_rng(0)


def run(env, steps, kernel=None, fused=None, host=False):
    if kernel is not None:
        env.set_option(N.OPT_PF_KERNEL, kernel)
    if fused is not None:
        env.set_option(N.OPT_FUSED, fused)
    if host:
        soc = rng.uniform(10, 40, size=(env.num_storage, env.num_envs))
        env.reset_host(soc)
        for _ in range(steps):
            env.step_host(torch.as_tensor(rng.uniform(-1, 1, size=(env.act_dim, env.num_envs))).pin_memory())
        torch.cuda.synchronize()
        return
    E = env.num_envs
    soc = rng.uniform(10, 40, size=(env.num_storage, E))
    env.reset_batch(soc)
    for _ in range(steps):
        env.step_batch(torch.as_tensor(rng.uniform(-1, 1, size=(env.act_dim, E))).cuda())
    env.stats()
    torch.cuda.synchronize()


for k in (0, 1, 2):
    run(PNS.CoordinatedMultiBuildingControlEnv(
        **S.buildings_scenario(PNS, PNS.OpenDSSSolver, 1.2), num_envs=130), 3, k, fused=0)
for host in (False, True):
    run(PNS.CoordinatedMultiBuildingControlEnv(
        **S.buildings_scenario(PNS, PNS.OpenDSSSolver, 1.2), num_envs=130), 3, 2, fused=2, host=host)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    run(PNS.MultiAgentEnv(**S.der123_scenario(PNS, PNS.OpenDSSSolver), num_envs=130), 2, 2)
s = PNS.OpenDSSSolver(**dict(S.IEEE13, system_load_rescale_factor=0.7))
s.calculate_power_flow(current_time="01-01-2021 05:00:00")
s._host_env.set_option(N.OPT_PF_KERNEL, 2)
s.calculate_power_flow(current_time="01-01-2021 05:00:00")
cfg = SH.two_vehicles(HNS)
run(PNS.MultiAgentEnv(
    common_config={"start_time": cfg["start_time"], "end_time": "01-01-2031 00:00:00",
                   "control_timedelta": cfg["control_timedelta"]},
    pf_config=None, num_envs=70,
    agents=[{"name": "house", "bus": None, "cls": HNS.HSMultiComponentEnv,
             "config": house_agent_config(cfg)}]), 4)
run(PNS.MultiAgentEnv(**S.ev_pv_storage_scenario(PNS), num_envs=70), 4)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    run(PNS.MultiAgentEnv(**S.randomized_ev_scenario(PNS, PNS.OpenDSSSolver), num_envs=70), 4)
print("sanitize_small: done")
