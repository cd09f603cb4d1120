"""Home-Steward house through the spec compiler + the host build of the device arithmetic
(tests/emu, same component_math.cuh the CUDA kernel compiles) against the golden traces recorded
from the unmodified reference -- the CPU-only tier's check of the HS device code."""
import os

import numpy as np
import pandas as pd
import pytest

import powergridworld_b200 as pgw
from powergridworld_b200.base_hs import house_agent_config
from tests import scenarios_hs as SH
from tests.emu.harness import EmulatedEnv
from tests.product_hs_ns import PRODUCT_HS_NS as PNS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def house_batch(name, num_envs=1, **kw):
    cfg = SH.VARIANTS[name](PNS)
    return pgw.MultiAgentEnv(
        common_config={"start_time": cfg["start_time"], "end_time": "01-01-2031 00:00:00",
                       "control_timedelta": cfg["control_timedelta"]},
        pf_config=None, num_envs=num_envs,
        agents=[{"name": "house", "bus": None, "cls": PNS.HSMultiComponentEnv,
                 "config": house_agent_config(cfg)}], **kw)


@pytest.mark.parametrize("name", list(SH.VARIANTS))
def test_hs_device_arithmetic_replays_reference_trace(name):
    g = np.load(os.path.join(GOLD, f"hs_{name}.npz"))
    env = house_batch(name, _dry_run=True)
    assert env.episode_length == 288 and env.act_dim == 4 and env.obs_dim == 12
    emu = EmulatedEnv(env)
    obs0 = emu.reset(g["init_soc"].reshape(1, 1))
    np.testing.assert_allclose(obs0[:, 0], g["obs0"], rtol=0, atol=1e-15)
    house = env.agents[0]
    off, _ = house._begin._slot["sd"]
    for t in range(g["actions"].shape[0]):
        obs, rew, done = emu.step(g["actions"][t].reshape(4, 1))
        np.testing.assert_allclose(obs[:, 0], g["obs"][t], rtol=0, atol=1e-14, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew[0, 0], g["rew"][t], rtol=1e-14, atol=1e-14, err_msg=f"rew t={t}")
        assert bool(done) == bool(g["done"][t]), t
        np.testing.assert_allclose(emu.agent_p[0, 0], g["real_power"][t], rtol=1e-15, atol=0)
        # meta rows: pv_power, es_power, es_cost, pv_cost, grid_power vs golden
        # (grid_cost, es_cost, grid_power, pv_power, es_power, pv_cost)
        m = emu.sd[off:off + 5, 0]
        np.testing.assert_allclose([m[3 - 3], m[1], m[2], m[3], m[4]],
                                   [g["meta"][t][3], g["meta"][t][4], g["meta"][t][1],
                                    g["meta"][t][5], g["meta"][t][2]], rtol=1e-15, atol=0,
                                   err_msg=f"meta t={t}")
    # bit-exact share of the whole trace
    assert done


def test_hs_second_episode_keeps_storage_cost_and_meta():
    env = house_batch("shipped", _dry_run=True)
    emu = EmulatedEnv(env)
    g = np.load(os.path.join(GOLD, "hs_shipped.npz"))
    emu.reset(g["init_soc"].reshape(1, 1))
    for t in range(30):
        emu.step(g["actions"][t].reshape(4, 1))
    es = env.agents[0].envs[1]._slot["sd"][0]
    cost, soc = emu.sd[es + 1, 0], emu.sd[es, 0]
    assert cost != 0.25847
    emu.reset(g["init_soc"].reshape(1, 1))
    assert emu.sd[es + 1, 0] == cost and emu.sd[es, 0] == 8.1 != soc
