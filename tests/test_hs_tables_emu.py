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


def house_batch(name, num_envs=1, step_meta=None, **kw):
    cfg = dict(SH.VARIANTS[name](PNS), step_meta=step_meta)
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


@pytest.mark.parametrize("name", ["shipped", "two_vehicles"])
def test_hs_telemetry_rows_match_reference_step_meta(name):
    """PGW_F_TELEMETRY: the numbers of the reference's per-device step_meta records (cost, reward,
    raw action, solar / battery / grid power consumed, device_custom_info) against the golden."""
    g = np.load(os.path.join(GOLD, f"hs_{name}.npz"))
    env = house_batch(name, step_meta=True, _dry_run=True)
    plain = house_batch(name, _dry_run=True)
    assert env._b.sd_rows == plain._b.sd_rows + 4 * 13
    emu = EmulatedEnv(env)
    emu.reset(g["init_soc"].reshape(1, 1))
    house = env.agents[0]
    for t in range(g["actions"].shape[0]):
        obs, rew, _ = emu.step(g["actions"][t].reshape(4, 1))
        np.testing.assert_allclose(obs[:, 0], g["obs"][t], rtol=0, atol=1e-14)
        for k, comp in enumerate(house.envs):
            off, n = house.telemetry_rows(comp)
            want = g["telemetry"][t, k]
            have = emu.sd[off:off + n, 0]
            m = ~np.isnan(want)
            np.testing.assert_allclose(have[m], want[m], rtol=1e-13, atol=1e-13,
                                       err_msg=f"t={t} {comp.name}")


def test_hs_device_arithmetic_replays_random_houses_over_two_episodes():
    """tests/golden/hs_random_configs.npz: the reference's house on six random parameter sets,
    two consecutive episodes each (what survives a reset: storage cost, meta state; the initial
    SOC is the reference's own unclipped draw, fed through the drawn-SOC path)."""
    import json
    g = np.load(os.path.join(GOLD, "hs_random_configs.npz"))
    for i, m in enumerate(json.loads(str(g["meta"]))):
        cfg = dict(SH.parametrised(PNS, m["hp"]), step_meta=None)
        env = pgw.MultiAgentEnv(
            common_config={"start_time": cfg["start_time"], "end_time": "01-01-2031 00:00:00",
                           "control_timedelta": cfg["control_timedelta"]},
            pf_config=None, num_envs=1, _dry_run=True,
            agents=[{"name": "house", "bus": None, "cls": PNS.HSMultiComponentEnv,
                     "config": house_agent_config(cfg)}])
        emu = EmulatedEnv(env)
        for ep in range(2):
            key = f"{i}_{ep}"
            np.random.seed(m["seed"] + ep)
            soc = env._reset_draws(None)                       # the product's own draw ...
            np.testing.assert_array_equal(soc[:, 0], g["soc_" + key])   # ... is the reference's
            obs0 = emu.reset(soc, drawn=True)
            np.testing.assert_allclose(obs0[:, 0], g["obs0_" + key], rtol=0, atol=1e-14, err_msg=key)
            A = g["act_" + key]
            for t in range(A.shape[0]):
                obs, rew, done = emu.step(A[t].reshape(4, 1))
                np.testing.assert_allclose(obs[:, 0], g["obs_" + key][t], rtol=0, atol=1e-13,
                                           err_msg=f"{key} obs t={t}")
                np.testing.assert_allclose(rew[0, 0], g["rew_" + key][t], rtol=1e-13, atol=1e-13,
                                           err_msg=f"{key} rew t={t}")
                np.testing.assert_allclose(emu.agent_p[0, 0], g["p_" + key][t], rtol=1e-14, atol=1e-14)
            assert done
