"""GPU tier: the reference-facing API surface beyond MultiAgentEnv.reset/step -- the list
interface of the MADDPG example, components stepped on their own (the reference's
tests/agents/*.py pattern, min / max / random policies), checkpoint-resume, and seeded
random component configurations against the oracle."""
import os

import numpy as np
import pytest

import powergridworld_b200 as pgw
from tests import scenarios as S
from tests.flatten import flat_obs, unflatten_action
from tests.oracle_ns import ORACLE_NS as ONS, storage_socs_to_dict
from tests.product_ns import PRODUCT_NS as PNS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _torch():
    import torch
    return torch


def test_list_interface_replays_reference_trace():
    """examples/marl/openai/train.py:165-188: MultiAgentListInterfaceEnv over the coordinated
    buildings env, driven with per-agent action lists."""
    g = np.load(os.path.join(GOLD, "c0_buildings.npz"))
    env = pgw.MultiAgentListInterfaceEnv(pgw.CoordinatedMultiBuildingControlEnv,
                                         S.buildings_scenario(PNS, PNS.OpenDSSSolver, 1.2))
    assert env.n == 3 and [s.shape for s in env.observation_space] == [(17,)] * 3
    obs = env.ma_env.reset(init_storage=g["init_soc"])
    obs = env.convert_to_list_obs(obs)
    np.testing.assert_allclose(np.concatenate(obs), g["obs0"], rtol=0, atol=1e-7)
    for t in range(60):
        acts = [g["actions"][t][8 * i:8 * i + 8] for i in range(3)]
        obs, rew, done, _ = env.step(acts)
        np.testing.assert_allclose(np.concatenate(obs), g["obs"][t], rtol=0, atol=1e-7)
        np.testing.assert_allclose(rew, g["rew"][t], rtol=1e-5, atol=2e-5)
        assert done == [bool(g["done"][t])] * 3


def test_list_interface_batched_views_are_zero_copy():
    torch = _torch()
    env = pgw.MultiAgentListInterfaceEnv(pgw.CoordinatedMultiBuildingControlEnv,
                                         S.buildings_scenario(PNS, PNS.OpenDSSSolver, 1.2),
                                         num_envs=64)
    obs = env.reset_batch()
    assert [tuple(o.shape) for o in obs] == [(17, 64)] * 3
    assert obs[1].data_ptr() == env.ma_env.obs[17:].data_ptr()
    act = torch.zeros((env.ma_env.act_dim, 64), dtype=torch.float64, device="cuda")
    views = env.action_views(act)
    views[2].fill_(0.5)
    assert float(act[16:24].min()) == 0.5 and float(act[:16].max()) == 0.0
    obs, rew, done, all_done = env.step_batch(act)
    assert len(rew) == 3 and tuple(rew[0].shape) == (64,)


def test_standalone_ev_station_reproduces_notebook_totals():
    """examples/envs/ev-charging.ipynb cells 5-7, through EVChargingEnv.reset()/step()."""
    g = np.load(os.path.join(GOLD, "ev_totals.npz"))
    cfg = {"num_vehicles": 100, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
           "peak_threshold": 250., "vehicle_multiplier": 5., "rescale_spaces": False}
    pol = {"high": lambda e: e.action_space.high, "low": lambda e: e.action_space.low,
           "const0.8": lambda e: np.array([.8])}
    for key, want in zip(g["keys"], g["notebook"]):
        env = PNS.EVChargingEnv(**cfg)
        obs, meta = env.reset()
        assert obs.shape == (6,) and meta == {}
        done, tot, n = False, 0.0, 0
        while not done:
            obs, r, done, _ = env.step(pol[str(key)](env))
            tot += r
            n += 1
        assert n == 286
        np.testing.assert_allclose(tot * env.reward_scale, want, rtol=1e-12)


@pytest.mark.parametrize("kind", ["low", "high", "random"])
def test_standalone_components_run_like_reference_agent_tests(kind):
    """tests/agents/test_{energy_storage,pv,building,ev_charging}.py of the reference: full
    episodes under min / max / random policies -- here additionally checked step by step
    against the oracle classes."""
    rng = np.random.default_rng(0)

    def act(space):
        if kind == "low":
            return space.low.copy()
        if kind == "high":
            return space.high.copy()
        return rng.uniform(space.low, space.high)

    cases = [
        (PNS.EnergyStorageEnv, ONS.EnergyStorageEnv, dict(name="storage"), 287),
        (PNS.PVEnv, ONS.PVEnv, dict(name="pv", profile_csv="pv_profile.csv", scaling_factor=10.), 286),
        (PNS.FiveZoneROMThermalEnergyEnv, ONS.FiveZoneROMThermalEnergyEnv,
         dict(name="b", start_time="08-12-2020 00:00:00", end_time="08-13-2020 00:00:00"), 285),
        (PNS.EVChargingEnv, ONS.EVChargingEnv,
         dict(name="ev", num_vehicles=100, max_charge_rate_kw=7., peak_threshold=250.,
              vehicle_multiplier=5., rescale_spaces=False), 286),
    ]
    for pcls, ocls, cfg, length in cases:
        p, o = pcls(**cfg), ocls(**cfg)
        kw = {"init_storage": 31.5} if pcls is PNS.EnergyStorageEnv else {}
        p.reset(**kw)
        o.reset(**kw)
        done, n = False, 0
        while not done:
            a = act(p.action_space)
            po, pr, done, _ = p.step(a)
            oo, orr, od, _ = o.step(a)
            np.testing.assert_allclose(po, oo, rtol=1e-9, atol=1e-9, err_msg=f"{pcls.__name__} t={n}")
            np.testing.assert_allclose(pr, orr, rtol=1e-9, atol=1e-12)
            assert done == od
            n += 1
        assert n == length, (pcls.__name__, n)


def test_checkpoint_resume_is_bit_exact():
    torch = _torch()
    E = 96
    mk = lambda: PNS.MultiAgentEnv(**S.heterogeneous_scenario(PNS, PNS.OpenDSSSolver, 0.65), num_envs=E)
    a, b = mk(), mk()
    rng = np.random.default_rng(8)
    soc = rng.uniform(10, 200, size=(a.num_storage, E))
    acts = [torch.as_tensor(rng.uniform(-1, 1, size=(a.act_dim, E))).cuda() for _ in range(30)]
    a.reset_batch(soc)
    for t in range(12):
        a.step_batch(acts[t])
    b.load_state_dict(a.state_dict())              # b was never reset
    for t in range(12, 30):
        oa, ra, _, _ = a.step_batch(acts[t])
        ob, rb, _, _ = b.step_batch(acts[t])
        assert torch.equal(oa, ob) and torch.equal(ra, rb)
    assert torch.equal(a.get_field(0), b.get_field(0)) and torch.equal(a.get_field(1), b.get_field(1))
    assert torch.equal(a.get_field(8), b.get_field(8))


def test_random_component_configurations_match_oracle():
    """Seeded random scenario parameters (ranges, efficiencies, multipliers, observation sets,
    scaled / raw spaces, grid-aware observations) -- 10 scenarios x 3 envs x 40 steps."""
    torch = _torch()
    rng = np.random.default_rng(2024)
    import pandas as pd
    for case in range(10):
        rs = bool(rng.integers(2))
        lo = float(rng.uniform(1, 10)); hi = lo + float(rng.uniform(20, 300))
        obs_keys = ["zone_temp", "zone_upper_viol", "zone_lower_viol", "comfort_lower",
                    "comfort_upper", "outdoor_temp", "p_consumed", "time_of_day",
                    "bus_voltage", "min_voltage", "max_voltage"]
        bounds = {"zone_temp": (16., 40.), "zone_upper_viol": (-10., 10.), "zone_lower_viol": (-12., 9.),
                  "comfort_lower": (20., 25.), "comfort_upper": (25., 30.), "outdoor_temp": (0., 56.),
                  "p_consumed": (0., 150.), "time_of_day": (0., 1.), "bus_voltage": (0.9, 1.1),
                  "min_voltage": (0.85, 1.1), "max_voltage": (0.9, 1.15)}
        picked = [k for k in obs_keys if rng.random() < 0.6] or ["zone_temp"]
        n_veh = int(rng.integers(3, 70))

        def build(ns):
            comps = [
                {"name": "building", "cls": ns.FiveZoneROMThermalEnergyEnv,
                 "config": {"obs_config": {k: bounds[k] for k in picked}, "rescale_spaces": rs}},
                {"name": "pv", "cls": ns.PVEnv,
                 "config": {"profile_csv": ["pv_profile.csv", "off-peak.csv", "constant.csv"][case % 3],
                            "scaling_factor": 5. + 7 * case, "rescale_spaces": rs,
                            "grid_aware": bool(case % 2)}},
                {"name": "storage", "cls": ns.EnergyStorageEnv,
                 "config": {"storage_range": (lo, hi), "max_power": 5. + 3 * case,
                            "charge_efficiency": 0.8 + 0.02 * case,
                            "discharge_efficiency": 0.99 - 0.02 * case, "rescale_spaces": rs}},
            ]
            return {
                "common_config": {"start_time": "03-0%d-2021 0%d:00:00" % (1 + case % 9, case % 10),
                                  "end_time": "03-1%d-2021 00:00:00" % (case % 9),
                                  "control_timedelta": pd.Timedelta(300, "s")},
                "pf_config": {"cls": ns.OpenDSSSolver,
                              "config": dict(S.IEEE13, system_load_rescale_factor=0.3 + 0.1 * case)},
                "max_episode_steps": 60,
                "agents": [
                    {"name": "house", "bus": ["675c", "634a", "645", "670b"][case % 4],
                     "cls": ns.MultiComponentEnv, "config": {"components": comps}},
                    {"name": "ev", "bus": ["675a", "634c", "684c"][case % 3], "cls": ns.EVChargingEnv,
                     "config": {"num_vehicles": n_veh, "max_charge_rate_kw": 3. + case,
                                "vehicle_multiplier": float(1 + case), "peak_threshold": 20. * (1 + case),
                                "rescale_spaces": rs}},
                    {"name": "farm", "bus": "675c", "cls": ns.GridAwarePVEnv,
                     "config": {"profile_csv": "constant.csv", "scaling_factor": 100. * (1 + case),
                                "rescale_spaces": rs, "grid_aware": True}},
                ]}

        E, T = 3, 40
        env = PNS.MultiAgentEnv(**build(PNS), num_envs=E)
        assert env.episode_length == 59
        soc = rng.uniform(lo, hi, size=(env.num_storage, E))
        acts = rng.uniform(-1.05, 1.05, size=(T, env.act_dim, E))
        if not rs:
            acts = np.abs(acts)
        obs0 = env.reset_batch(soc).cpu().numpy().copy()
        O, R = [], []
        for t in range(T):
            o, r, _, _ = env.step_batch(torch.as_tensor(acts[t]).cuda())
            O.append(o.cpu().numpy().copy())
            R.append(r.cpu().numpy().copy())
        for e in range(E):
            ref = ONS.MultiAgentEnv(**build(ONS))
            o0 = ref.reset(init_storage=storage_socs_to_dict(ref, soc[:, e]))
            np.testing.assert_allclose(obs0[:, e], flat_obs(ref, o0), rtol=0, atol=1e-7,
                                       err_msg=f"case {case} reset")
            for t in range(T):
                o, r, _, _ = ref.step(unflatten_action(ref, acts[t][:, e]))
                np.testing.assert_allclose(O[t][:, e], flat_obs(ref, o), rtol=0, atol=1e-7,
                                           err_msg=f"case {case} env {e} t={t}")
                np.testing.assert_allclose(R[t][:, e], [r[a.name] for a in ref.agents],
                                           rtol=1e-5, atol=2e-5, err_msg=f"case {case} env {e} t={t}")


@pytest.mark.gpu
def test_pf_kernel_keyword_and_house_readme_example():
    """MultiAgentEnv(pf_kernel=...) selects the solver ("auto" = tc2 where the feeder fits);
    the README's Home-Steward snippet runs as written."""
    import numpy as np
    import powergridworld_b200 as pgw
    from tests import scenarios as S
    from tests.product_ns import PRODUCT_NS as PNS
    envs = {k: PNS.CoordinatedMultiBuildingControlEnv(
        **S.buildings_scenario(PNS, PNS.OpenDSSSolver, 1.2), num_envs=64, pf_kernel=k)
        for k in ("fp64", "auto")}
    soc = np.full((3, 64), 25.0)
    v = {k: (e.reset_batch(soc), e.get_field(3).cpu().numpy())[1] for k, e in envs.items()}
    assert 0 < np.abs(v["auto"] - v["fp64"]).max() < 1e-6        # a different (fp32) solver ran
    with pytest.raises(ValueError):
        PNS.CoordinatedMultiBuildingControlEnv(
            **S.buildings_scenario(PNS, PNS.OpenDSSSolver, 1.2), pf_kernel="fastest", _dry_run=True)
    from powergridworld_b200.scenarios.heterogeneous_hs import make_env_config
    house = pgw.HSMultiComponentEnv(**make_env_config())
    obs = house.reset()
    obs, reward, done, meta = house.step(
        {name: space.sample() for name, space in house.action_space.items()})
    assert set(obs) == {"pv", "storage", "ev-charging", "other-devices"} and not done
    assert len(meta["step_meta"]) == 4 and np.isfinite(reward)


def test_update_tables_abi_and_checkpoint_of_randomised_rosters():
    """pgw_update_tables: sizes are those of pgw_create (a changed parameter length is refused,
    null tables are left alone); a checkpoint of a batch with EVChargingEnv(randomize=True)
    carries the drawn rosters, so a fresh env resumes it bit for bit."""
    import ctypes as C
    from powergridworld_b200 import _native as N
    torch = _torch()
    E = 40
    mk = lambda: PNS.MultiAgentEnv(**S.randomized_ev_scenario(PNS, PNS.OpenDSSSolver), num_envs=E)
    a, b = mk(), mk()
    dp = a._dpar.ctypes.data_as(C.POINTER(C.c_double))
    assert a._lib.pgw_update_tables(a._h, dp, len(a._b.dpar) - 1, None, None, a._stream()) != 0
    with pytest.raises(N.NativeError, match="length changed"):
        N.check(a._lib.pgw_update_tables(a._h, dp, len(a._b.dpar) + 2, None, None, a._stream()))
    assert a._lib.pgw_update_tables(a._h, None, 0, None, None, a._stream()) == 0
    rng = np.random.default_rng(3)
    acts = [torch.as_tensor(rng.uniform(-1, 1, size=(a.act_dim, E))).cuda() for _ in range(140)]
    np.random.seed(77)
    a.reset_batch()
    for t in range(100):                             # well into the day: vehicles are parked
        a.step_batch(acts[t])
    np.random.seed(78)
    b.reset_batch()                                  # another draw, overwritten by the checkpoint
    b.load_state_dict(a.state_dict())
    for t in range(100, 140):
        oa, ra, _, _ = a.step_batch(acts[t])
        ob, rb, _, _ = b.step_batch(acts[t])
        assert torch.equal(oa, ob) and torch.equal(ra, rb)
    assert torch.equal(a.get_field(0), b.get_field(0)) and torch.equal(a.get_field(1), b.get_field(1))


def test_drawn_initial_storage_is_not_clipped_but_explicit_is():
    """energy_storage_env.py:82-89: the SOC the reference draws itself is used as is (here:
    mean 30 +- 5 against a 0..20 kWh range), an explicit init_storage is clipped to the range."""
    cfg = dict(storage_range=(0., 20.), rescale_spaces=False)
    dev, ora = PNS.EnergyStorageEnv(name="s", **cfg), ONS.EnergyStorageEnv(name="s", **cfg)
    np.random.seed(5)
    od, _ = dev.reset()
    np.random.seed(5)
    oo, _ = ora.reset()
    assert oo[0] > 20.0
    np.testing.assert_array_equal(od, oo)
    for t in range(5):                               # discharging brings it back into the range
        a = np.array([0.9])
        rd, ro = dev.step(a), ora.step(a)
        np.testing.assert_allclose(rd[0], ro[0], rtol=0, atol=1e-12)
    od, _ = dev.reset(init_storage=33.0)
    oo, _ = ora.reset(init_storage=33.0)
    np.testing.assert_array_equal(od, oo)
    assert od[0] == 20.0


def test_per_object_protocol_on_random_configurations_vs_reference_traces():
    """tests/golden/component_configs.npz, the cases that need no grid input, through the
    reference's per-object protocol (component.reset() / component.step(action)) on the GPU."""
    from tests.component_cases import build_component
    from tests.test_host_tables_emu import _standalone_cases
    g, cases = _standalone_cases()
    for ci, m in cases[::2]:                         # every second case: a device handle each
        dev = build_component(getattr(PNS, m["cls"]), m["cfg"])
        np.random.seed(m["seed"])
        r0 = dev.reset(**m["reset_kw"])
        o0 = r0[0] if isinstance(r0, tuple) else r0
        if g[f"obs0_{ci}"].size and o0 is not None:
            np.testing.assert_allclose(np.asarray(o0, float), g[f"obs0_{ci}"], rtol=1e-13, atol=1e-13,
                                       err_msg=f"case {ci} ({m['cls']}) reset")
        A = g[f"act_{ci}"]
        for t in range(min(A.shape[0], 120)):
            ob, rew, done, _ = dev.step(A[t])
            np.testing.assert_allclose(np.asarray(ob, float), g[f"obs_{ci}"][t], rtol=1e-12, atol=1e-12,
                                       err_msg=f"case {ci} ({m['cls']}) t={t}")
            np.testing.assert_allclose(rew, g[f"rew_{ci}"][t], rtol=1e-11, atol=1e-14)
            assert bool(done) == bool(g[f"done_{ci}"][t])
