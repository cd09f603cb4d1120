"""Property tests: random component parameters and random action sequences --
including actions outside [-1, 1], which exercise the clip of to_raw -- through the spec compiler
and the host build of the device arithmetic, against the oracle classes stepped with the same
inputs.  Component-only envs (storage / PV / EV station), so no power flow is involved."""
import numpy as np
import pandas as pd
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from tests.emu.harness import EmulatedEnv
from tests.flatten import flat_obs, unflatten_action
from tests.oracle_ns import ORACLE_NS as ONS, storage_socs_to_dict
from tests.product_ns import PRODUCT_NS as PNS

COMMON = {"start_time": "08-12-2021 00:00:00", "end_time": "08-13-2021 00:00:00",
          "control_timedelta": pd.Timedelta(300, "s")}

storage_cfg = st.fixed_dictionaries({
    "max_power": st.floats(1.0, 60.0), "charge_efficiency": st.floats(0.6, 1.0),
    "discharge_efficiency": st.floats(0.6, 1.0),
    "storage_range": st.tuples(st.floats(0.0, 10.0), st.floats(20.0, 90.0)),
    "rescale_spaces": st.booleans()})
pv_cfg = st.fixed_dictionaries({
    "profile_csv": st.sampled_from(["pv_profile.csv", "off-peak.csv", "constant.csv"]),
    "scaling_factor": st.floats(0.5, 80.0), "rescale_spaces": st.booleans()})
ev_cfg = st.fixed_dictionaries({
    "num_vehicles": st.integers(1, 40), "max_charge_rate_kw": st.floats(2.0, 40.0),
    "peak_threshold": st.floats(5.0, 300.0), "vehicle_multiplier": st.sampled_from([1, 2.0, 5.0]),
    "unserved_penalty": st.floats(0.0, 3.0), "peak_penalty": st.floats(0.0, 3.0),
    "rescale_spaces": st.booleans()})


def scenario(ns, s_cfg, p_cfg, e_cfg):
    return {"common_config": COMMON, "pf_config": None, "agents": [
        {"name": "battery", "bus": None, "cls": ns.EnergyStorageEnv, "config": dict(s_cfg)},
        {"name": "pv", "bus": None, "cls": ns.PVEnv, "config": dict(p_cfg)},
        {"name": "station", "bus": None, "cls": ns.EVChargingEnv, "config": dict(e_cfg)},
        {"name": "combo", "bus": None, "cls": ns.MultiComponentEnv, "config": {"components": [
            {"name": "pv", "cls": ns.PVEnv, "config": dict(p_cfg)},
            {"name": "battery", "cls": ns.EnergyStorageEnv, "config": dict(s_cfg)}]}}]}


class _NoPF:
    def __init__(self, **k):
        pass

    def calculate_power_flow(self, *a, **k):
        pass

    def get_bus_voltages(self):
        return {}

    def get_bus_voltage_by_name(self, n):
        return 1.0


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(s_cfg=storage_cfg, p_cfg=pv_cfg, e_cfg=ev_cfg, seed=st.integers(0, 2 ** 31 - 1),
       wild=st.floats(1.0, 1.6))
def test_components_match_oracle_on_random_parameters_and_actions(s_cfg, p_cfg, e_cfg, seed, wild):
    env = PNS.MultiAgentEnv(**scenario(PNS, s_cfg, p_cfg, e_cfg), _dry_run=True)
    oscn = scenario(ONS, s_cfg, p_cfg, e_cfg)
    oscn["pf_config"] = {"cls": _NoPF, "config": {}}
    ref = ONS.MultiAgentEnv(**oscn)
    emu = EmulatedEnv(env)
    rng = np.random.default_rng(seed)
    lo, hi = s_cfg["storage_range"]
    soc = rng.uniform(lo - 2.0, hi + 2.0, size=(env.num_storage, 1))       # reset clips it
    o0 = emu.reset(soc)
    r0 = ref.reset(init_storage=storage_socs_to_dict(ref, soc[:, 0]))
    np.testing.assert_allclose(o0[:, 0], flat_obs(ref, r0), rtol=0, atol=1e-12)
    for t in range(40):
        a = rng.uniform(-wild, wild, size=(env.act_dim, 1))
        o, r, d = emu.step(a)
        ro, rr, rd, _ = ref.step(unflatten_action(ref, a[:, 0]))
        np.testing.assert_allclose(o[:, 0], flat_obs(ref, ro), rtol=1e-12, atol=1e-12, err_msg=f"t={t}")
        np.testing.assert_allclose(r[:, 0], [rr[x.name] for x in ref.agents], rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(emu.agent_p[:, 0], [x.real_power for x in ref.agents],
                                   rtol=1e-12, atol=1e-12)
        assert bool(d) == bool(rd["__all__"])
        if d:
            break


@settings(max_examples=8, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(e_cfg=ev_cfg, s_cfg=storage_cfg, seed=st.integers(0, 2 ** 31 - 1), roster_seed=st.integers(0, 2 ** 31 - 1))
def test_randomized_station_matches_oracle_over_two_draws(e_cfg, s_cfg, seed, roster_seed):
    """EVChargingEnv(randomize=True): the host draws (roster and storage SOC, in the reference's
    order), the rebuilt station tables and the device arithmetic against the oracle, for two
    consecutive resets = two different rosters, through the busy part of the day."""
    e_cfg = dict(e_cfg, randomize=True)
    mk = lambda ns: {"common_config": COMMON, "pf_config": None, "agents": [
        {"name": "station", "bus": None, "cls": ns.EVChargingEnv, "config": dict(e_cfg)},
        {"name": "battery", "bus": None, "cls": ns.EnergyStorageEnv, "config": dict(s_cfg)}]}
    env = PNS.MultiAgentEnv(**mk(PNS), _dry_run=True)
    oscn = mk(ONS)
    oscn["pf_config"] = {"cls": _NoPF, "config": {}}
    ref = ONS.MultiAgentEnv(**oscn)
    emu = EmulatedEnv(env)
    rng = np.random.default_rng(seed)
    for ep in range(2):
        np.random.seed((roster_seed + ep) % 2 ** 32)
        o0 = emu.reset(env._reset_draws(None), drawn=True)
        np.random.seed((roster_seed + ep) % 2 ** 32)
        r0 = ref.reset()
        np.testing.assert_allclose(o0[:, 0], flat_obs(ref, r0), rtol=0, atol=1e-12)
        for t in range(170):
            a = rng.uniform(-1.2, 1.2, size=(env.act_dim, 1))
            o, r, d = emu.step(a)
            ro, rr, rd, _ = ref.step(unflatten_action(ref, a[:, 0]))
            np.testing.assert_allclose(o[:, 0], flat_obs(ref, ro), rtol=1e-12, atol=1e-12,
                                       err_msg=f"ep={ep} t={t}")
            np.testing.assert_allclose(r[:, 0], [rr[x.name] for x in ref.agents], rtol=1e-11, atol=1e-14)


# ------------------------------------------------------------------ Home-Steward house
house_params = st.fixed_dictionaries({
    "pv_scale": st.floats(0.3, 3.0), "max_power": st.floats(2.0, 12.0),
    "lo": st.floats(0.5, 4.0), "hi": st.floats(10.0, 30.0),
    "eta_c": st.floats(0.7, 1.0), "eta_d": st.floats(0.7, 1.0),
    "init_cost": st.floats(0.0, 0.5), "mult": st.sampled_from([1.0, 2.0, 3.0]),
    "rate": st.floats(3.0, 15.0), "grid": st.floats(60.0, 120.0),
    "rescale": st.tuples(st.booleans(), st.booleans(), st.booleans(), st.booleans())})


def house_config(ns, hp):
    from tests import scenarios_hs as SH
    return SH.parametrised(ns, hp)


@settings(max_examples=20, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(hp=house_params, seed=st.integers(0, 2 ** 31 - 1))
def test_house_matches_hs_oracle_on_random_parameters_and_actions(hp, seed):
    from powergridworld_b200.base_hs import house_agent_config
    from tests.oracle_hs_ns import ORACLE_HS_NS as OHS
    from tests.product_hs_ns import PRODUCT_HS_NS as HNS
    cfg = house_config(HNS, hp)
    env = PNS.MultiAgentEnv(
        common_config={"start_time": cfg["start_time"], "end_time": "01-01-2031 00:00:00",
                       "control_timedelta": cfg["control_timedelta"]},
        pf_config=None, _dry_run=True,
        agents=[{"name": "house", "bus": None, "cls": HNS.HSMultiComponentEnv,
                 "config": house_agent_config(cfg)}])
    ref = OHS.HSMultiComponentEnv(**house_config(OHS, hp))
    emu = EmulatedEnv(env)
    rng = np.random.default_rng(seed)
    flat = lambda ob: np.concatenate([np.asarray(ob[c.name], dtype=np.float64).ravel() for c in ref.envs])
    for episode in range(2):                      # costs and meta state survive the reset
        soc = rng.uniform(hp["lo"], hp["hi"])
        o0 = emu.reset(np.array([[soc]]))
        np.testing.assert_allclose(o0[:, 0], flat(ref.reset(init_storage=soc)), rtol=0, atol=1e-12)
        for t in range(45):
            a = np.array([rng.uniform(-1.3, 1.3) if c.rescale_spaces else
                          rng.uniform(c._action_space.low[0], c._action_space.high[0])
                          for c in ref.envs])
            o, r, d = emu.step(a.reshape(4, 1))
            ro, rr, rd, _ = ref.step({c.name: a[k:k + 1] for k, c in enumerate(ref.envs)})
            np.testing.assert_allclose(o[:, 0], flat(ro), rtol=1e-11, atol=1e-12,
                                       err_msg=f"episode {episode} t={t}")
            np.testing.assert_allclose(r[0, 0], rr, rtol=1e-10, atol=1e-13)
            np.testing.assert_allclose(emu.agent_p[0, 0], ref.real_power, rtol=1e-12, atol=1e-13)
