"""CPU tier: the spec compiler's tables + the device arithmetic (host build of
csrc/component_math.cuh) + a NumPy statement of the Z-bus fixed point, replayed
against the golden traces recorded from the unmodified reference.  The GPU tier
(tests/test_gpu_parity.py) runs the same traces through the CUDA kernels."""
import os

import numpy as np
import pytest

from tests import scenarios as S
from tests.emu.harness import EmulatedEnv
from powergridworld_b200 import _native as N
from tests.product_ns import PRODUCT_NS as NS

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "c0_buildings": lambda **k: NS.CoordinatedMultiBuildingControlEnv(
        **S.buildings_scenario(NS, NS.OpenDSSSolver, 1.2), **k),
    "heterogeneous": lambda **k: NS.MultiAgentEnv(
        **S.heterogeneous_scenario(NS, NS.OpenDSSSolver, 0.65), **k),
    "heterogeneous_max250": lambda **k: NS.MultiAgentEnv(
        **S.heterogeneous_scenario(NS, NS.OpenDSSSolver, 0.6, max_episode_steps=250), **k),
    "test_heterogeneous": lambda **k: NS.MultiAgentEnv(
        **S.test_heterogeneous_scenario(NS, NS.OpenDSSSolver), **k),
}
for _v in S.TIME_BASE_VARIANTS:                     # other control intervals / day windows
    CASES["timebase_" + _v] = (lambda v: lambda **k: NS.MultiAgentEnv(
        **S.time_base_scenario(NS, NS.OpenDSSSolver, v), **k))(_v)
# tolerances of BASELINE.json north_star: voltages 1e-4 p.u., states 1e-6 rel, rewards 1e-5 rel;
# what is asserted here is much tighter because both sides are float64.
# The shared voltage penalty multiplies a ~3e-9 p.u. solver difference (two formulations of the
# same fixed point, cond(Y) ~ 2e8) by 1e4, hence the reward tolerance.
# The two solvers assemble the same network in different orders; next to the 1e7 S switch
# that alone moves voltages by ~1e-9 p.u.  Voltage obs are scaled x10; the voltage-threshold
# rewards multiply a voltage difference by up to 1e4 (hence the absolute reward floor).
OBS_ATOL, REW_RTOL, REW_ATOL, V_ATOL = 1e-7, 1e-5, 2e-5, 1e-7


@pytest.mark.parametrize("name", list(CASES))
def test_emulated_product_replays_reference_trace(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    env = CASES[name](_dry_run=True)
    T = g["actions"].shape[0]
    assert env.episode_length == T
    assert env.act_dim == g["actions"].shape[1] and env.obs_dim == g["obs"].shape[1]
    emu = EmulatedEnv(env)
    obs0 = emu.reset(g["init_soc"].reshape(-1, 1))
    np.testing.assert_allclose(obs0[:, 0], g["obs0"], rtol=0, atol=OBS_ATOL)
    names = [str(n) for n in g["node_names"]]
    perm = [env.pf_solver.feeder.node_index(n) for n in names]
    np.testing.assert_allclose(emu.vmag[perm, 0], g["volt"][0], rtol=0, atol=V_ATOL)
    for t in range(T):
        obs, rew, done = emu.step(g["actions"][t].reshape(-1, 1))
        np.testing.assert_allclose(obs[:, 0], g["obs"][t], rtol=0, atol=OBS_ATOL, err_msg=f"obs t={t}")
        np.testing.assert_allclose(rew[:, 0], g["rew"][t], rtol=REW_RTOL, atol=1e-6, err_msg=f"rew t={t}")
        np.testing.assert_allclose(emu.agent_p[:, 0], g["agent_p"][t], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(emu.vmag[perm, 0], g["volt"][t + 1], rtol=0, atol=V_ATOL)
        assert bool(done) == bool(g["done"][t])


def test_exogenous_recipes_agree():
    from oracle.exogenous import synthetic_exogenous_table
    from powergridworld_b200.agents.buildings.exogenous import synthetic_table
    np.testing.assert_array_equal(synthetic_exogenous_table(), synthetic_table())


def test_der123_scenario_tables_and_arithmetic_vs_oracle():
    """BASELINE C3 composition (123-bus-class feeder, 100 heterogeneous agents incl. loads with
    models 2 and 5): compiled tables + device arithmetic + NumPy Z-bus vs the oracle."""
    import warnings
    from tests.flatten import flat_obs, unflatten_action
    from tests.oracle_ns import ORACLE_NS as ONS, storage_socs_to_dict
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NS.MultiAgentEnv(**S.der123_scenario(NS, NS.OpenDSSSolver), _dry_run=True)
    assert (env.act_dim, env.obs_dim, len(env.agents)) == (240, 470, 100)
    f = env.pf_solver.feeder
    assert (f.nn, f.nb, f.nl) == (251, 85, 85)
    emu = EmulatedEnv(env)
    ref = ONS.MultiAgentEnv(**S.der123_scenario(ONS, ONS.OpenDSSSolver))
    rng = np.random.default_rng(0)
    soc = rng.uniform(10, 50, size=(env.num_storage, 1))
    o0 = emu.reset(soc)
    r0 = ref.reset(init_storage=storage_socs_to_dict(ref, soc[:, 0]))
    np.testing.assert_allclose(o0[:, 0], flat_obs(ref, r0), rtol=0, atol=1e-12)
    names = f.node_names
    for t in range(6):
        a = rng.uniform(-1, 1, size=(env.act_dim, 1))
        o, r, d = emu.step(a)
        ro, rr, rd, _ = ref.step(unflatten_action(ref, a[:, 0]))
        np.testing.assert_allclose(o[:, 0], flat_obs(ref, ro), rtol=0, atol=1e-12)
        np.testing.assert_allclose(r[:, 0], [rr[x.name] for x in ref.agents], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(emu.vmag[:, 0], [ref.voltages[n] for n in names], rtol=0, atol=1e-10)


def test_randomized_rosters_host_draws_and_tables_vs_reference_trace():
    """EVChargingEnv(randomize=True): the product's host side draws the rosters and the
    storage SOCs of a reset in the reference's RNG order and rebuilds the station tables in
    place; with the device arithmetic emulated this replays the reference's two episodes."""
    from tests.test_oracle_golden import _replay_randomized
    state = {}

    def make():
        env = NS.MultiAgentEnv(**S.randomized_ev_scenario(NS, NS.OpenDSSSolver), _dry_run=True)
        state["emu"] = EmulatedEnv(env)
        return env

    def reset(env):
        soc = env._reset_draws(None)                 # what reset_batch / reset_host do first
        assert soc.shape == (2, 1)
        return state["emu"].reset(soc, drawn=True)[:, 0]

    def step(env, a):
        obs, rew, done = state["emu"].step(a.reshape(-1, 1))
        return obs[:, 0], rew[:, 0], bool(done)

    _replay_randomized(make, reset, step, exact=False)
    env = state["emu"].env
    g = np.load(os.path.join(GOLD, "ev_randomized.npz"))
    rosters = [o._rows for o in env._b.objs if getattr(o, "randomize", False)]
    for k, r in enumerate(rosters):                  # last episode's draw, row for row
        np.testing.assert_array_equal(r, g[f"roster1_{k}"])


def _standalone_cases():
    from tests.component_cases import load_cases
    g, meta = load_cases()
    # a component on its own has no feeder: cases with voltage inputs or an injected set point are
    # covered inside full envs (tests/test_gpu_api.py::test_random_component_configurations_match_oracle)
    return g, [(ci, m) for ci, m in enumerate(meta)
               if not m["ext_keys"] and "p_setpoint" not in m["reset_kw"]]


def test_standalone_components_on_random_configurations_vs_reference_traces():
    """tests/golden/component_configs.npz (recorded from the unmodified reference): the spec
    compiler's tables + the device arithmetic for every case that needs no grid input --
    1..30-minute EV steps, episode caps, storage ranges / efficiencies / control intervals with
    drawn and explicit initial SOC, the three PV profiles, building observation subsets."""
    import pandas as pd
    from tests.component_cases import build_component
    g, cases = _standalone_cases()
    assert len(cases) >= 28
    common = {"start_time": "01-01-2021 00:00:00", "end_time": "01-01-2031 00:00:00",
              "control_timedelta": pd.Timedelta(300, "s")}
    for ci, m in cases:
        me = build_component(getattr(NS, m["cls"]), m["cfg"])
        me.name = me.name or "x"
        env = NS.MultiAgentEnv(common_config=common, pf_config=None, _dry_run=True, agents=[
            {"name": me.name, "bus": None, "cls": lambda name, me=me, **kw: me, "config": {}}])
        emu = EmulatedEnv(env)
        np.random.seed(m["seed"])
        if "init_storage" in m["reset_kw"]:
            o0 = emu.reset(np.array([[m["reset_kw"]["init_storage"]]]))
        else:
            o0 = emu.reset(env._reset_draws(None), drawn=True)
        if g[f"obs0_{ci}"].size:                     # PVEnv.reset returns nothing in the reference
            np.testing.assert_allclose(o0[:, 0], g[f"obs0_{ci}"], rtol=1e-13, atol=1e-13,
                                       err_msg=f"case {ci} ({m['cls']}) reset")
        A = g[f"act_{ci}"]
        assert env.episode_length >= A.shape[0] or not g[f"done_{ci}"][-1]
        for t in range(A.shape[0]):
            o, r, d = emu.step(A[t].reshape(-1, 1))
            np.testing.assert_allclose(o[:, 0], g[f"obs_{ci}"][t], rtol=1e-12, atol=1e-12,
                                       err_msg=f"case {ci} ({m['cls']}) {m['cfg']} t={t}")
            np.testing.assert_allclose(r[0, 0], g[f"rew_{ci}"][t], rtol=1e-11, atol=1e-14)
            assert bool(d) == bool(g[f"done_{ci}"][t]), (ci, t)


def test_agent_on_unknown_load_name_is_ignored_by_the_power_flow_like_the_reference():
    """opendss.py:115-129 walks the feeder's own load names: an agent whose ``bus`` is not one of
    them still steps, but its power never reaches the circuit.  Same here (with a warning); the
    trajectories agree with the oracle, which follows the reference."""
    from tests.flatten import flat_obs, unflatten_action
    from tests.oracle_ns import ORACLE_NS as ONS, storage_socs_to_dict

    def scen(ns):
        cfg = S.heterogeneous_scenario(ns, ns.OpenDSSSolver, 0.65)
        cfg["agents"][2]["bus"] = "nowhere"          # the EV station
        return cfg
    with pytest.warns(UserWarning, match="no load named 'nowhere'"):
        env = NS.MultiAgentEnv(**scen(NS), _dry_run=True)
    assert env._agent_recs[2].load_slot == -1 and env._agent_recs[0].load_slot >= 0
    ref = ONS.MultiAgentEnv(**scen(ONS))
    emu = EmulatedEnv(env)
    soc = np.array([[30.0]])
    o0 = emu.reset(soc)
    r0 = ref.reset(init_storage=storage_socs_to_dict(ref, soc[:, 0]))
    np.testing.assert_allclose(o0[:, 0], flat_obs(ref, r0), rtol=0, atol=OBS_ATOL)
    rng = np.random.default_rng(3)
    for t in range(140):                             # into the hours when vehicles charge
        a = rng.uniform(-1, 1, size=(env.act_dim, 1))
        o, r, _ = emu.step(a)
        ro, rr, _, _ = ref.step(unflatten_action(ref, a[:, 0]))
        np.testing.assert_allclose(o[:, 0], flat_obs(ref, ro), rtol=0, atol=OBS_ATOL, err_msg=f"t={t}")
        np.testing.assert_allclose(r[:, 0], [rr[x.name] for x in ref.agents], rtol=REW_RTOL, atol=REW_ATOL)
    assert emu.agent_p[2, 0] > 0.0                   # the station does draw power


def test_composite_on_its_own_replays_reference_trace():
    """gridworld/base.py:108-172 -- a MultiComponentEnv stepped outside any MultiAgentEnv (the
    reference's tests/conftest.py fixture: 6-entry building observation + PV + storage, raw
    spaces, actions slightly outside the spaces): tests/golden/composite_standalone.npz through
    the spec compiler and the device arithmetic; the storage SOC is the reference's own draw
    (np.random.seed(3) before the constructor and reset, as in the recording)."""
    import pandas as pd
    g = np.load(os.path.join(GOLD, "composite_standalone.npz"))
    np.random.seed(3)
    me = NS.MultiComponentEnv(name="house", components=S.test_multicomponent_components(NS))
    common = {"start_time": "01-01-2021 00:00:00", "end_time": "01-01-2031 00:00:00",
              "control_timedelta": pd.Timedelta(300, "s")}
    env = NS.MultiAgentEnv(common_config=common, pf_config=None, _dry_run=True, agents=[
        {"name": "house", "bus": None, "cls": lambda name, **kw: me, "config": {}}])
    emu = EmulatedEnv(env)
    soc = env._reset_draws(None)
    np.testing.assert_array_equal(soc[:, 0], g["init_soc"])
    o0 = emu.reset(soc, drawn=True)
    np.testing.assert_allclose(o0[:, 0], g["obs0"], rtol=0, atol=1e-13)
    T = g["actions"].shape[0]
    assert T == 285                                  # the building ends the episode first
    for t in range(T):
        o, r, d = emu.step(g["actions"][t].reshape(-1, 1))
        np.testing.assert_allclose(o[:, 0], g["obs"][t], rtol=1e-12, atol=1e-12, err_msg=f"t={t}")
        np.testing.assert_allclose(r[0, 0], g["rew"][t], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(emu.agent_p[0, 0], g["real_power"][t], rtol=1e-13, atol=1e-13)
        assert bool(d) == bool(g["done"][t])


def test_three_consecutive_episodes_on_one_env_vs_reference_trace():
    """tests/golden/heterogeneous_3episodes.npz through the spec compiler and the device
    arithmetic: the building's state vector survives pgw_reset like the reference's does."""
    g = np.load(os.path.join(GOLD, "heterogeneous_3episodes.npz"))
    env = NS.MultiAgentEnv(**S.heterogeneous_scenario(NS, NS.OpenDSSSolver, 0.65, max_episode_steps=60),
                           _dry_run=True)
    emu = EmulatedEnv(env)
    for ep in range(int(g["episodes"])):
        o0 = emu.reset(g[f"init_soc{ep}"].reshape(-1, 1))
        np.testing.assert_allclose(o0[:, 0], g[f"obs0_{ep}"], rtol=0, atol=OBS_ATOL, err_msg=f"ep={ep}")
        A = g[f"actions{ep}"]
        for t in range(A.shape[0]):
            o, r, d = emu.step(A[t].reshape(-1, 1))
            np.testing.assert_allclose(o[:, 0], g[f"obs{ep}"][t], rtol=0, atol=OBS_ATOL, err_msg=f"ep={ep} t={t}")
            np.testing.assert_allclose(r[:, 0], g[f"rew{ep}"][t], rtol=REW_RTOL, atol=REW_ATOL)
        assert bool(d)


def test_randomised_rosters_in_a_batch_are_per_env_and_match_the_oracle_per_env():
    """A batch of three envs with randomised stations: every env instance draws its OWN roster at
    every reset, like every instance of the reference's class (ev_charging_env.py:154-157); with
    explicit per-env storage SOCs and per-env actions every env replays an oracle env that was
    handed the same roster (two resets = two draws per env)."""
    from tests.flatten import flat_obs, unflatten_action
    from tests.oracle_ns import ORACLE_NS as ONS, storage_socs_to_dict
    E = 3
    env = NS.MultiAgentEnv(**S.randomized_ev_scenario(NS, NS.OpenDSSSolver), num_envs=E, _dry_run=True)
    stations = [o for o in env._b.objs if getattr(o, "randomize", False)]
    assert all(o._per_env for o in stations)
    assert all(c.flags & N.F_EV_PER_ENV for c in env._b.comps if c.type == N.EV)
    emu = EmulatedEnv(env)
    refs = [ONS.MultiAgentEnv(**S.randomized_ev_scenario(ONS, ONS.OpenDSSSolver)) for _ in range(E)]
    rng = np.random.default_rng(0)
    seen = []
    for ep in range(2):
        soc = rng.uniform(10, 45, size=(env.num_storage, E))
        np.random.seed(40 + ep)
        assert env._reset_draws(soc) is soc          # explicit SOCs: only the rosters are drawn
        o0 = emu.reset(soc)
        rosters = [np.array(o._rows) for o in stations]            # [E, n] each
        assert all(r.shape == (E, o.num_vehicles) for r, o in zip(rosters, stations))
        assert not np.array_equal(rosters[0][0], rosters[0][1])    # env 0 and env 1 park different vehicles
        seen.append(rosters)
        orig = np.random.choice
        for e, r in enumerate(refs):
            q = [x[e] for x in rosters]
            np.random.choice = lambda n, size=None, replace=True: q.pop(0)
            try:
                r0 = r.reset(init_storage=storage_socs_to_dict(r, soc[:, e]))
            finally:
                np.random.choice = orig
            np.testing.assert_allclose(o0[:, e], flat_obs(r, r0), rtol=0, atol=OBS_ATOL)
        for t in range(285):
            a = rng.uniform(-1, 1, size=(env.act_dim, E))
            o, rew, _ = emu.step(a)
            for e, r in enumerate(refs):
                ro, rr, _, _ = r.step(unflatten_action(r, a[:, e]))
                np.testing.assert_allclose(o[:, e], flat_obs(r, ro), rtol=0, atol=OBS_ATOL,
                                           err_msg=f"ep={ep} t={t} env={e}")
                np.testing.assert_allclose(rew[:, e], [rr[x.name] for x in r.agents],
                                           rtol=REW_RTOL, atol=REW_ATOL)
    assert not np.array_equal(seen[0][0], seen[1][0])           # a new draw per reset


def test_per_env_roster_path_equals_the_shared_table_path_bit_for_bit():
    """The two EV code paths -- per-event window lists compiled on the host (one env / shared roster)
    and per-env window words evaluated on the device (PGW_F_EV_PER_ENV) -- on the SAME roster, SOCs
    and actions: observations, rewards and states bit for bit, over a whole day."""
    E = 4
    batch = NS.MultiAgentEnv(**S.randomized_ev_scenario(NS, NS.OpenDSSSolver), num_envs=E, _dry_run=True)
    single = NS.MultiAgentEnv(**S.randomized_ev_scenario(NS, NS.OpenDSSSolver), num_envs=1, _dry_run=True)
    sb = [o for o in batch._b.objs if getattr(o, "randomize", False)]
    ss = [o for o in single._b.objs if getattr(o, "randomize", False)]
    assert all(o._per_env for o in sb) and not any(o._per_env for o in ss)
    rng = np.random.default_rng(11)
    np.random.seed(123)
    soc1 = rng.uniform(10, 45, size=(single.num_storage, 1))
    single._reset_draws(soc1)                         # draws one roster per station
    for ob, os_ in zip(sb, ss):
        ob._rows = np.tile(os_._rows, (E, 1))         # every env of the batch parks the same vehicles
    batch._rebuild_roster_tables()
    eb, es = EmulatedEnv(batch), EmulatedEnv(single)
    ob0, os0 = eb.reset(np.tile(soc1, (1, E))), es.reset(soc1)
    for e in range(E):
        np.testing.assert_array_equal(ob0[:, e], os0[:, 0])
    for t in range(285):
        a = rng.uniform(-1, 1, size=(single.act_dim, 1))
        o_b, r_b, _ = eb.step(np.tile(a, (1, E)))
        o_s, r_s, _ = es.step(a)
        for e in range(E):
            np.testing.assert_array_equal(o_b[:, e], o_s[:, 0], err_msg=f"t={t} env={e}")
            np.testing.assert_array_equal(r_b[:, e], r_s[:, 0])
    for ob, os_ in zip(sb, ss):                       # remaining energy per vehicle
        np.testing.assert_array_equal(eb.sd[ob._slot["sd"][0]:ob._slot["sd"][0] + ob.num_vehicles, 0],
                                      es.sd[os_._slot["sd"][0]:os_._slot["sd"][0] + os_.num_vehicles, 0])


def test_scenario_factories_read_like_the_reference():
    """gridworld/scenarios/{buildings,heterogeneous}.py: same signatures, same class names inside
    the config dicts (``OpenDSSSolver``, ``ThisPVEnv``, ...), and the configs build."""
    import inspect
    from powergridworld_b200.scenarios import buildings as PB, heterogeneous as PH
    assert str(inspect.signature(PB.make_env_config)) == \
        "(building_config=None, pv_config=None, storage_config=None, system_load_rescale_factor=0.65, num_buildings=3)"
    assert str(inspect.signature(PH.make_env_config)) == "(system_load_rescale_factor=0.65, rescale_spaces=True)"
    cfg = PH.make_env_config()
    assert cfg["pf_config"]["cls"].__name__ == "OpenDSSSolver"
    assert [a["cls"].__name__ for a in cfg["agents"]] == ["MultiComponentEnv", "ThisPVEnv", "EVChargingEnv"]
    assert [c["cls"].__name__ for c in cfg["agents"][0]["config"]["components"]] == \
        ["FiveZoneROMThermalEnergyEnv", "PVEnv", "EnergyStorageEnv"]
    env = NS.MultiAgentEnv(**cfg, _dry_run=True)
    assert env.episode_length == 286 and (env.act_dim, env.obs_dim) == (10, 25)
    cfg = PB.make_env_config(building_config={},     # (None fails in the reference too: cls(**None))
                             pv_config={"profile_csv": "pv_profile.csv", "scaling_factor": 40.},
                             storage_config={"max_power": 15., "storage_range": (3., 50.)})
    assert cfg["pf_config"]["cls"].__name__ == "OpenDSSSolver" and len(cfg["agents"]) == 3
    env = NS.CoordinatedMultiBuildingControlEnv(**cfg, _dry_run=True)
    assert (env.act_dim, env.obs_dim) == (24, 51)
