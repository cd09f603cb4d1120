"""The CPU oracle against golden traces recorded from the unmodified reference
(tests/golden/make_golden.py) and against the reference's own known answers."""
import os

import numpy as np
import pytest

from tests import scenarios as S
from tests.flatten import flat_obs, unflatten_action
from tests.oracle_ns import ORACLE_NS as NS, storage_socs_to_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "c0_buildings": lambda: (NS.CoordinatedMultiBuildingControlEnv,
                             S.buildings_scenario(NS, NS.OpenDSSSolver, 1.2)),
    "heterogeneous": lambda: (NS.MultiAgentEnv,
                              S.heterogeneous_scenario(NS, NS.OpenDSSSolver, 0.65)),
    "heterogeneous_max250": lambda: (NS.MultiAgentEnv, S.heterogeneous_scenario(
        NS, NS.OpenDSSSolver, 0.6, max_episode_steps=250)),
    "test_heterogeneous": lambda: (NS.MultiAgentEnv,
                                   S.test_heterogeneous_scenario(NS, NS.OpenDSSSolver)),
}


for _v in S.TIME_BASE_VARIANTS:                     # other control intervals / day windows
    CASES["timebase_" + _v] = (lambda v: lambda: (NS.MultiAgentEnv, S.time_base_scenario(
        NS, NS.OpenDSSSolver, v)))(_v)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_replays_reference_trace(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cls, cfg = CASES[name]()
    env = cls(**cfg)
    obs0 = env.reset(init_storage=storage_socs_to_dict(env, g["init_soc"]))
    # same float64 arithmetic in the same order: the oracle is expected to be exact
    np.testing.assert_array_equal(flat_obs(env, obs0), g["obs0"])
    names = [str(n) for n in g["node_names"]]
    np.testing.assert_array_equal(np.array([env.voltages[k] for k in names]), g["volt"][0])
    T = g["actions"].shape[0]
    for t in range(T):
        ob, rew, dn, _ = env.step(unflatten_action(env, g["actions"][t]))
        np.testing.assert_array_equal(flat_obs(env, ob), g["obs"][t], err_msg=f"obs t={t}")
        np.testing.assert_array_equal(
            np.array([rew[a.name] for a in env.agents]), g["rew"][t], err_msg=f"rew t={t}")
        np.testing.assert_array_equal(
            np.array([env.voltages[k] for k in names]), g["volt"][t + 1], err_msg=f"volt t={t}")
        assert dn["__all__"] == bool(g["done"][t])
    assert dn["__all__"]


def test_episode_lengths_match_reference_facts():
    # PV profile (287 rows) ends first -> 286; max_episode_steps=250 -> 249
    assert np.load(os.path.join(GOLD, "c0_buildings.npz"))["actions"].shape[0] == 286
    assert np.load(os.path.join(GOLD, "heterogeneous_max250.npz"))["actions"].shape[0] == 249


def test_ev_notebook_totals_bit_exact():
    """examples/envs/ev-charging.ipynb cells 5-7 (always-max, always-min, constant 0.8)."""
    g = np.load(os.path.join(GOLD, "ev_totals.npz"))
    np.testing.assert_array_equal(g["notebook"], g["reference_here"])
    cfg = {"num_vehicles": 100, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
           "peak_threshold": 250., "vehicle_multiplier": 5., "rescale_spaces": False}
    pol = {"high": lambda e: e.action_space.high, "low": lambda e: e.action_space.low,
           "const0.8": lambda e: np.array([.8])}
    for key, want in zip(g["keys"], g["notebook"]):
        env = NS.EVChargingEnv(**cfg)
        env.reset()
        done, tot, n = False, 0.0, 0
        while not done:
            _, r, done, _ = env.step(pol[str(key)](env))
            tot += r
            n += 1
        assert n == 286
        assert tot * env.reward_scale == want


def test_storage_episode_is_287_steps():
    env = NS.EnergyStorageEnv(name="s")
    env.reset(init_storage=30.0)
    n, done = 0, False
    while not done:
        _, _, done, _ = env.step(np.array([0.3]))
        n += 1
    assert n == 287


def _replay_randomized(make_env, reset, step, exact):
    """Shared by the oracle (here) and the product tiers: two episodes of
    tests/golden/ev_randomized.npz, np.random.seed(100 + ep) before each reset like the
    recording; the storages' drawn SOCs and the rosters must come out of the same RNG stream."""
    g = np.load(os.path.join(GOLD, "ev_randomized.npz"))
    env = make_env()
    cmp = np.testing.assert_array_equal if exact else \
        (lambda a, b, err_msg="": np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-7, err_msg=err_msg))
    for ep in range(int(g["episodes"])):
        np.random.seed(100 + ep)
        cmp(reset(env), g[f"obs0_{ep}"], err_msg=f"obs0 ep={ep}")
        A = g[f"actions{ep}"]
        for t in range(A.shape[0]):
            ob, rew, done = step(env, A[t])
            cmp(ob, g[f"obs{ep}"][t], err_msg=f"obs ep={ep} t={t}")
            cmp(rew, g[f"rew{ep}"][t], err_msg=f"rew ep={ep} t={t}")
        assert done


def test_oracle_replays_randomized_rosters():
    """EVChargingEnv(randomize=True): pandas' df.sample(n) == np.random.choice(N, n, False)."""
    def step(env, a):
        ob, rew, dn, _ = env.step(unflatten_action(env, a))
        return flat_obs(env, ob), np.array([rew[x.name] for x in env.agents]), dn["__all__"]
    _replay_randomized(
        lambda: NS.MultiAgentEnv(**S.randomized_ev_scenario(NS, NS.OpenDSSSolver)),
        lambda env: flat_obs(env, env.reset()), step, exact=True)


def test_oracle_matches_reference_on_random_component_configurations():
    """tests/golden/component_configs.npz: 40 seeded random configurations of the four component
    classes recorded from the unmodified reference (tests/golden/make_golden_configs.py), each
    stepped on its own with out-of-range actions and external voltage / set-point inputs."""
    from tests.component_cases import build_component, load_cases
    g, meta = load_cases()
    assert len(meta) == 40
    for ci, m in enumerate(meta):
        env = build_component(getattr(NS, m["cls"]), m["cfg"])
        np.random.seed(m["seed"])
        r0 = env.reset(**m["reset_kw"], **m["ext0"])
        o0 = r0[0] if isinstance(r0, tuple) else r0
        if o0 is None:
            assert g[f"obs0_{ci}"].size == 0
        else:
            np.testing.assert_array_equal(np.asarray(o0, float), g[f"obs0_{ci}"], err_msg=f"case {ci} reset")
        fixed = {k: v for k, v in m["reset_kw"].items() if k == "p_setpoint"}
        A = g[f"act_{ci}"]
        for t in range(A.shape[0]):
            ob, rew, done, _ = env.step(A[t], **dict(zip(m["ext_keys"], g[f"ext_{ci}"][t])), **fixed)
            np.testing.assert_array_equal(np.asarray(ob, float), g[f"obs_{ci}"][t],
                                          err_msg=f"case {ci} ({m['cls']}) t={t}")
            assert float(rew) == g[f"rew_{ci}"][t] and bool(done) == bool(g[f"done_{ci}"][t]), (ci, t)


def test_oracle_composite_on_its_own_replays_reference_trace():
    """tests/golden/composite_standalone.npz: MultiComponentEnv outside a MultiAgentEnv."""
    g = np.load(os.path.join(GOLD, "composite_standalone.npz"))
    np.random.seed(3)
    env = NS.MultiComponentEnv(name="house", components=S.test_multicomponent_components(NS))
    obs0, _ = env.reset()
    flat = lambda o: np.concatenate([np.atleast_1d(np.asarray(o[e.name], float)) for e in env.envs])
    np.testing.assert_array_equal(flat(obs0), g["obs0"])
    k = 0
    for t in range(g["actions"].shape[0]):
        act, k = {}, 0
        for e in env.envs:
            n = int(np.prod(e.action_space.shape))
            act[e.name] = g["actions"][t][k:k + n]
            k += n
        ob, rew, done, _ = env.step(act)
        np.testing.assert_array_equal(flat(ob), g["obs"][t], err_msg=f"t={t}")
        assert float(rew) == g["rew"][t] and float(env.real_power) == g["real_power"][t]
        assert bool(done) == bool(g["done"][t])


def test_oracle_replays_three_consecutive_episodes_on_one_env():
    """tests/golden/heterogeneous_3episodes.npz: what a reset keeps (the building's state vector,
    five_zone_rom_env.py:147-180) and what it restores, on the reference itself."""
    g = np.load(os.path.join(GOLD, "heterogeneous_3episodes.npz"))
    env = NS.MultiAgentEnv(**S.heterogeneous_scenario(NS, NS.OpenDSSSolver, 0.65, max_episode_steps=60))
    for ep in range(int(g["episodes"])):
        obs0 = env.reset(init_storage=storage_socs_to_dict(env, g[f"init_soc{ep}"]))
        np.testing.assert_array_equal(flat_obs(env, obs0), g[f"obs0_{ep}"], err_msg=f"ep={ep}")
        A = g[f"actions{ep}"]
        for t in range(A.shape[0]):
            ob, rew, dn, _ = env.step(unflatten_action(env, A[t]))
            np.testing.assert_array_equal(flat_obs(env, ob), g[f"obs{ep}"][t], err_msg=f"ep={ep} t={t}")
            np.testing.assert_array_equal(np.array([rew[a.name] for a in env.agents]), g[f"rew{ep}"][t])
        assert dn["__all__"] and A.shape[0] == 59
    # the episodes differ although the scenario restarts at the same clock: carried-over state
    assert not np.array_equal(g["obs0_0"], g["obs0_1"])
