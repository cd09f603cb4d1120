#!/usr/bin/env python
"""Record golden traces from the UNMODIFIED reference (authoring container only).

    python tests/golden/make_golden.py

Imports /root/reference/gridworld through oracle/ref_harness.py (gym stub, pandas
copy-on-write shim, synthetic exogenous table), builds the reference's own
``MultiAgentEnv`` / ``CoordinatedMultiBuildingControlEnv`` for each scenario of
tests/scenarios.py with this repo's CPU power-flow oracle plugged in through the
reference's ``pf_config["cls"]`` hook (gridworld/multiagent_env.py:80; the real
OpenDSS engine is not installable here), steps one full episode under seeded
actions and stores inputs and outputs as ``tests/golden/<scenario>.npz``:

  actions[T, act_dim]   flat, agent-major then component order
  init_soc[n_storage]   the SOCs the reference drew (np.random.seed(0))
  obs0[obs_dim]         reset observation
  obs[T, obs_dim], rew[T, A], done[T], agent_p[T, A], volt[T+1, n_nodes]
  node_names, act_dim, obs_dim

Also stores the three EV-station episode totals printed in
examples/envs/ev-charging.ipynb cells 5-7 next to what the reference computes
here (they agree bit for bit).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from tests import scenarios as S  # noqa: E402
from tests.flatten import flat_obs, unflatten_action, action_layout  # noqa: E402
from oracle.ref_harness import load_reference, quiet_stdout, reference_namespace  # noqa: E402
from oracle.powerflow import OracleOpenDSSSolver  # noqa: E402


def storage_socs(ref, env):
    out = []
    for a in env.agents:
        for e in getattr(a, "envs", [a]):
            if isinstance(e, ref.EnergyStorageEnv):
                out.append(e.current_storage)
    return np.array(out, dtype=np.float64)


def draw_actions(layout, rng):
    """U(-1.2, 1.2) on rescaled dims (exercises the clip), U(low, high) on raw ones."""
    parts = []
    for _, _, low, high, rescaled in layout:
        if rescaled:
            parts.append(rng.uniform(-1.2, 1.2, size=low.shape))
        else:
            parts.append(rng.uniform(low, high))
    return np.concatenate(parts)


def record(name, env_cls, cfg, ref, seed=1234):
    with quiet_stdout():
        np.random.seed(0)
        env = env_cls(**cfg)
        obs0 = env.reset()
    socs = storage_socs(ref, env)
    layout = action_layout(env)
    rng = np.random.default_rng(seed)
    names = list(env.pf_solver.get_bus_voltages().keys())
    A, O, R, D, P, V = [], [], [], [], [], [np.array([env.voltages[k] for k in names])]
    done = False
    while not done:
        a = draw_actions(layout, rng)
        with quiet_stdout():
            ob, rew, dn, _ = env.step(unflatten_action(env, a))
        A.append(a)
        O.append(flat_obs(env, ob))
        R.append(np.array([rew[ag.name] for ag in env.agents], dtype=np.float64))
        D.append(dn["__all__"])
        P.append(np.array(env.history["agent_power_p"][-1], dtype=np.float64))
        V.append(np.array([env.voltages[k] for k in names]))
        done = dn["__all__"]
    out = dict(actions=np.array(A), init_soc=socs, obs0=flat_obs(env, obs0), obs=np.array(O),
               rew=np.array(R), done=np.array(D), agent_p=np.array(P), volt=np.array(V),
               node_names=np.array(names))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: T={len(A)} act_dim={out['actions'].shape[1]} obs_dim={out['obs'].shape[1]} "
          f"vmin={out['volt'].min():.4f} vmax={out['volt'].max():.4f}")


def record_randomized(ref, ns, pf, episodes=2, seed=4321):
    """Two episodes of tests/scenarios.py::randomized_ev_scenario.  np.random.seed(100 + ep) is
    set right before each reset; the file keeps what the reference then drew (storage SOCs,
    and -- through the trace -- the rosters), so a replay checks the draws themselves and
    their order, not only the dynamics."""
    with quiet_stdout():
        env = ns.MultiAgentEnv(**S.randomized_ev_scenario(ns, pf))
    layout = action_layout(env)
    rng = np.random.default_rng(seed)
    out = {}
    for ep in range(episodes):
        np.random.seed(100 + ep)
        with quiet_stdout():
            obs0 = env.reset()
        socs = storage_socs(ref, env)               # what this reset drew
        A, O, R, P = [], [], [], []
        done = False
        while not done:
            a = draw_actions(layout, rng)
            with quiet_stdout():
                ob, rew, dn, _ = env.step(unflatten_action(env, a))
            A.append(a)
            O.append(flat_obs(env, ob))
            R.append(np.array([rew[ag.name] for ag in env.agents], dtype=np.float64))
            P.append(np.array(env.history["agent_power_p"][-1], dtype=np.float64))
            done = dn["__all__"]
        rosters = [np.asarray(e.df["index"].values) for a in env.agents
                   for e in getattr(a, "envs", [a]) if isinstance(e, ref.EVChargingEnv)]
        out.update({f"init_soc{ep}": socs, f"obs0_{ep}": flat_obs(env, obs0),
                    f"actions{ep}": np.array(A), f"obs{ep}": np.array(O), f"rew{ep}": np.array(R),
                    f"agent_p{ep}": np.array(P)})
        for k, r in enumerate(rosters):
            out[f"roster{ep}_{k}"] = r
    np.savez_compressed(os.path.join(HERE, "ev_randomized.npz"), episodes=episodes, **out)
    print(f"ev_randomized: {episodes} episodes, T={len(A)}, rosters "
          f"{[out[f'roster0_{k}'][:4].tolist() for k in range(len(rosters))]}")


def record_consecutive(ref, ns, pf, episodes=3, seed=77):
    """Three consecutive 59-step episodes of the heterogeneous scenario on ONE env object: what a
    reset keeps (the building's state vector is not zeroed, five_zone_rom_env.py:147-180) and
    what it restores (EV energies, PV index, the storage SOC it draws: np.random.seed(10 + ep)
    before each reset)."""
    with quiet_stdout():
        np.random.seed(0)
        env = ns.MultiAgentEnv(**S.heterogeneous_scenario(ns, pf, 0.65, max_episode_steps=60))
    layout = action_layout(env)
    rng = np.random.default_rng(seed)
    out = {}
    for ep in range(episodes):
        np.random.seed(10 + ep)
        with quiet_stdout():
            obs0 = env.reset()
        socs = storage_socs(ref, env)
        A, O, R = [], [], []
        done = False
        while not done:
            a = draw_actions(layout, rng)
            with quiet_stdout():
                ob, rew, dn, _ = env.step(unflatten_action(env, a))
            A.append(a); O.append(flat_obs(env, ob))
            R.append(np.array([rew[ag.name] for ag in env.agents], dtype=np.float64))
            done = dn["__all__"]
        out.update({f"init_soc{ep}": socs, f"obs0_{ep}": flat_obs(env, obs0), f"actions{ep}": np.array(A),
                    f"obs{ep}": np.array(O), f"rew{ep}": np.array(R)})
    np.savez_compressed(os.path.join(HERE, "heterogeneous_3episodes.npz"), episodes=episodes, **out)
    print(f"heterogeneous_3episodes: {episodes} x {len(A)} steps")


def ev_totals(ref):
    cfg = {"num_vehicles": 100, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
           "peak_threshold": 250., "vehicle_multiplier": 5., "rescale_spaces": False}
    notebook = {"high": -934170.2851237846, "low": -2659771.95782906,
                "const0.8": -1161670.9270816303}
    got = {}
    for k, pol in [("high", lambda e: e.action_space.high), ("low", lambda e: e.action_space.low),
                   ("const0.8", lambda e: np.array([.8]))]:
        env = ref.EVChargingEnv(**cfg)
        env.reset()
        done, tot = False, 0.
        while not done:
            _, r, done, _ = env.step(pol(env))
            tot += r
        got[k] = tot * env.reward_scale
        assert got[k] == notebook[k], (k, got[k], notebook[k])
    np.savez(os.path.join(HERE, "ev_totals.npz"),
             keys=np.array(list(notebook)), notebook=np.array([notebook[k] for k in notebook]),
             reference_here=np.array([got[k] for k in notebook]))
    print("EV totals reproduced bit-exactly:", got)


def main():
    ref = load_reference()
    ns = reference_namespace(ref)
    pf = OracleOpenDSSSolver
    record("c0_buildings", ns.CoordinatedMultiBuildingControlEnv,
           S.buildings_scenario(ns, pf, system_load_rescale_factor=1.2), ref)
    record("heterogeneous", ns.MultiAgentEnv, S.heterogeneous_scenario(ns, pf, 0.65), ref)
    record("heterogeneous_max250", ns.MultiAgentEnv,
           S.heterogeneous_scenario(ns, pf, 0.6, max_episode_steps=250), ref)
    record("test_heterogeneous", ns.MultiAgentEnv, S.test_heterogeneous_scenario(ns, pf), ref)
    record_randomized(ref, ns, pf)
    record_consecutive(ref, ns, pf)
    for v in S.TIME_BASE_VARIANTS:
        record("timebase_" + v, ns.MultiAgentEnv, S.time_base_scenario(ns, pf, v), ref)
    ev_totals(ref)


if __name__ == "__main__":
    main()
