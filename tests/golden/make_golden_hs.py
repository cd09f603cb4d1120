#!/usr/bin/env python
"""Record golden traces of the Home-Steward house env from the UNMODIFIED reference
(authoring container only).

    python tests/golden/make_golden_hs.py

Builds the reference's ``HSMultiComponentEnv`` (gridworld/base_hs.py) from its own
``make_env_config()`` (gridworld/scenarios/heterogeneous_hs.py:47-58) and from two variants
(tests/scenarios_hs.py), steps one full episode under seeded actions and stores
``tests/golden/hs_<variant>.npz``:

  actions[T, 4]   one action per component, component order
  obs0[obs_dim], obs[T, obs_dim], rew[T], done[T], real_power[T]
  meta[T, 6]      grid_cost, es_cost, grid_power, pv_power, es_power, pv_cost after the step
  telemetry[T, 4, 13]  the numbers of the per-device step_meta records (base_hs.py:158-164):
                  cost, reward, raw action, solar / es / grid power consumed, then the
                  device_custom_info values in their dict order (nan-padded)
  init_soc        the storage level the reference started from
"""
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle.components_hs import META_KEYS  # noqa: E402
from oracle.ref_harness import load_reference, quiet_stdout  # noqa: E402
from tests import scenarios_hs as SH  # noqa: E402


def reference_hs_namespace():
    """The reference's HS classes (plus the pd.read_json shim pandas 3 needs for the literal
    JSON string of ev_charging_env_hs.py:69)."""
    import types

    import pandas as pd
    load_reference()
    if not getattr(pd.read_json, "_pgw_patched", False):
        orig = pd.read_json

        def read_json(obj, *a, **k):
            if isinstance(obj, str) and obj.lstrip().startswith(("{", "[")):
                obj = io.StringIO(obj)
            return orig(obj, *a, **k)
        read_json._pgw_patched = True
        pd.read_json = read_json
    from gridworld.agents.devices import HSDevicesEnv
    from gridworld.agents.energy_storage.energy_storage_env_hs import HSEnergyStorageEnv
    from gridworld.agents.pv.pv_profile_env_hs import HSPVEnv
    from gridworld.agents.vehicles import HSEVChargingEnv
    from gridworld.base_hs import HSMultiComponentEnv
    return types.SimpleNamespace(HSPVEnv=HSPVEnv, HSEnergyStorageEnv=HSEnergyStorageEnv,
                                 HSEVChargingEnv=HSEVChargingEnv, HSDevicesEnv=HSDevicesEnv,
                                 HSMultiComponentEnv=HSMultiComponentEnv, _is_reference=True)


def flat(env, obs):
    return np.concatenate([np.asarray(obs[e.name], dtype=np.float64).ravel() for e in env.envs])


def record(name, ns, seed):
    cfg = SH.VARIANTS[name](ns)
    with quiet_stdout():
        np.random.seed(0)
        env = ns.HSMultiComponentEnv(**cfg)
        obs0 = env.reset()
    storage = [e for e in env.envs if hasattr(e, "current_storage")]
    init_soc = np.array([storage[0].current_storage if storage else np.nan])
    rng = np.random.default_rng(seed)
    A, O, R, D, P, M, TL = [], [], [], [], [], [], []
    done = False
    while not done:
        a = np.array([SH.draw_action(e, rng) for e in env.envs])
        with quiet_stdout():
            ob, rew, done, meta = env.step({e.name: a[k:k + 1] for k, e in enumerate(env.envs)})
        A.append(a); O.append(flat(env, ob)); R.append(rew); D.append(done)
        P.append(env.real_power); M.append([float(meta[k]) for k in META_KEYS])
        tel = np.full((len(env.envs), 13), np.nan)
        for k, rec in enumerate(meta["step_meta"]):
            vals = [rec["cost"], rec["reward"], np.asarray(rec["action"]).ravel()[0],
                    rec["solar_power_consumed"], rec["es_power_consumed"], rec["grid_power_consumed"]]
            vals += [float(v) for v in rec["device_custom_info"].values()]
            tel[k, :len(vals)] = vals
            assert rec["device_id"] == env.envs[k].name
        TL.append(tel)
    np.savez_compressed(os.path.join(HERE, f"hs_{name}.npz"), actions=np.array(A),
                        obs0=flat(env, obs0), obs=np.array(O), rew=np.array(R, dtype=np.float64),
                        done=np.array(D), real_power=np.array(P, dtype=np.float64),
                        meta=np.array(M), init_soc=init_soc, telemetry=np.array(TL))
    print(f"hs_{name}: T={len(A)} obs_dim={len(O[0])} total reward {np.sum(R):.6f} "
          f"min grid_power {np.min(np.array(M)[:, 2]):.3f}")


def main():
    ns = reference_hs_namespace()
    for k, name in enumerate(SH.VARIANTS):
        record(name, ns, 100 + k)


if __name__ == "__main__":
    main()
