#!/usr/bin/env python
"""Record the UNMODIFIED reference's Home-Steward house on seeded random parameter sets, TWO
consecutive episodes each (authoring container only):

    python tests/golden/make_golden_hs_configs.py   ->  tests/golden/hs_random_configs.npz

Pins what the three shipped-house traces do not: PV / storage / charger parameters and rescale
flags drawn at random (tests/scenarios_hs.py::parametrised), the initial SOC the reference draws
itself (np.random.seed before every reset; never clipped), and everything that survives a reset
(storage cost, meta state: energy_storage_env_hs.py:39, base_hs.py:53-61)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden_hs as MH  # noqa: E402
from oracle.components_hs import META_KEYS  # noqa: E402
from oracle.ref_harness import quiet_stdout  # noqa: E402
from tests import scenarios_hs as SH  # noqa: E402

N_CASES, EPISODES = 6, 2


def params(i):
    r = np.random.default_rng(700 + i)
    return dict(pv_scale=float(r.uniform(.3, 3)), max_power=float(r.uniform(2, 12)),
                lo=float(r.uniform(.5, 4)), hi=float(r.uniform(10, 30)), eta_c=float(r.uniform(.7, 1)),
                eta_d=float(r.uniform(.7, 1)), init_cost=float(r.uniform(0, .5)),
                mult=float(r.choice([1., 2., 3.])), rate=float(r.uniform(3, 15)),
                grid=float(r.uniform(60, 120)), rescale=[bool(x) for x in r.integers(0, 2, 4)])


def main():
    ns = MH.reference_hs_namespace()
    out, meta = {}, []
    for i in range(N_CASES):
        hp = params(i)
        with quiet_stdout():
            env = ns.HSMultiComponentEnv(**SH.parametrised(ns, hp))
        rng = np.random.default_rng(i)
        for ep in range(EPISODES):
            np.random.seed(50 + i + ep)
            with quiet_stdout():
                obs0 = env.reset()
            soc = [e for e in env.envs if hasattr(e, "current_storage")][0].current_storage
            A, O, R, P, M = [], [], [], [], []
            done = False
            while not done:
                a = np.array([SH.draw_action(e, rng) for e in env.envs])
                with quiet_stdout():
                    ob, rew, done, m = env.step({e.name: a[k:k + 1] for k, e in enumerate(env.envs)})
                A.append(a); O.append(MH.flat(env, ob)); R.append(rew); P.append(env.real_power)
                M.append([float(m[k]) for k in META_KEYS])
            key = f"{i}_{ep}"
            out["obs0_" + key], out["soc_" + key] = MH.flat(env, obs0), np.array([soc])
            out["act_" + key], out["obs_" + key] = np.array(A), np.array(O)
            out["rew_" + key], out["p_" + key], out["meta_" + key] = np.array(R, float), np.array(P, float), np.array(M)
        meta.append({"hp": hp, "seed": 50 + i})
    np.savez_compressed(os.path.join(HERE, "hs_random_configs.npz"), meta=np.array(json.dumps(meta)), **out)
    print(f"hs_random_configs: {N_CASES} houses x {EPISODES} episodes x {len(A)} steps")


if __name__ == "__main__":
    main()
