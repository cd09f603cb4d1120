#!/usr/bin/env python
"""Record the step METAS (4th return value of MultiAgentEnv.step) of the UNMODIFIED reference
(authoring container only):   python tests/golden/make_golden_meta.py

Same harness, scenarios, seeds and action draws as make_golden.py; stores for the first T steps
of c0_buildings and heterogeneous a [T, K] float array and the K key paths
("agent/component/key" or "agent/key"; array-valued entries get "/i")."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle.flatten import action_layout, unflatten_action  # noqa: E402
from oracle.powerflow import OracleOpenDSSSolver  # noqa: E402
from oracle.ref_harness import load_reference, quiet_stdout, reference_namespace  # noqa: E402
from powergridworld_b200.scenarios import catalog as S  # noqa: E402
from tests.golden.make_golden import draw_actions, storage_socs  # noqa: E402

T = 16


def flatten_meta(meta, prefix=""):
    out = {}
    for k, v in meta.items():
        path = f"{prefix}{k}"
        if isinstance(v, dict):
            out.update(flatten_meta(v, path + "/"))
        else:
            a = np.asarray(v, dtype=np.float64).reshape(-1)
            if a.size == 1 and not isinstance(v, np.ndarray):
                out[path] = float(a[0])
            else:
                for i, x in enumerate(a):
                    out[f"{path}/{i}"] = float(x)
    return out


def record(name, env_cls, cfg, ref):
    with quiet_stdout():
        np.random.seed(0)
        env = env_cls(**cfg)
        env.reset()
    socs = storage_socs(ref, env)
    layout = action_layout(env)
    rng = np.random.default_rng(1234)
    rows, keys, acts = [], None, []
    for t in range(T):
        a = draw_actions(layout, rng)
        with quiet_stdout():
            _, _, _, meta = env.step(unflatten_action(env, a))
        flat = flatten_meta(meta)
        if keys is None:
            keys = list(flat.keys())
        assert list(flat.keys()) == keys
        rows.append([flat[k] for k in keys])
        acts.append(a)
    np.savez_compressed(os.path.join(HERE, f"meta_{name}.npz"), keys=np.array(keys), meta=np.array(rows),
                        actions=np.array(acts), init_soc=socs)
    print(f"meta_{name}: T={T}, {len(keys)} entries, e.g. {keys[:4]} ... {keys[-3:]}")


def main():
    ref = load_reference()
    ns = reference_namespace(ref)
    record("c0_buildings", ns.CoordinatedMultiBuildingControlEnv,
           S.buildings_scenario(ns, OracleOpenDSSSolver, 1.2), ref)
    record("heterogeneous", ns.MultiAgentEnv, S.heterogeneous_scenario(ns, OracleOpenDSSSolver, 0.65), ref)


if __name__ == "__main__":
    main()
