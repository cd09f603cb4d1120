#!/usr/bin/env python
"""Record the UNMODIFIED reference's ``MultiComponentEnv`` stepped on its own (gridworld/base.py:
108-172, the fixture of the reference's tests/conftest.py:113-148: building with a 6-entry
observation set + PV + storage, raw spaces) for one full episode (authoring container only):

    python tests/golden/make_golden_composite.py  ->  tests/golden/composite_standalone.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as M  # noqa: E402
from tests import scenarios as S  # noqa: E402


def main():
    ref = M.load_reference()
    ns = M.reference_namespace(ref)
    with M.quiet_stdout():
        np.random.seed(3)
        env = ns.MultiComponentEnv(name="house", components=S.test_multicomponent_components(ns))
        obs0, _ = env.reset()
    soc = [e for e in env.envs if hasattr(e, "current_storage")][0].current_storage
    flat = lambda o: np.concatenate([np.atleast_1d(np.asarray(o[e.name], float)) for e in env.envs])
    rng = np.random.default_rng(0)
    A, O, R, P, D = [], [], [], [], []
    done = False
    while not done:
        act = {e.name: rng.uniform(np.asarray(e.action_space.low, float) - 0.05,
                                   np.asarray(e.action_space.high, float) + 0.05) for e in env.envs}
        with M.quiet_stdout():
            ob, rew, done, _ = env.step(act)
        A.append(np.concatenate([np.atleast_1d(act[e.name]) for e in env.envs]))
        O.append(flat(ob)); R.append(float(rew)); P.append(float(env.real_power)); D.append(bool(done))
    np.savez_compressed(os.path.join(HERE, "composite_standalone.npz"), init_soc=np.array([soc]),
                        obs0=flat(obs0), actions=np.array(A), obs=np.array(O), rew=np.array(R),
                        real_power=np.array(P), done=np.array(D))
    print(f"composite_standalone: T={len(A)} obs_dim={len(O[0])} drawn SOC {soc:.4f}")


if __name__ == "__main__":
    main()
