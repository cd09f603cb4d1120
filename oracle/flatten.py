"""Flat <-> nested (dict) conversion shared by golden generation, the parity tests, smoke() and
bench.py's CPU arm (test infrastructure).

Order everywhere: agents in list order, then components in list order
(the order of gridworld/multiagent_list_interface_env.py:80-111)."""
import numpy as np


def _components(agent):
    return getattr(agent, "envs", None)


def action_layout(env):
    """[(agent, component|None, low, high, rescaled?)] in flat order."""
    out = []
    for ag in env.agents:
        comps = _components(ag)
        for e in (comps if comps is not None else [ag]):
            sp = e.action_space
            low = np.asarray(sp.low, dtype=np.float64).reshape(-1)
            high = np.asarray(sp.high, dtype=np.float64).reshape(-1)
            out.append((ag.name, e.name if comps is not None else None, low, high,
                        bool(getattr(e, "rescale_spaces", True))))
    return out


def unflatten_action(env, flat):
    act, k = {}, 0
    for ag in env.agents:
        comps = _components(ag)
        if comps is None:
            n = int(np.prod(ag.action_space.shape))
            act[ag.name] = np.array(flat[k:k + n])
            k += n
        else:
            act[ag.name] = {}
            for e in comps:
                n = int(np.prod(e.action_space.shape))
                act[ag.name][e.name] = np.array(flat[k:k + n])
                k += n
    assert k == len(flat)
    return act


def flat_obs(env, obs):
    parts = []
    for ag in env.agents:
        o = obs[ag.name]
        if isinstance(o, dict):
            for e in _components(ag):
                parts.append(np.atleast_1d(np.asarray(o[e.name], dtype=np.float64)))
        else:
            parts.append(np.atleast_1d(np.asarray(o, dtype=np.float64)))
    return np.concatenate(parts)
