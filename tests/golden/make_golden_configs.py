#!/usr/bin/env python
"""Record the UNMODIFIED reference's component classes on seeded random configurations
(authoring container only; imports /root/reference through oracle/ref_harness.py):

    python tests/golden/make_golden_configs.py   ->  tests/golden/component_configs.npz

Every case is one component stepped on its own, the way the reference's tests/agents/*.py do:
constructor kwargs (JSON), reset kwargs, per-step action and external kwargs, and what the
reference returned (reset observation, then obs / reward / done per step).  The oracle classes
are then held to these traces bit for bit (tests/test_oracle_golden.py), which pins them on
parameter ranges the shipped scenarios never touch: EV stations with 1..30-minute steps and
episode caps, storages with every efficiency / range / control interval (drawn and explicit
initial SOC), all three PV profiles, buildings with random observation subsets and bounds,
voltage and set-point inputs, reward weights."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle.ref_harness import load_reference, quiet_stdout  # noqa: E402
from tests.component_cases import build_component as build  # noqa: E402

BUILDING_KEYS = ["zone_temp", "zone_upper_viol", "zone_lower_viol", "comfort_lower", "comfort_upper",
                 "outdoor_temp", "p_setpoint", "p_consumed", "time_of_day", "bus_voltage",
                 "min_voltage", "max_voltage"]


def cases(default_obs_config):
    out = []
    for i in range(10):
        r = np.random.default_rng(100 + i)
        out.append(("EVChargingEnv", dict(
            num_vehicles=int(r.integers(1, 120)), minutes_per_step=int(r.choice([1, 5, 10, 15, 30])),
            max_charge_rate_kw=float(r.uniform(2, 30)), peak_threshold=float(r.uniform(5, 300)),
            vehicle_multiplier=float(r.choice([1, 2, 5, 40])), rescale_spaces=bool(r.integers(2)),
            unserved_penalty=float(r.uniform(0, 2)), peak_penalty=float(r.uniform(0, 2)),
            reward_scale=float(r.choice([1e5, 1.0, 1e3])),
            max_episode_steps=(None if r.integers(2) else int(r.integers(5, 400)))), {}, {}, 300))
    for i in range(10):
        r = np.random.default_rng(200 + i)
        lo, hi = float(r.uniform(0, 10)), float(r.uniform(20, 90))
        out.append(("EnergyStorageEnv", dict(
            storage_range=[lo, hi], initial_storage_mean=float(r.uniform(lo, hi)),
            initial_storage_std=float(r.uniform(0, 10)), charge_efficiency=float(r.uniform(.6, 1)),
            discharge_efficiency=float(r.uniform(.6, 1)), max_power=float(r.uniform(1, 60)),
            max_episode_steps=int(r.integers(3, 300)),
            control_timedelta_s=int(r.choice([60, 300, 900])), rescale_spaces=bool(r.integers(2)), name="s"),
            ({} if i % 2 else {"init_storage": float(i * 7 % 100)}), {}, 300))
    for i in range(8):
        r = np.random.default_rng(300 + i)
        out.append(("PVEnv", dict(
            profile_csv=str(r.choice(["pv_profile.csv", "off-peak.csv", "constant.csv"])),
            scaling_factor=float(r.uniform(.5, 80)), rescale_spaces=bool(r.integers(2)),
            max_episode_steps=(None if r.integers(2) else int(r.integers(5, 300))), name="p"), {}, {}, 300))
    for i in range(12):
        r = np.random.default_rng(500 + i)
        keys = [k for k in BUILDING_KEYS if r.integers(2)] or ["zone_temp"]
        obs_config = {}
        for k in keys:
            lo, hi = default_obs_config[k]
            if r.integers(2) and np.isfinite(lo) and np.isfinite(hi):
                w = hi - lo
                lo, hi = lo + 0.1 * w * r.uniform(), hi - 0.1 * w * r.uniform()
            obs_config[k] = [float(lo), float(hi)]
        cfg = dict(start_time="08-12-2020 00:00:00", end_time="08-13-2020 00:00:00",
                   rescale_spaces=bool(r.integers(2)), name="b", obs_config=obs_config)
        if r.integers(2):
            cfg["reward_structure"] = {"alpha": float(r.uniform(0, 1))}
        ext = {k: True for k in ("bus_voltage", "min_voltage", "max_voltage") if k in keys}
        fixed = {"p_setpoint": float(r.uniform(0, 50))} if ("p_setpoint" in keys and r.integers(2)) else {}
        out.append(("FiveZoneROMThermalEnergyEnv", cfg, fixed, ext, 80))
    return out


def main():
    ref = load_reference()
    from gridworld.agents.buildings.obs_space import DEFAULT_OBS_CONFIG
    out, meta = {}, []
    rng = np.random.default_rng(2024)
    for ci, (cls, cfg, fixed_kw, ext_keys, T) in enumerate(cases(DEFAULT_OBS_CONFIG)):
        with quiet_stdout():
            env = build(getattr(ref, cls), cfg)
        lo = np.asarray(env.action_space.low, float)
        hi = np.asarray(env.action_space.high, float)
        vr = np.random.default_rng(9000 + ci)
        ext0 = {k: float(vr.uniform(0.93, 1.06)) for k in ext_keys}
        np.random.seed(ci)
        with quiet_stdout():
            r0 = env.reset(**fixed_kw, **ext0)
        o0 = r0[0] if isinstance(r0, tuple) else r0
        A, X, O, R, D = [], [], [], [], []
        for t in range(T):
            a = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo))
            ext = [float(vr.uniform(0.93, 1.06)) for _ in ext_keys]
            with quiet_stdout():
                ob, rew, done, _ = env.step(a, **dict(zip(ext_keys, ext)),
                                            **({k: v for k, v in fixed_kw.items() if k == "p_setpoint"}))
            A.append(a); X.append(ext); O.append(np.asarray(ob, float)); R.append(float(rew)); D.append(bool(done))
            if done:
                break
        meta.append({"cls": cls, "cfg": cfg, "reset_kw": fixed_kw, "ext_keys": list(ext_keys),
                     "ext0": ext0, "seed": ci})
        out[f"obs0_{ci}"] = np.zeros(0) if o0 is None else np.asarray(o0, float)
        out[f"act_{ci}"], out[f"ext_{ci}"] = np.array(A), np.array(X).reshape(len(A), -1)
        out[f"obs_{ci}"], out[f"rew_{ci}"], out[f"done_{ci}"] = np.array(O), np.array(R), np.array(D)
    np.savez_compressed(os.path.join(HERE, "component_configs.npz"), meta=np.array(json.dumps(meta)), **out)
    print(f"component_configs: {len(meta)} cases, {sum(len(out[f'act_{i}']) for i in range(len(meta)))} steps")


if __name__ == "__main__":
    main()
