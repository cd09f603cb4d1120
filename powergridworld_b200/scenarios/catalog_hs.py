"""Home-Steward house scenarios shared by bench.py, the golden generator and the parity tests.

Each builder takes a namespace ``ns`` providing ``HSPVEnv, HSEnergyStorageEnv, HSEVChargingEnv,
HSDevicesEnv, HSMultiComponentEnv`` and returns the kwargs of ``HSMultiComponentEnv`` -- the
same dict builds the reference (tests/golden/make_golden_hs.py), the oracle and the product.

  shipped        gridworld/scenarios/heterogeneous_hs.py:47-58 make_env_config() on the packaged
                 copy of gridworld/scenarios/data/env_config.json (one vehicle, 48 kW grid limit)
  two_vehicles   the station of gridworld/agents/vehicles/vehicles_hs.csv (two vehicles, one
                 parked at reset), multiplier 2, a smaller battery, 2.5 x PV.  The grid limit stays
                 non-binding: with every source exhausted the reference divides by zero
                 (devices_env_hs.py:188, energy_storage_env_hs.py:233)
  raw_spaces     the shipped house with rescale_spaces=False everywhere
"""
import copy
import json
import os

import numpy as np
import pandas as pd

_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "hs_env_config.json")


def _load(ns):
    with open(_CFG) as fh:
        cfg = json.load(fh)
    for c in cfg["components"]:
        c["cls"] = getattr(ns, c["cls"])
    cfg["control_timedelta"] = pd.Timedelta(cfg["control_timedelta"])
    return cfg


def shipped(ns):
    return _load(ns)


def two_vehicles(ns):
    cfg = _load(ns)
    cfg["max_grid_power"] = 60
    by = {c["name"]: c["config"] for c in cfg["components"]}
    by["ev-charging"].update(profile_data={}, num_vehicles=2, vehicle_multiplier=2.0,
                             max_charge_rate_kw=7.0, unserved_penalty=0.5)
    by["storage"].update(max_power=10, storage_range=[2.0, 20.0], initial_storage_mean=12.0,
                         charge_efficiency=0.9, discharge_efficiency=0.85,
                         initial_storage_cost=0.3)
    by["pv"].update(scaling_factor=2.5)
    return cfg


def raw_spaces(ns):
    cfg = _load(ns)
    for c in cfg["components"]:
        c["config"]["rescale_spaces"] = False
    return cfg


VARIANTS = {"shipped": shipped, "two_vehicles": two_vehicles, "raw_spaces": raw_spaces}


def draw_action(comp, rng):
    """U(-1.2, 1.2) on rescaled components (exercises the clip), U(low, high) on raw ones."""
    if comp.rescale_spaces:
        return float(rng.uniform(-1.2, 1.2))
    return float(rng.uniform(comp._action_space.low[0], comp._action_space.high[0]))


def clone(cfg):
    out = copy.copy(cfg)
    out["components"] = [dict(c, config=copy.deepcopy(c["config"])) for c in cfg["components"]]
    return out


def parametrised(ns, hp):
    """two_vehicles with the numeric knobs of ``hp`` (pv_scale, max_power, lo, hi, eta_c, eta_d,
    init_cost, mult, rate, grid, rescale[4]) -- shared by the property test, the random-config
    goldens (tests/golden/make_golden_hs_configs.py) and their replays."""
    cfg = two_vehicles(ns)
    cfg["max_grid_power"] = hp["grid"]
    by = {c["name"]: c["config"] for c in cfg["components"]}
    by["pv"].update(scaling_factor=hp["pv_scale"], rescale_spaces=bool(hp["rescale"][0]))
    by["storage"].update(max_power=hp["max_power"], storage_range=[hp["lo"], hp["hi"]],
                         charge_efficiency=hp["eta_c"], discharge_efficiency=hp["eta_d"],
                         initial_storage_cost=hp["init_cost"], rescale_spaces=bool(hp["rescale"][1]))
    by["ev-charging"].update(vehicle_multiplier=hp["mult"], max_charge_rate_kw=hp["rate"],
                             rescale_spaces=bool(hp["rescale"][2]))
    by["other-devices"].update(rescale_spaces=bool(hp["rescale"][3]))
    return cfg
