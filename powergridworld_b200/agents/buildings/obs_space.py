"""Observation layout of the building (gridworld/agents/buildings/obs_space.py:30-101)."""
from collections import OrderedDict

import numpy as np

from powergridworld_b200 import spaces

DEFAULT_OBS_CONFIG = OrderedDict({
    "zone_temp": (16., 40.), "zone_upper_viol": (-10., 10.), "zone_lower_viol": (-10., 10.),
    "comfort_lower": (20., 23.), "comfort_upper": (23., 26.), "outdoor_temp": (0., 56.),
    "p_setpoint": (0., 200.), "p_consumed": (0., 200.), "time_of_day": (0., 1.),
    "bus_voltage": (0.90, 1.10), "min_voltage": (0.90, 1.10), "max_voltage": (0.90, 1.10)})
MULTIZONE_KEYS = ["zone_temp", "zone_upper_viol", "zone_lower_viol"]

# Order in which FiveZoneROMEnv.get_obs gathers the *values* (five_zone_rom_env.py:256-269);
# it differs from the label/bounds order above for p_setpoint -- replicated, not fixed.
STATE_ORDER = ([f"zone_temp_{z}" for z in range(5)] + [f"zone_upper_viol_{z}" for z in range(5)]
               + [f"zone_lower_viol_{z}" for z in range(5)]
               + ["comfort_lower", "comfort_upper", "outdoor_temp", "p_consumed", "time_of_day",
                  "bus_voltage", "min_voltage", "max_voltage", "p_setpoint"])


def make_obs_space(num_zones, config):
    for key in config:
        assert key in DEFAULT_OBS_CONFIG, "invalid key {}".format(key)
    labels, low, high = [], [], []
    for key in [k for k in DEFAULT_OBS_CONFIG if k in config]:
        reps = num_zones if key in MULTIZONE_KEYS else 1
        for z in range(reps):
            labels.append(f"{key}_{z}" if key in MULTIZONE_KEYS else key)
            low.append(config[key][0])
            high.append(config[key][1])
    return spaces.Box(np.array(low, dtype=float), np.array(high, dtype=float),
                      dtype=np.float64), labels
