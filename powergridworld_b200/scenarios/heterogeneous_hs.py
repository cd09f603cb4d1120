"""gridworld/scenarios/heterogeneous_hs.py:47-58 on the packaged copy of the scenario data
(powergridworld_b200/data/hs_env_config.json, numbers only)."""
import sys

import pandas as pd

from powergridworld_b200.agents.devices import HSDevicesEnv  # noqa: F401
from powergridworld_b200.agents.energy_storage import HSEnergyStorageEnv  # noqa: F401
from powergridworld_b200.agents.pv import HSPVEnv  # noqa: F401
from powergridworld_b200.agents.pv.pv_profile_env_hs import packaged_hs_config
from powergridworld_b200.agents.vehicles import HSEVChargingEnv  # noqa: F401


def make_env_config():
    env_config = packaged_hs_config()
    for elem in env_config["components"]:
        elem["cls"] = getattr(sys.modules[__name__], elem["cls"])
    env_config["control_timedelta"] = pd.Timedelta(env_config["control_timedelta"])
    return env_config
