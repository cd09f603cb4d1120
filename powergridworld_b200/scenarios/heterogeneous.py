"""Mirror of gridworld/scenarios/heterogeneous.py:13-112: a composite building, a
grid-aware PV farm rewarded for voltage support, and an EV charging station."""
import pandas as pd

from powergridworld_b200 import MultiComponentEnv
from powergridworld_b200.agents.buildings import FiveZoneROMThermalEnergyEnv
from powergridworld_b200.agents.energy_storage import EnergyStorageEnv
from powergridworld_b200.agents.pv import GridAwarePVEnv, PVEnv
from powergridworld_b200.agents.vehicles import EVChargingEnv
from powergridworld_b200.distribution_system import OpenDSSSolver


class ThisPVEnv(GridAwarePVEnv):
    """gridworld/scenarios/heterogeneous.py:46-52: the PV farm rewarded for voltage support
    (the reward itself is evaluated on the device, see GridAwarePVEnv)."""


def make_env_config(system_load_rescale_factor=0.65, rescale_spaces=True):
    building_components = [
        {"name": "building", "cls": FiveZoneROMThermalEnergyEnv,
         "config": {"reward_structure": {"alpha": 0.0}, "rescale_spaces": rescale_spaces}},
        {"name": "pv", "cls": PVEnv,
         "config": {"profile_csv": "off-peak.csv", "scaling_factor": 40.,
                    "rescale_spaces": rescale_spaces}},
        {"name": "storage", "cls": EnergyStorageEnv,
         "config": {"max_power": 20., "storage_range": (3., 250.),
                    "rescale_spaces": rescale_spaces}},
    ]
    common_config = {"start_time": "08-12-2020 00:00:00", "end_time": "08-13-2020 00:00:00",
                     "control_timedelta": pd.Timedelta(300, "s")}
    pf_config = {"cls": OpenDSSSolver,
                 "config": {"feeder_file": "ieee_13_dss/IEEE13Nodeckt.dss",
                            "loadshape_file": "ieee_13_dss/annual_hourly_load_profile.csv",
                            "system_load_rescale_factor": system_load_rescale_factor}}
    agents = [
        {"name": "building", "bus": "675c", "cls": MultiComponentEnv,
         "config": {"components": building_components}},
        {"name": "pv", "bus": "675c", "cls": ThisPVEnv,
         "config": {"profile_csv": "constant.csv", "scaling_factor": 400.,
                    "rescale_spaces": rescale_spaces, "grid_aware": True}},
        {"name": "ev-charging", "bus": "675c", "cls": EVChargingEnv,
         "config": {"num_vehicles": 25, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
                    "peak_threshold": 200., "vehicle_multiplier": 40.,
                    "rescale_spaces": rescale_spaces}},
    ]
    return {"common_config": common_config, "pf_config": pf_config, "agents": agents}
