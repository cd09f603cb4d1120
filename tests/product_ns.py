"""Product plugin namespace for tests/scenarios.py builders."""
import types

import powergridworld_b200 as pgw
from powergridworld_b200.agents.buildings import FiveZoneROMThermalEnergyEnv
from powergridworld_b200.agents.energy_storage import EnergyStorageEnv
from powergridworld_b200.agents.pv import GridAwarePVEnv, PVEnv
from powergridworld_b200.agents.vehicles import EVChargingEnv
from powergridworld_b200.distribution_system import OpenDSSSolver

PRODUCT_NS = types.SimpleNamespace(
    MultiComponentEnv=pgw.MultiComponentEnv,
    FiveZoneROMThermalEnergyEnv=FiveZoneROMThermalEnergyEnv,
    PVEnv=PVEnv, GridAwarePVEnv=GridAwarePVEnv, EnergyStorageEnv=EnergyStorageEnv,
    EVChargingEnv=EVChargingEnv, MultiAgentEnv=pgw.MultiAgentEnv,
    CoordinatedMultiBuildingControlEnv=pgw.CoordinatedMultiBuildingControlEnv,
    OpenDSSSolver=OpenDSSSolver)
