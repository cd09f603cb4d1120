"""Plugin namespaces of the product classes for the scenario catalog's builders (the same
builders take the reference's or the oracle's classes in the tests)."""
import types

import powergridworld_b200 as pgw
from powergridworld_b200.agents.buildings import FiveZoneROMThermalEnergyEnv
from powergridworld_b200.agents.devices import HSDevicesEnv
from powergridworld_b200.agents.energy_storage import EnergyStorageEnv, HSEnergyStorageEnv
from powergridworld_b200.agents.pv import GridAwarePVEnv, HSPVEnv, PVEnv
from powergridworld_b200.agents.vehicles import EVChargingEnv, HSEVChargingEnv
from powergridworld_b200.base_hs import HSMultiComponentEnv
from powergridworld_b200.distribution_system import OpenDSSSolver

PRODUCT_NS = types.SimpleNamespace(
    MultiComponentEnv=pgw.MultiComponentEnv,
    FiveZoneROMThermalEnergyEnv=FiveZoneROMThermalEnergyEnv,
    PVEnv=PVEnv, GridAwarePVEnv=GridAwarePVEnv, EnergyStorageEnv=EnergyStorageEnv,
    EVChargingEnv=EVChargingEnv, MultiAgentEnv=pgw.MultiAgentEnv,
    CoordinatedMultiBuildingControlEnv=pgw.CoordinatedMultiBuildingControlEnv,
    OpenDSSSolver=OpenDSSSolver)

PRODUCT_HS_NS = types.SimpleNamespace(
    HSPVEnv=HSPVEnv, HSEnergyStorageEnv=HSEnergyStorageEnv, HSEVChargingEnv=HSEVChargingEnv,
    HSDevicesEnv=HSDevicesEnv, HSMultiComponentEnv=HSMultiComponentEnv)
