"""Product plugin namespace for tests/scenarios_hs.py builders."""
import types

from powergridworld_b200.agents.devices import HSDevicesEnv
from powergridworld_b200.agents.energy_storage import HSEnergyStorageEnv
from powergridworld_b200.agents.pv import HSPVEnv
from powergridworld_b200.agents.vehicles import HSEVChargingEnv
from powergridworld_b200.base_hs import HSMultiComponentEnv

PRODUCT_HS_NS = types.SimpleNamespace(
    HSPVEnv=HSPVEnv, HSEnergyStorageEnv=HSEnergyStorageEnv, HSEVChargingEnv=HSEVChargingEnv,
    HSDevicesEnv=HSDevicesEnv, HSMultiComponentEnv=HSMultiComponentEnv)
