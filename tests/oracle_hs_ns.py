"""Oracle plugin namespace for tests/scenarios_hs.py builders."""
import types

import oracle.components_hs as oh

ORACLE_HS_NS = types.SimpleNamespace(
    HSPVEnv=oh.HSPVEnv, HSEnergyStorageEnv=oh.HSEnergyStorageEnv,
    HSEVChargingEnv=oh.HSEVChargingEnv, HSDevicesEnv=oh.HSDevicesEnv,
    HSMultiComponentEnv=oh.HSMultiComponentEnv)
