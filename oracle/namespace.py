"""Oracle plugin namespaces for the scenario catalog's builders (test infrastructure: the same
builders that configure the product build the CPU oracle from these classes)."""
import types

import oracle.components as oc
import oracle.components_hs as oh
import oracle.multiagent as om
from oracle.powerflow import OracleOpenDSSSolver

ORACLE_NS = types.SimpleNamespace(
    MultiComponentEnv=oc.MultiComponentEnv,
    FiveZoneROMThermalEnergyEnv=oc.FiveZoneROMThermalEnergyEnv,
    PVEnv=oc.PVEnv, GridAwarePVEnv=oc.GridAwarePVEnv, EnergyStorageEnv=oc.EnergyStorageEnv,
    EVChargingEnv=oc.EVChargingEnv, MultiAgentEnv=om.MultiAgentEnv,
    CoordinatedMultiBuildingControlEnv=om.CoordinatedMultiBuildingControlEnv,
    OpenDSSSolver=OracleOpenDSSSolver)


def storage_socs_to_dict(env, socs):
    """Flat SOC vector (agent order, component order) -> oracle ``init_storage`` dict."""
    out, k = {}, 0
    for a in env.agents:
        comps = getattr(a, "envs", None)
        for e in (comps if comps is not None else [a]):
            if isinstance(e, oc.EnergyStorageEnv):
                if comps is None:
                    out[a.name] = socs[k]
                else:
                    out.setdefault(a.name, {})[e.name] = socs[k]
                k += 1
    assert k == len(socs)
    return out


ORACLE_HS_NS = types.SimpleNamespace(
    HSPVEnv=oh.HSPVEnv, HSEnergyStorageEnv=oh.HSEnergyStorageEnv,
    HSEVChargingEnv=oh.HSEVChargingEnv, HSDevicesEnv=oh.HSDevicesEnv,
    HSMultiComponentEnv=oh.HSMultiComponentEnv)
