"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import harness for the *unmodified* reference (``/root/reference/gridworld``).

The reference cannot be imported as-is in this image: ``gym`` is absent,
pandas 3 makes ``DataFrame.values`` read-only, and the building env's
``data/exogenous_data.csv`` is missing from the mount.
This module installs three shims and then imports the reference classes, so that
``tests/golden/make_golden.py`` can record golden traces from the real reference
code.  The reference tree does not exist on the GPU box; nothing that runs
there may import this module (the recorded ``.npz`` fixtures travel instead).

Shims (all behaviour-neutral for the hot path):
  1. ``gym`` stub: ``gym.Env``, ``gym.spaces.{Box,Dict,Discrete}`` -- only the
     attributes the reference touches (low/high/shape/dtype/sample/items).
  2. ``pd.read_csv(...).values`` must be writable: ``PVEnv.__init__``
     multiplies the profile in place (gridworld/agents/pv/pv_profile_env.py:68-69).
  3. ``five_zone_rom_env.load_data`` (gridworld/agents/buildings/five_zone_rom_env.py:30-52)
     is replaced by a loader that serves the deterministic synthetic exogenous
     table of ``oracle.exogenous`` (same slicing semantics) plus the real
     ``state_space_model.p``.
"""
from __future__ import annotations

import copy
import os
import pickle
import sys
import types

import numpy as np

def _find_reference_root() -> str:
    """$PGW_REFERENCE_ROOT, the authoring container's mount, or the git-ignored install that
    oracle/install_reference.py leaves in baseline/_ref (the only one present on a GPU box)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("PGW_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "gridworld")):
            return cand
    return os.environ.get("PGW_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "gridworld"))


# --------------------------------------------------------------------------- gym stub
def _install_gym_stub() -> None:
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Env:  # noqa: D401 - minimal stand-in for gym.Env
        metadata: dict = {}

        def __init__(self, *a, **k):
            pass

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float64):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

    class Discrete(Space):
        def __init__(self, n):
            self.n = n
            self.shape = ()

        def sample(self):
            return np.random.randint(self.n)

    class Dict(Space):
        def __init__(self, spaces_dict):
            self.spaces = dict(spaces_dict)

        def __getitem__(self, k):
            return self.spaces[k]

        def __iter__(self):
            return iter(self.spaces)

        def items(self):
            return self.spaces.items()

        def keys(self):
            return self.spaces.keys()

        def values(self):
            return self.spaces.values()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

    spaces.Space, spaces.Box, spaces.Discrete, spaces.Dict = Space, Box, Discrete, Dict
    gym.Env, gym.spaces = Env, spaces
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


# --------------------------------------------------------------------------- pandas shim
def _install_pandas_shim() -> None:
    import pandas as pd

    if getattr(pd.DataFrame, "_pgw_values_patched", False):
        return
    orig = pd.DataFrame.values

    def writable_values(self):
        v = orig.fget(self)
        if not v.flags.writeable:
            v = v.copy()
        return v

    pd.DataFrame.values = property(writable_values)
    pd.DataFrame._pgw_values_patched = True


# --------------------------------------------------------------------------- building data shim
def _patched_load_data(start_time=None, end_time=None):
    """Same contract as five_zone_rom_env.load_data (:30-52) on synthetic data."""
    import pandas as pd

    from oracle.exogenous import synthetic_exogenous_frame

    df = synthetic_exogenous_frame()
    start_time = pd.Timestamp(start_time) if start_time else df.index[0]
    end_time = pd.Timestamp(end_time) if end_time else df.index[-1]
    _df = df.loc[start_time:end_time]
    if _df is None or len(_df) == 0:
        raise ValueError("empty exogenous slice")
    path = os.path.join(
        REFERENCE_ROOT, "gridworld/agents/buildings/data/state_space_model.p")
    with open(path, "rb") as f:
        models = pickle.load(f)
    return _df, models


_REF = None


def load_reference():
    """Import the reference package through the shims; returns a namespace."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_gym_stub()
    _install_pandas_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)   # appended: our own `tests` package must win
    import logging

    import gridworld  # noqa: F401  (the reference)
    from gridworld.agents.buildings import five_zone_rom_env
    from gridworld.log import logger

    logger.setLevel(logging.ERROR)
    five_zone_rom_env.load_data = _patched_load_data

    ns = types.SimpleNamespace()
    from gridworld import MultiAgentEnv, MultiComponentEnv
    from gridworld.agents.buildings import FiveZoneROMEnv, FiveZoneROMThermalEnergyEnv
    from gridworld.agents.energy_storage import EnergyStorageEnv
    from gridworld.agents.pv import PVEnv
    from gridworld.agents.vehicles import EVChargingEnv
    from gridworld.distribution_system.powerflow import PowerFlowSolver

    ns.MultiAgentEnv = MultiAgentEnv
    ns.MultiComponentEnv = MultiComponentEnv
    ns.FiveZoneROMEnv = FiveZoneROMEnv
    ns.FiveZoneROMThermalEnergyEnv = FiveZoneROMThermalEnergyEnv
    ns.EnergyStorageEnv = EnergyStorageEnv
    ns.PVEnv = PVEnv
    ns.EVChargingEnv = EVChargingEnv
    ns.PowerFlowSolver = PowerFlowSolver
    ns.root = REFERENCE_ROOT
    _REF = ns
    return ns


def reference_namespace(ref=None):
    """The reference's classes as a plugin namespace for the scenario catalog, plus the two
    subclasses its own example scripts define outside the package."""
    ref = ref or load_reference()

    class ThisPVEnv(ref.PVEnv):                       # scenarios/heterogeneous.py:46-52
        def step_reward(self, **kwargs):
            v = kwargs["min_voltage"]
            viol = min(0, v - 0.95) + min(0, 1.05 - v)
            return -(1000 * viol) ** 2, {}

    class Coordinated(ref.MultiAgentEnv):             # examples/marl/openai/train.py:37-88
        VOLTAGE_LIMITS = [0.95, 1.05]
        VV_UNIT_PENALTY = 1e4

        def reward_transform(self, rew_dict):
            pen = self.get_voltage_violation() * self.VV_UNIT_PENALTY
            n = len(rew_dict)
            for k in rew_dict.keys():
                rew_dict[k] -= (pen / n)
            return rew_dict

        def get_voltage_violation(self):
            bus_id = list(set(self.agent_name_bus_map.values()))[0]
            v = self.pf_solver.get_bus_voltage_by_name(bus_id)
            return max([0.0, self.VOLTAGE_LIMITS[0] - v, v - self.VOLTAGE_LIMITS[1]])

    return types.SimpleNamespace(
        MultiComponentEnv=ref.MultiComponentEnv,
        FiveZoneROMThermalEnergyEnv=ref.FiveZoneROMThermalEnergyEnv,
        PVEnv=ref.PVEnv, GridAwarePVEnv=ThisPVEnv, EnergyStorageEnv=ref.EnergyStorageEnv,
        EVChargingEnv=ref.EVChargingEnv, MultiAgentEnv=ref.MultiAgentEnv,
        CoordinatedMultiBuildingControlEnv=Coordinated)


class quiet_stdout:
    """EnergyStorageEnv.get_obs prints every call (energy_storage_env.py:172)."""

    def __enter__(self):
        self._saved = sys.stdout
        sys.stdout = open(os.devnull, "w")
        return self

    def __exit__(self, *exc):
        sys.stdout.close()
        sys.stdout = self._saved
        return False
