"""Home-Steward house (mirrors gridworld/base_hs.py:12-199): a MultiComponentEnv whose
components share the step's available solar / battery / grid power through a meta state and
pay a blended energy cost.  Like every class of this package it is a *description*: the
dynamics run on the GPU (csrc/component_math.cuh, hs_* functions), for one house through the
per-object protocol below or for ``num_envs`` houses as an agent of ``MultiAgentEnv``:

    cfg = make_env_config()                                   # scenarios/heterogeneous_hs.py
    house = HSMultiComponentEnv(**cfg); obs = house.reset(); obs, rew, done, meta = house.step(a)

    batch = MultiAgentEnv(common_config={...}, pf_config=None, num_envs=65536,
                          agents=[{"name": "house", "bus": None, "cls": HSMultiComponentEnv,
                                   "config": house_agent_config(cfg)}])
"""
from __future__ import annotations

from typing import List

import numpy as np
import pandas as pd

from powergridworld_b200 import _native as N
from powergridworld_b200.base import MultiComponentEnv

META_KEYS = ("pv_power", "es_power", "es_cost", "pv_cost", "grid_power")   # state-row order


class _HsBegin:
    """Descriptor object of the leading pseudo-component (no action, no observation)."""
    _act_dim = 0
    _obs_dim = 0


class HSMultiComponentEnv(MultiComponentEnv):

    def __init__(self, name: str = None, components: List[dict] = None, start_time: str = '',
                 end_time: str = '', control_timedelta=pd.Timedelta(300, "s"),
                 max_grid_power: float = 48, max_episode_steps: int = None,
                 rescale_spaces: bool = True, step_meta: bool = None, **kwargs):
        # step_meta: produce the per-device telemetry records of the reference's meta_state
        # (13 extra state rows per component).  None = only when the house is stepped on its
        # own through reset()/step(); batches leave it off unless asked.
        self._step_meta = step_meta
        self.max_grid_power = max_grid_power
        super().__init__(name=name, components=components)
        if len(self.envs) > N.HS_MAX_COMPONENTS:
            raise ValueError(f"a house holds at most {N.HS_MAX_COMPONENTS} components")
        for e in self.envs:
            if not hasattr(e, "_emit") or type(e).__name__[:2] != "HS":
                raise TypeError("HSMultiComponentEnv takes the HS* component classes")
        self.rescale_spaces = rescale_spaces
        self._grid_cost_data = list(kwargs["grid_cost"])                           # :42
        self._timestamps = list(kwargs["timestamps"])
        self.max_episode_steps = max_episode_steps if max_episode_steps is not None else np.inf
        self.meta_state = {"timestamp": None, "grid_cost": None, "es_cost": 0.0,
                           "grid_power": self.max_grid_power, "pv_power": None, "es_power": 0.0,
                           "pv_cost": 0.0, "step_meta": None}                      # :53-61
        self._begin = _HsBegin()
        self.time_index = 0

    # ---- spec compiler
    def _runner(self):
        if self._standalone is None and self._step_meta is None:
            self._step_meta = True
        return super()._runner()

    def _emit(self, b, agent_index, standalone):
        for e in self.envs:
            e._telemetry = bool(self._step_meta)
        cost, last = self._grid_cost_data, len(self._grid_cost_data) - 1
        b.add_component(self._begin, N.HS_BEGIN, agent_index, dpar=[self.max_grid_power],
                        sd_rows=5, dtab_width=1,
                        dtab_fn=lambda r: [cost[min(max(r - 1, 0), last)]])        # :124
        for e in self.envs:
            e._emit(b, agent_index, standalone=False)

    # ---- the reference's per-object protocol (one house, still the CUDA path)
    def _refresh_meta(self):
        r = self._runner()
        off, n = self._begin._slot["sd"]
        rows = r.get_field(N.FIELD_STATE_D)[off:off + n, 0].cpu().numpy()
        k = min(max(self.time_index - 1, 0), len(self._grid_cost_data) - 1)
        self.meta_state.update({key: float(v) for key, v in zip(META_KEYS, rows)})
        self.meta_state["grid_cost"] = self._grid_cost_data[k]
        self.meta_state["timestamp"] = self._timestamps[k] if k < len(self._timestamps) else None
        self.meta_state["step_meta"] = self._step_meta_records(r, self.meta_state["timestamp"]) \
            if self._step_meta else []

    _CUSTOM = {
        "HSPVEnv": (0, ("pv_available_power", "pv_actionable_power")),
        "HSEnergyStorageEnv": (2, ("current_storage", "power_ask", "solar_power_available",
                                   "grid_power_available", "es_power_available")),
        "HSEVChargingEnv": (None, ("power_ask", "power_unserved", "charging_vehicle",
                                   "vehicle_charged", "solar_power_available", "es_power_available",
                                   "grid_power_available")),
        "HSDevicesEnv": (0, ("power_ask", "solar_power_available", "es_power_available",
                             "grid_power_available"))}

    def telemetry_rows(self, comp) -> tuple:
        """(first row, count) of ``comp``'s telemetry block in the double state (FIELD_STATE_D)."""
        own, _ = self._CUSTOM[type(comp).__name__]
        if own is None:
            own = len(comp._roster_energy) + 1
        return comp._slot["sd"][0] + own, N.HS_TEL_ROWS

    def _step_meta_records(self, runner, timestamp, env_index: int = 0):
        """The reference's per-device records (base_hs.py:158-164) from the telemetry rows."""
        sd = runner.get_field(N.FIELD_STATE_D)[:, env_index].cpu().numpy()
        out = []
        for comp in self.envs:
            off, _ = self.telemetry_rows(comp)
            t = sd[off:off + N.HS_TEL_ROWS]
            names = self._CUSTOM[type(comp).__name__][1]
            custom = {k: float(t[6 + i]) for i, k in enumerate(names)}
            for k in ("charging_vehicle", "vehicle_charged"):
                if k in custom:
                    custom[k] = int(custom[k])
            out.append({"device_id": comp.name, "timestamp": timestamp, "cost": float(t[0]),
                        "reward": float(t[1]), "action": [float(t[2])],
                        "solar_power_consumed": float(t[3]), "es_power_consumed": float(t[4]),
                        "grid_power_consumed": float(t[5]), "device_custom_info": custom})
        return out

    def reset(self, **kwargs):
        obs, _ = super().reset(**kwargs)
        self.time_index = 0
        return obs                                  # base_hs.py:92

    def step(self, action: dict, **kwargs):
        obs, rew, done, _ = super().step(action, **kwargs)
        self.time_index += 1
        self._refresh_meta()
        return obs, rew, done, self.meta_state


def house_agent_config(env_config: dict) -> dict:
    """``env_config`` without the keys ``MultiAgentEnv`` passes through ``common_config``."""
    return {k: v for k, v in env_config.items()
            if k not in ("name", "start_time", "end_time", "control_timedelta")}
