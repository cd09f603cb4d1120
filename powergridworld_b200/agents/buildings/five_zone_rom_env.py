"""Five-zone reduced-order building description
(gridworld/agents/buildings/five_zone_rom_env.py:60-335 + five_zone_rom_dynamics.py).
Dynamics: csrc/component_math.cuh building_step / building_reset."""
from typing import Union

import numpy as np
import pandas as pd

from powergridworld_b200 import _native as N
from powergridworld_b200 import assets, spaces
from powergridworld_b200.agents.buildings import defaults
from powergridworld_b200.agents.buildings.exogenous import load_exogenous
from powergridworld_b200.agents.buildings.obs_space import STATE_ORDER, make_obs_space
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space

MAX_FLOW_RATE = [2.2, 2.2, 2.2, 2.2, 3.2]
MIN_FLOW_RATE = [.22, .22, .22, .22, .32]
MAX_DISCHARGE_TEMP = 16.0
MIN_DISCHARGE_TEMP = 10.0
DEFAULT_COMFORT_BOUNDS = (22., 28.)


class FiveZoneROMEnv(ComponentEnv):
    _thermal_energy_reward = False

    def __init__(self, name: str = None, obs_config: dict = None,
                 start_time: Union[str, pd.Timestamp] = None,
                 end_time: Union[str, pd.Timestamp] = None,
                 comfort_bounds: Union[tuple, np.ndarray] = None,
                 zone_temp_init: np.ndarray = None, max_episode_steps: int = None,
                 rescale_spaces: bool = True, exogenous_csv: str = None, **kwargs):
        super().__init__(name=name)
        self.rescale_spaces = rescale_spaces
        self.num_zones = 5
        self.obs_config = obs_config if obs_config is not None else defaults.obs_config
        self.zone_temp_init = np.array(zone_temp_init, dtype=np.float64) \
            if zone_temp_init is not None else 27. * np.ones(5, dtype=np.float64)
        self.exo, self.exo_index = load_exogenous(start_time, end_time, exogenous_csv)
        max_steps = self.exo.shape[0] - 3                                   # :97
        self.max_episode_steps = max_steps if max_episode_steps is None \
            else min(max_episode_steps, max_steps)
        self.comfort_bounds = comfort_bounds if comfort_bounds is not None \
            else DEFAULT_COMFORT_BOUNDS
        rows = self.exo.shape[0]
        if isinstance(self.comfort_bounds, tuple):
            self._comfort = np.tile(np.asarray(self.comfort_bounds, dtype=np.float64), (rows, 1))
        else:
            self._comfort = np.asarray(self.comfort_bounds, dtype=np.float64)[:rows, :2]
        self.act_low = np.array(MIN_FLOW_RATE + [MIN_DISCHARGE_TEMP])
        self.act_high = np.array(MAX_FLOW_RATE + [MAX_DISCHARGE_TEMP])
        self._action_space = spaces.Box(low=self.act_low, high=self.act_high, dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)
        self._observation_space, self._obs_labels = make_obs_space(5, self.obs_config)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)

    def _terminal_after(self):
        return self.max_episode_steps - 1           # time_index == max_episode_steps - 1 (:301)

    def _meta(self, ctx) -> dict:
        """The state dict get_obs builds and step returns as meta (:223, :256-271): zone
        temperatures C x + mean from the device state, comfort margins against the event's
        bounds, exogenous entries of the event row, the lagged grid variables the step saw."""
        from collections import OrderedDict
        Cm, mean = assets.array("building/ss_C"), assets.array("building/mean_output")
        s0, d0 = self._slot["sd"][0], self._slot["dtab"][0]
        T = [Cm[z] * ctx.sd(s0 + z) + mean[z] for z in range(5)]
        lb, ub = ctx.dtab(d0 + 12), ctx.dtab(d0 + 13)
        st = OrderedDict()
        for z in range(5):
            st[f"zone_temp_{z}"] = T[z]
        for z in range(5):
            st[f"zone_upper_viol_{z}"] = T[z] - ub
        for z in range(5):
            st[f"zone_lower_viol_{z}"] = lb - T[z]
        st["comfort_lower"], st["comfort_upper"] = ctx.scalar(lb), ctx.scalar(ub)
        st["outdoor_temp"] = ctx.scalar(ctx.dtab(d0 + 11))
        st["p_consumed"] = ctx.sd(s0 + 5)
        st["time_of_day"] = ctx.scalar(ctx.dtab(d0 + 14))
        # nominal 1.0 unless delivered; min / max default to the BUS voltage (sic, :265-267),
        # then the delivered grid variables overwrite their keys (:271)
        bv = ctx.grid("bus_voltage") if "bus_voltage" in self._obs_labels else None
        for key in ("bus_voltage", "min_voltage", "max_voltage"):
            st[key] = bv if bv is not None else 1.0
        st["p_setpoint"] = np.inf
        for key in ("min_voltage", "max_voltage"):
            if key in self._obs_labels:
                st[key] = ctx.grid(key)
        return st

    def _emit(self, b, agent_index, standalone):
        if not self._thermal_energy_reward:
            raise NotImplementedError(
                "FiveZoneROMEnv.step_reward returns a 5-vector (five_zone_rom_env.py:286-294) "
                "that no shipped scenario uses; use FiveZoneROMThermalEnergyEnv")
        A = assets.array("building/ss_A")
        B32 = assets.array("building/ss_B").astype(np.float32).astype(np.float64)   # dynamics.py:51
        Cm = assets.array("building/ss_C")
        K = assets.array("building/ss_K")
        mean = assets.array("building/mean_output")
        sel = assets.array("building/input_sel_list") - 1
        nbr = assets.array("building/neighbors")
        alpha = 0.2                                                          # :318
        low, high = self._observation_space.low, self._observation_space.high
        # observation slots: selected sources in state-dict order (:256-276)
        obs_src = [src for src, key in enumerate(STATE_ORDER) if key in self._obs_labels]
        assert len(obs_src) == self._obs_dim
        # model inputs: input_sel_list picks 4 of the 8 candidates of build_u_vector
        # (0 outdoor, 1 solar, 2 internal, 3-6 neighbour i, 7 cooling), dynamics.py:12-41
        u_kind, u_arg = [], []
        for z in range(5):
            for j in range(4):
                cand = int(sel[z, j])
                if cand <= 2:
                    u_kind.append(cand); u_arg.append(0)
                elif cand <= 6:
                    u_kind.append(3); u_arg.append(int(nbr[z, cand - 3]))
                else:
                    u_kind.append(4); u_arg.append(0)
        dpar = list(A) + list(B32.reshape(-1)) + list(Cm) + list(K) + list(mean) + \
            list(self.zone_temp_init) + [alpha * 0.5, 1. - alpha] + list(low) + list(high) + \
            list(1.0 / (high - low))
        ipar = u_kind + u_arg + obs_src
        exo, comfort, mes = self.exo, self._comfort, self.max_episode_steps

        def dtab_fn(r):
            if r == 0:       # reset: row 0, Q_cool drives the filter (:160-168)
                return [exo[0, 0], *exo[0, 1:6], *exo[0, 6:11], exo[0, 0],
                        comfort[0, 0], comfort[0, 1], 0.0, comfort[0, 0], comfort[0, 1]]
            t = r - 1        # dynamics on row t, observation of row t+1 (:203-223)
            return [exo[t, 0], *exo[t, 1:6], *exo[t, 11:16], exo[t + 1, 0],
                    comfort[t + 1, 0], comfort[t + 1, 1], 1. * (t + 1) / mes,
                    comfort[t, 0], comfort[t, 1]]

        grid = any(k in self._obs_labels for k in ("bus_voltage", "min_voltage", "max_voltage"))
        flags = (N.F_RESCALE if self.rescale_spaces else 0) | \
                (N.F_STALE_REWARD if standalone else 0) | (N.F_GRID_AWARE if grid else 0)
        # straight-line kernel path for the shipped model + default observation set
        default_src = list(range(5, 15)) + [15, 16, 17, 18, 19]
        if u_kind == [0, 4, 3, 1] * 5 and obs_src == default_src and not standalone:
            flags |= N.F_BUILDING_FAST
        b.add_component(self, N.BUILDING, agent_index, flags=flags, dpar=dpar, ipar=ipar,
                        sd_rows=6, dtab_width=17, dtab_fn=dtab_fn, needs_grid=grid,
                        max_events=self.exo.shape[0] - 1)


class FiveZoneROMThermalEnergyEnv(FiveZoneROMEnv):
    """Same physics, reward balancing energy and comfort (:312-335)."""
    _thermal_energy_reward = True
