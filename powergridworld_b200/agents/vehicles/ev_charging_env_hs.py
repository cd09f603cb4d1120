"""Home-Steward EV charger description (gridworld/agents/vehicles/ev_charging_env_hs.py:15-336).
Dynamics: csrc/component_math.cuh hs_ev_advance / hs_ev_reset (the charging pass is the stock
station's ev_charge_pass; the window is evaluated at the fork's lagging clock, :260)."""
from collections import OrderedDict

import numpy as np

from powergridworld_b200 import _native as N
from powergridworld_b200 import assets, spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space

_COLS = ("start_time_min", "end_time_park_min", "energy_required_kwh")


class HSEVChargingEnv(ComponentEnv):

    def __init__(self, num_vehicles: int = 100, minutes_per_step: int = 5,
                 max_charge_rate_kw: float = 7.0, max_episode_steps: int = None,
                 unserved_penalty: float = 1., peak_penalty: float = 1.,
                 peak_threshold: float = 10., reward_scale: float = 1e5, name: str = None,
                 randomize: bool = False, vehicle_csv: str = None, vehicle_multiplier: int = 1,
                 rescale_spaces: bool = True, max_charge_cost: float = 0.55,
                 profile_data: dict = {}, **kwargs):
        super().__init__(name=name)
        self.num_vehicles, self.max_charge_rate_kw = num_vehicles, max_charge_rate_kw
        self.minutes_per_step, self.randomize = minutes_per_step, randomize
        self.vehicle_multiplier, self.rescale_spaces = vehicle_multiplier, rescale_spaces
        self.unserved_penalty, self.peak_penalty = unserved_penalty, peak_penalty
        self.peak_threshold, self.reward_scale = peak_threshold, reward_scale
        self.max_charge_cost = max_charge_cost
        mes = max_episode_steps if max_episode_steps is not None else np.inf
        self.max_episode_steps = min(mes, 24 * 60 / minutes_per_step)             # :54-55
        self.simulation_times = np.arange(
            0, (self.max_episode_steps + 1) * minutes_per_step, minutes_per_step)
        if profile_data:                            # pd.read_json(..., orient="split") (:68-69)
            names = profile_data["columns"]
            cols = {c: np.array([row[names.index(c)] for row in profile_data["data"]],
                                dtype=np.float64) for c in _COLS}
        elif vehicle_csv:
            import pandas as pd
            df = pd.read_csv(vehicle_csv)
            cols = {c: df[c].values.astype(np.float64) for c in _COLS}
        else:                                       # vehicles_hs.csv next to the module (:72-73)
            cols = {c: assets.array(f"hs/vehicles/{c}") for c in _COLS}
        # unlike the stock station the fork keeps EVERY row of the table (:142-143)
        self._roster_energy = cols["energy_required_kwh"] * self.vehicle_multiplier
        rnd = lambda x: x - x % self.minutes_per_step                             # :333-335
        self._roster_start = rnd(cols["start_time_min"])
        self._roster_end = rnd(cols["end_time_park_min"])
        emax = self._roster_energy.max()
        obs_bounds = OrderedDict({
            "time": (0, self.simulation_times[-1]),
            "num_active_vehicles": (0, self.num_vehicles),
            "real_power_consumed": (0, self.num_vehicles * self.max_charge_rate_kw),
            "real_power_demand": (0, self.num_vehicles * emax),
            "mean_charge_rate_deficit": (0, emax / (self.minutes_per_step / 60.)),
            "real_power_unserved": (0, emax),
            "current_cost": (0, max_charge_cost)})
        self._observation_space = spaces.Box(
            low=np.array([x[0] for x in obs_bounds.values()], dtype=np.float64),
            high=np.array([x[1] for x in obs_bounds.values()], dtype=np.float64),
            shape=(len(obs_bounds),), dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(low=0., high=1., shape=(1,), dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)
        self.state = OrderedDict({k: None for k in obs_bounds.keys()})
        self._obs_labels = list(self.state.keys())

    def _terminal_after(self):
        return int(self.max_episode_steps)          # time_index == max before the increment (:326)

    def _emit(self, b, agent_index, standalone):
        if standalone:
            raise NotImplementedError("HS components are stepped inside an HSMultiComponentEnv")
        n = len(self._roster_energy)
        start, end = np.floor(self._roster_start), np.floor(self._roster_end)
        times = self.simulation_times
        idx = np.arange(n)
        last = len(times) - 1

        def t_eval(r):
            # the hidden step of reset (event 0) and step 0 both see times[0]; step t sees
            # times[t]: the fork assigns self.time = simulation_times[time_index] BEFORE the
            # increment (:260, :329)
            return times[min(max(r - 1, 0), last)]

        def window(r):
            t = t_eval(r)
            return idx[(t >= start) & (t <= end)]          # :207-211, ascending index

        n_ev = len(times) + 1
        wins = [window(r) for r in range(n_ev)]
        lefts = [np.array([], dtype=int)] + [np.setdiff1d(wins[r - 1], wins[r]) for r in range(1, n_ev)]
        cap = max(1, max(len(w) for w in wins), max(len(l) for l in lefts))
        words = (n + 31) // 32
        if getattr(self, "_telemetry", False) and words > 4:
            raise NotImplementedError("step_meta telemetry supports up to 128 vehicles per charger")

        end_raw = self._roster_end

        def dtab_fn(r):
            r = min(r, n_ev - 1)
            left = np.zeros(cap)
            left[:len(wins[r])] = (end_raw[wins[r]] - t_eval(r)) / 60.            # :240
            with np.errstate(divide="ignore"):
                inv = np.where(left > 0, 1.0 / left, 0.0)
            return [t_eval(r), times[min(r, last)]] + list(left) + list(inv)

        def itab_fn(r):
            r = min(r, n_ev - 1)
            row = np.zeros(2 + 2 * cap, dtype=np.int32)
            row[0], row[1] = len(wins[r]), len(lefts[r])
            row[2:2 + len(wins[r])] = wins[r]
            row[2 + cap:2 + cap + len(lefts[r])] = lefts[r]
            return row

        hi = self._observation_space.high
        dpar = [self.max_charge_rate_kw, self.minutes_per_step / 60., float(self.vehicle_multiplier),
                self.unserved_penalty, self.peak_penalty, self.peak_threshold, self.reward_scale]
        with np.errstate(divide="ignore"):
            dpar += list(hi[:6]) + list(1.0 / hi[:6]) + [1.0 / self.reward_scale, 1.0 / 60.0]
        dpar += list(self._roster_end) + list(self._roster_energy)
        dpar += [self.max_charge_cost, 60.0 / self.minutes_per_step, 1.0 / self.max_charge_cost]
        b.add_component(self, N.HS_EV, agent_index,
                        flags=(N.F_RESCALE if self.rescale_spaces else 0)
                        | (N.F_TELEMETRY if getattr(self, "_telemetry", False) else 0),
                        dpar=dpar, ipar=[n, words, cap],
                        sd_rows=n + 1 + (N.HS_TEL_ROWS if getattr(self, "_telemetry", False) else 0),
                        si_rows=words,
                        dtab_width=2 + 2 * cap, dtab_fn=dtab_fn, itab_width=2 + 2 * cap, itab_fn=itab_fn)
