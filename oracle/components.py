"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's agent models.

Plain-Python/NumPy float64 restatement of the device models and of the agent
composition on PowerGridworld's ``MultiAgentEnv.step`` hot path.  One object =
one env instance, stepped scalar-wise, exactly like the reference.  Every class
keeps the reference's constructor keywords so the same scenario builder can be
instantiated against the reference classes, these oracle classes and the
product classes (``tests/scenarios.py``).

Pinned against the reference itself: ``tests/golden/make_golden.py`` imports the
unmodified reference (through ``oracle/ref_harness.py``) and records traces;
``tests/test_oracle_golden.py`` replays them through this file.  The three EV
notebook totals of ``examples/envs/ev-charging.ipynb`` cells 5-7 are pinned
bit-exactly.

Reference files restated here (paths relative to the reference root):
  gridworld/utils.py:9-43                               to_scaled / to_raw
  gridworld/agents/energy_storage/energy_storage_env.py EnergyStorageEnv
  gridworld/agents/pv/pv_profile_env.py                 PVEnv
  gridworld/agents/vehicles/ev_charging_env.py          EVChargingEnv
  gridworld/agents/buildings/five_zone_rom_env.py       FiveZoneROM*Env
  gridworld/agents/buildings/five_zone_rom_dynamics.py  zone dynamics
  gridworld/agents/buildings/obs_space.py, defaults.py  observation layout
  gridworld/base.py:74-182                              MultiComponentEnv
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

_ASSETS = None
_ASSET_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                           "powergridworld_b200", "data", "assets.npz")


def assets():
    """The binary bundle of the reference's input data (tools/import_reference_data.py)."""
    global _ASSETS
    if _ASSETS is None:
        with np.load(_ASSET_PATH) as z:
            _ASSETS = {k: z[k] for k in z.files}
    return _ASSETS


class Box:
    """Bounds holder standing in for gym.spaces.Box (low/high/shape only)."""

    def __init__(self, low, high, shape=None):
        low = np.asarray(low, dtype=np.float64)
        high = np.asarray(high, dtype=np.float64)
        if shape is None:
            shape = np.broadcast(low, high).shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(low, self.shape).copy()
        self.high = np.broadcast_to(high, self.shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high)


def _unit_box(box, rescale):
    # utils.py:46-53 maybe_rescale_box_space
    return Box(-1.0, 1.0, box.shape) if rescale else box


def to_scaled(x, low, high):
    """utils.py:9-24: clip to [low, high] then map affinely onto [-1, 1]."""
    x = np.clip(x, low, high)
    return (2 * x - (low + high)) / (high - low)


def to_raw(y, low, high):
    """utils.py:27-43: clip to [-1, 1] then map affinely onto [low, high]."""
    y = np.clip(y, -np.ones_like(y), np.ones_like(y))
    return (y * (high - low) + (high + low)) / 2.0


class ComponentEnv:
    """base.py:12-71 -- the agent plugin protocol."""

    def __init__(self, name=None, **kwargs):
        self.name = name
        self._real_power = 0.0
        self._reactive_power = 0.0   # base.py:24, never reassigned by any shipped agent
        self._obs_labels = []

    @property
    def real_power(self):
        return self._real_power

    @property
    def reactive_power(self):
        return self._reactive_power

    @property
    def obs_labels(self):
        return self._obs_labels


# ======================================================================== storage
class EnergyStorageEnv(ComponentEnv):
    """energy_storage_env.py:20-181."""

    def __init__(self, name=None, storage_range=(3.0, 50.0), initial_storage_mean=30.0,
                 initial_storage_std=5.0, charge_efficiency=0.95, discharge_efficiency=0.9,
                 max_power=15.0, max_episode_steps=288, control_timedelta=None,
                 rescale_spaces=True, **kwargs):
        super().__init__(name=name)
        self.lo, self.hi = float(storage_range[0]), float(storage_range[1])
        self.initial_storage_mean = initial_storage_mean
        self.initial_storage_std = initial_storage_std
        self.eta_c = charge_efficiency
        self.eta_d = discharge_efficiency
        self.max_power = max_power
        self.rescale_spaces = rescale_spaces
        self.max_episode_steps = max_episode_steps
        seconds = 300 if control_timedelta is None else control_timedelta.seconds
        self.dt_h = seconds / 3600.0                       # :49
        self.soc = None
        self.simulation_step = 0
        self._obs_labels = ["stage_of_charge"]             # sic, :51
        self._observation_space = Box(self.lo, self.hi, (1,))
        self._action_space = Box(-1.0, 1.0, (1,))
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self.action_space = _unit_box(self._action_space, rescale_spaces)

    def reset(self, **kwargs):
        # :72-97
        self.simulation_step = 0
        init = kwargs.get("init_storage")
        if init is None:
            from scipy.stats import truncnorm
            self.soc = float(truncnorm(-1, 1).rvs() * self.initial_storage_std
                             + self.initial_storage_mean)
        else:
            self.soc = float(np.clip(float(init), self.lo, self.hi))
        return self.get_obs(**kwargs)

    def feasible_power(self, p):
        # validate_power :100-128.  Note the clamp formulas omit the efficiencies.
        if p > 0:
            if self.soc - p * self.dt_h / self.eta_d < self.lo:
                p = max(self.soc - self.lo, 0.0) / self.dt_h
        elif p < 0:
            if self.soc - self.eta_c * p * self.dt_h > self.hi:
                p = -max(self.hi - self.soc, 0.0) / self.dt_h
        return p

    def step(self, action, **kwargs):
        # :131-157
        action = np.asarray(action, dtype=np.float64)
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        p = self.feasible_power(action[0] * self.max_power)
        if p < 0.0:
            self.soc -= self.eta_c * p * self.dt_h
            self.soc = min(self.soc, self.hi)
        elif p > 0.0:
            self.soc -= p * self.dt_h / self.eta_d
            self.soc = max(self.soc, self.lo)
        self._real_power = -p
        obs, meta = self.get_obs()
        self.simulation_step += 1
        return obs, 0.0, self.is_terminal(), meta

    def step_reward(self, **kwargs):
        return 0.0, {}

    def get_obs(self, **kwargs):
        raw = np.array([self.soc])
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        return obs, {"state_of_charge": raw}

    def is_terminal(self):
        return self.simulation_step + 1 == self.max_episode_steps   # :181


# ======================================================================== PV
def load_profile(profile_csv, profile_path=None):
    """First CSV column, first line consumed as header (pv_profile_env.py:62-68)."""
    if profile_path is not None:
        rows = []
        with open(profile_path) as fh:
            next(fh)
            for line in fh:
                line = line.strip()
                if line:
                    rows.append(float(line.split(",")[0]))
        return np.array(rows, dtype=np.float64)
    return assets()[f"pv/{profile_csv}"].copy()


class PVEnv(ComponentEnv):
    """pv_profile_env.py:22-148."""

    def __init__(self, name=None, profile_csv=None, profile_path=None, scaling_factor=1.0,
                 rescale_spaces=True, grid_aware=False, max_episode_steps=None, **kwargs):
        super().__init__(name=name)
        self.rescale_spaces = rescale_spaces
        self.grid_aware = grid_aware
        self.data = load_profile(profile_csv, profile_path) * scaling_factor
        self.episode_length = len(self.data)
        if max_episode_steps is not None:
            self.episode_length = min(max_episode_steps, self.episode_length)
        self._obs_labels = ["real_power"] + (["min_voltage"] if grid_aware else [])
        low = [-np.max(self.data)] + ([0.9] if grid_aware else [])
        high = [0.0] + ([1.1] if grid_aware else [])
        self._observation_space = Box(np.array(low), np.array(high))
        self._action_space = Box(0.0, 1.0, (1,))
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self.index = None

    def get_obs(self, **kwargs):
        raw = [-self.data[self.index]]
        if self.grid_aware:
            raw.append(kwargs["min_voltage"])
        raw = np.array(raw)
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        return obs, {"real_power": raw[0]}

    def is_terminal(self):
        return self.index == self.episode_length - 1

    def step_reward(self, **kwargs):
        return 0.0, {}

    def reset(self, **kwargs):
        self.index = 0
        self.get_obs(**kwargs)      # returns None like the reference (:127-130)

    def step(self, action, **kwargs):
        # :133-148 -- the obs is built BEFORE the index advances.
        action = np.asarray(action, dtype=np.float64)
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        obs, meta = self.get_obs(**kwargs)
        self._real_power = np.float64((action * meta["real_power"]).squeeze())
        self.index += 1
        rew, _ = self.step_reward(**kwargs)
        return obs, rew, self.is_terminal(), meta


class GridAwarePVEnv(PVEnv):
    """``ThisPVEnv`` of gridworld/scenarios/heterogeneous.py:46-52 (lagged min voltage)."""

    def step_reward(self, **kwargs):
        v = kwargs["min_voltage"]
        viol = min(0, v - 0.95) + min(0, 1.05 - v)
        return -(1000 * viol) ** 2, {}


# ======================================================================== EV
class EVChargingEnv(ComponentEnv):
    """ev_charging_env.py:19-275."""

    def __init__(self, num_vehicles=100, minutes_per_step=5, max_charge_rate_kw=7.0,
                 max_episode_steps=None, unserved_penalty=1.0, peak_penalty=1.0,
                 peak_threshold=10.0, reward_scale=1e5, name=None, randomize=False,
                 vehicle_csv=None, vehicle_multiplier=1, rescale_spaces=True, **kwargs):
        super().__init__(name=name)
        self.randomize = randomize
        self.n = num_vehicles
        self.rate = max_charge_rate_kw
        self.minutes_per_step = minutes_per_step
        self.mult = vehicle_multiplier
        self.rescale_spaces = rescale_spaces
        self.unserved_penalty = unserved_penalty
        self.peak_penalty = peak_penalty
        self.peak_threshold = peak_threshold
        self.reward_scale = reward_scale
        mes = max_episode_steps if max_episode_steps is not None else np.inf
        self.max_episode_steps = min(mes, 24 * 60 / minutes_per_step)        # :54-55
        self.simulation_times = np.arange(
            0, self.max_episode_steps * minutes_per_step, minutes_per_step)  # :59-60
        a = assets()
        e_all = a["vehicles/energy_required_kwh"] * self.mult                # :72
        rnd = lambda x: x - x % minutes_per_step                             # :273-275
        self._start = rnd(a["vehicles/start_time_min"])
        self._end = rnd(a["vehicles/end_time_park_min"])
        self._e0 = e_all
        emax = e_all.max()
        low = np.zeros(6)
        high = np.array([self.simulation_times[-1], self.n, self.n * self.rate,
                         self.n * emax, emax / (minutes_per_step / 60.0), emax],
                        dtype=np.float64)                                    # :79-91
        self._observation_space = Box(low, high)
        self._action_space = Box(0.0, 1.0, (1,))
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self.state = OrderedDict((k, None) for k in [
            "time", "num_active_vehicles", "real_power_consumed", "real_power_demand",
            "mean_charge_rate_deficit", "real_power_unserved"])
        self._obs_labels = list(self.state.keys())
        self.time_index = None
        self.time = None
        self.energy = None
        self.charging_vehicles = None
        self.departed_vehicles = None

    def get_obs(self, **kwargs):
        raw = np.array(list(self.state.values()), dtype=np.float64)
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        return obs.copy(), self.state.copy()

    def is_terminal(self):
        return self.time_index == self.max_episode_steps - 1

    def step_reward(self, **kwargs):
        # :135-142
        unserved = -self.unserved_penalty * self.state["real_power_unserved"] ** 2
        peak = -self.peak_penalty * max(0, self.state["real_power_consumed"] - self.peak_threshold) ** 2
        return (unserved + peak) / self.reward_scale, \
            {"real_power_unserved": unserved, "peak_reward": peak}

    def reset(self, **kwargs):
        # :145-168 -- includes one hidden step with the minimum action.
        self.time_index = 0
        self.time = self.simulation_times[0]
        self.charging_vehicles = []
        self.departed_vehicles = []
        # :154-157  df.sample(n) draws np.random.choice(len(df), n, replace=False) on NumPy's global
        # RNG (pandas' random_state=None) and keeps the rows in drawn order
        rows = np.random.choice(len(self._e0), size=self.n, replace=False) if self.randomize \
            else np.arange(self.n)
        self.energy = self._e0[rows].copy()
        self.start = self._start[rows]
        self.end = self._end[rows]
        self._real_power = 0.0
        self.step(**kwargs)
        obs, _ = self.get_obs()
        return obs, {}

    def step(self, action=None, **kwargs):
        # :171-264
        action = np.asarray(action, dtype=np.float64) if action is not None \
            else self._action_space.low
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        kwh = (action[0] * self.rate) * (self.minutes_per_step / 60.0)

        arrived = np.where(self.time >= np.floor(self.start))[0]
        parked = np.where(self.time <= np.floor(self.end))[0]
        # Python-set semantics kept on purpose: iteration order of the
        # intersection / difference is what the reference sums in (:190-194).
        charging = list(set(list(arrived)).intersection(set(list(parked))))
        charging = [i for i in charging if self.energy[i] > 0.0]
        self.departed_vehicles = list(set(self.charging_vehicles) - set(charging))

        consumed = 0.0
        demand = 0.0
        deficits = []
        for i in charging:
            need = self.energy[i]
            demand += need
            if need <= 0.0:
                continue
            left_h = (self.end[i] - self.time) / 60.0
            if left_h <= 0:
                continue
            deficits.append(max(0, self.rate - need / left_h))
            delta = min(kwh, need)
            self.energy[i] -= delta
            consumed += delta

        self.time_index += 1
        self.time = self.simulation_times[self.time_index]
        self.charging_vehicles = charging

        unserved = 0.0
        for i in self.departed_vehicles:
            unserved += self.energy[i]
        self.state["real_power_unserved"] = unserved
        self.state["time"] = self.time
        self.state["num_active_vehicles"] = self.mult * len(charging)
        self.state["real_power_consumed"] = self.mult * consumed
        self.state["real_power_demand"] = self.mult * demand
        self.state["mean_charge_rate_deficit"] = 0 if len(deficits) == 0 else np.mean(deficits)
        self._real_power = self.mult * consumed      # kWh per step reported as kW (:255)

        obs, meta = self.get_obs(**kwargs)
        rew, rew_meta = self.step_reward(**kwargs)
        meta.update(rew_meta)
        return obs, rew, self.is_terminal(), meta


# ======================================================================== building
FLOW_HI = [2.2, 2.2, 2.2, 2.2, 3.2]       # five_zone_rom_env.py:22-26
FLOW_LO = [0.22, 0.22, 0.22, 0.22, 0.32]
T_DIS_HI, T_DIS_LO = 16.0, 10.0
COMFORT = (22.0, 28.0)                    # :27

# obs_space.py:30-43 (order matters: it is the order of the *bounds*)
OBS_KEYS_BOUND_ORDER = ["zone_temp", "zone_upper_viol", "zone_lower_viol", "comfort_lower",
                        "comfort_upper", "outdoor_temp", "p_setpoint", "p_consumed",
                        "time_of_day", "bus_voltage", "min_voltage", "max_voltage"]
PER_ZONE_KEYS = ["zone_temp", "zone_upper_viol", "zone_lower_viol"]
# five_zone_rom_env.py:256-269 (order of the *values*; differs for p_setpoint)
STATE_KEYS_VALUE_ORDER = ["zone_temp", "zone_upper_viol", "zone_lower_viol", "comfort_lower",
                          "comfort_upper", "outdoor_temp", "p_consumed", "time_of_day",
                          "bus_voltage", "min_voltage", "max_voltage", "p_setpoint"]
DEFAULT_BUILDING_OBS = OrderedDict([          # defaults.py:2-10
    ("zone_upper_viol", (-10.0, 10.0)), ("zone_lower_viol", (-10.0, 10.0)),
    ("comfort_lower", (20.0, 25.0)), ("comfort_upper", (25.0, 30)),
    ("outdoor_temp", (0.0, 56.0)), ("p_consumed", (0.0, 100.0)), ("time_of_day", (0.0, 1.0))])


def building_obs_layout(obs_config):
    """obs_space.py:66-101: labels and bounds in DEFAULT_OBS_CONFIG key order."""
    for k in obs_config:
        assert k in OBS_KEYS_BOUND_ORDER, f"invalid key {k}"
    labels, low, high = [], [], []
    for k in OBS_KEYS_BOUND_ORDER:
        if k not in obs_config:
            continue
        reps = 5 if k in PER_ZONE_KEYS else 1
        for z in range(reps):
            labels.append(f"{k}_{z}" if reps > 1 else k)
            low.append(float(obs_config[k][0]))
            high.append(float(obs_config[k][1]))
    return labels, np.array(low), np.array(high)


def exogenous_slice(start_time=None, end_time=None, table=None, index0=None):
    """Rows of the exogenous table between two timestamps, inclusive
    (five_zone_rom_env.py:30-41 ``df.loc[start:end]`` on a 5-minute index)."""
    import pandas as pd

    from oracle.exogenous import START, synthetic_exogenous_table
    if table is None:
        table = synthetic_exogenous_table()
        index0 = pd.Timestamp(START)
    n = table.shape[0]
    step = pd.Timedelta(300, "s")
    lo = 0 if not start_time else int(np.ceil((pd.Timestamp(start_time) - index0) / step))
    hi = n - 1 if not end_time else int(np.floor((pd.Timestamp(end_time) - index0) / step))
    lo, hi = max(lo, 0), min(hi, n - 1)
    if hi < lo:
        raise ValueError("start/end times select no exogenous rows")
    return table[lo:hi + 1]


class FiveZoneROMEnv(ComponentEnv):
    """five_zone_rom_env.py:60-308 + five_zone_rom_dynamics.py."""

    def __init__(self, name=None, obs_config=None, start_time=None, end_time=None,
                 comfort_bounds=None, zone_temp_init=None, max_episode_steps=None,
                 rescale_spaces=True, **kwargs):
        super().__init__(name=name)
        self.rescale_spaces = rescale_spaces
        self.obs_config = obs_config if obs_config is not None else DEFAULT_BUILDING_OBS
        self.zone_temp_init = np.array(zone_temp_init, dtype=np.float64) \
            if zone_temp_init is not None else 27.0 * np.ones(5)
        self.exo = exogenous_slice(start_time, end_time)     # [T, 16]: T_oa, Qsol5, Qcool5, Qint5
        a = assets()
        self.A = a["building/ss_A"].copy()
        # ss_B is rounded to float32 every step before the float64 product (dynamics.py:51)
        self.B = a["building/ss_B"].astype(np.float32).astype(np.float64)
        self.C = a["building/ss_C"].copy()
        self.K = a["building/ss_K"].copy()
        self.mean = a["building/mean_output"].copy()
        self.sel = a["building/input_sel_list"] - 1          # 1-based in the pickle
        self.nbr = a["building/neighbors"]
        self.x = a["building/x_k0"].copy()                   # persists across resets (:94)
        max_steps = self.exo.shape[0] - 3                    # :97
        self.max_episode_steps = max_steps if max_episode_steps is None \
            else min(max_episode_steps, max_steps)
        cb = comfort_bounds if comfort_bounds is not None else COMFORT
        if isinstance(cb, tuple):
            self.comfort = np.tile(np.array(cb, dtype=np.float64), (self.exo.shape[0], 1))
        else:
            self.comfort = np.asarray(cb, dtype=np.float64)[:self.exo.shape[0], :2]
        self.act_low = np.array(FLOW_LO + [T_DIS_LO])
        self.act_high = np.array(FLOW_HI + [T_DIS_HI])
        self._action_space = Box(self.act_low, self.act_high)
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self._obs_labels, lo, hi = building_obs_layout(self.obs_config)
        self._observation_space = Box(lo, hi)
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self.time_index = None
        self.state = None

    # ---- dynamics (five_zone_rom_dynamics.py)
    def _u(self, action, row, use_q_cool):
        # build_u_vector :12-41
        t_oa, q_sol, q_cool, q_int = row[0], row[1:6], row[6:11], row[11:16]
        T = self.zone_temp
        u = np.zeros((5, 4))
        for z in range(5):
            cand = np.zeros(8)
            cand[0] = t_oa - T[z]
            cand[1] = q_sol[z]
            cand[2] = q_int[z]
            for i, y in enumerate(self.nbr[z]):
                cand[3 + i] = T[y] - T[z]
            cand[7] = q_cool[z] if use_q_cool else action[z] * (action[-1] - T[z])
            u[z] = cand[self.sel[z]]
        return u

    def _advance_x(self, u):
        # state_update :44-55
        for z in range(5):
            self.x[z] = self.A[z] * self.x[z] + float(np.matmul(self.B[z].reshape(1, -1),
                                                                u[z].reshape(-1, 1)).squeeze())

    def _temps(self):
        return self.C * self.x + self.mean                  # temp_dynamics :75-85

    def reset(self, **obs_kwargs):
        # :147-180
        self.time_index = 0
        self.state = None
        self.zone_temp = self.zone_temp_init.copy()
        self.p_consumed = 0.0
        u = self._u(None, self.exo[0], use_q_cool=True)
        for _ in range(2):                                   # filter_update x2, dynamics.py:58-72
            self._advance_x(u)
            for z in range(5):
                self.x[z] += self.K[z] * ((self.zone_temp[z] - self.mean[z]) - self.C[z] * self.x[z])
        self.zone_temp = self._temps()
        obs, _ = self.get_obs(**obs_kwargs)
        return obs

    def step(self, action, **obs_kwargs):
        action = np.asarray(action, dtype=np.float64)
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        return self.step_(action, **obs_kwargs)

    def step_(self, action, **obs_kwargs):
        # :189-225
        action = np.array(action).squeeze()
        row = self.exo[self.time_index]                      # pre-increment exogenous row
        u = self._u(action, row, use_q_cool=False)
        self._advance_x(u)
        self.zone_temp = self._temps()
        flow = np.sum(action[:-1])
        self.p_consumed = (0.0076 * flow ** 3 + 4.8865) + max(0.0, flow * (row[0] - action[-1]))
        rew, _ = self.step_reward()                          # uses the STALE state dict (:215)
        self.time_index += 1
        obs, state = self.get_obs(**obs_kwargs)
        return np.array(obs), rew, self.is_terminal(), state

    def get_obs(self, **obs_kwargs):
        # :228-283
        lb, ub = self.comfort[self.time_index]
        t_oa = self.exo[self.time_index][0]
        bus_voltage = obs_kwargs.get("bus_voltage")
        p_setpoint = obs_kwargs.get("p_setpoint")
        st = OrderedDict()
        for z in range(5):
            st[f"zone_temp_{z}"] = self.zone_temp[z]
        for z in range(5):
            st[f"zone_upper_viol_{z}"] = self.zone_temp[z] - ub
        for z in range(5):
            st[f"zone_lower_viol_{z}"] = lb - self.zone_temp[z]
        st["comfort_lower"] = lb
        st["comfort_upper"] = ub
        st["outdoor_temp"] = t_oa
        st["p_consumed"] = self.p_consumed
        st["time_of_day"] = 1.0 * self.time_index / self.max_episode_steps
        for k in ("bus_voltage", "min_voltage", "max_voltage"):
            st[k] = bus_voltage if bus_voltage is not None else 1.0
        st["p_setpoint"] = p_setpoint if p_setpoint is not None else np.inf
        st.update(obs_kwargs)
        self.state = st
        obs = np.array([v for k, v in st.items() if k in self._obs_labels], dtype=np.float64)
        obs = np.clip(obs, self._observation_space.low, self._observation_space.high).squeeze()
        if self.rescale_spaces:
            obs = to_scaled(obs, self._observation_space.low, self._observation_space.high)
        return obs.copy(), st.copy()

    def step_reward(self, **kwargs):
        # base-class reward :286-294 (a 5-vector; both lists read the *upper* violation)
        v = np.array([self.state[f"zone_upper_viol_{z}"] for z in range(5)])
        return v ** 2 + v ** 2, {}

    def is_terminal(self):
        return self.time_index == self.max_episode_steps - 1

    @property
    def real_power(self):
        return self.state["p_consumed"]                      # :304-308


class FiveZoneROMThermalEnergyEnv(FiveZoneROMEnv):
    """five_zone_rom_env.py:312-335."""

    def step_reward(self, **kwargs):
        alpha = 0.2
        energy = -self.state["p_consumed"] / 12.0
        err = [max(self.state[f"zone_upper_viol_{z}"], self.state[f"zone_lower_viol_{z}"], 0.0)
               for z in range(5)]
        comfort = -(sum([e ** 2 for e in err]))
        return alpha * energy * 0.5 + (1.0 - alpha) * comfort, \
            {"comfort_rew": comfort, "energy_rew": energy}


# ======================================================================== composite agent
class MultiComponentEnv(ComponentEnv):
    """base.py:74-182: ordered list of components behind one agent."""

    def __init__(self, name=None, components=None, **kwargs):
        super().__init__(name=name)
        self.envs = [c["cls"](name=c["name"], **c["config"]) for c in components]
        self.observation_space = {e.name: e.observation_space for e in self.envs}
        self.action_space = {e.name: e.action_space for e in self.envs}
        self._obs_labels_dict = {e.name: e.obs_labels for e in self.envs}
        labels = []
        for e in self.envs:
            labels += e.obs_labels
        self._obs_labels = list(set(labels))

    def reset(self, **kwargs):
        for e in self.envs:
            e.reset(**kwargs)            # NB: unfiltered kwargs at reset (base.py:110)
        return self.get_obs(**kwargs)

    def step(self, action, **kwargs):
        real_power, obs, dones, metas = 0.0, {}, [], {}
        for e in self.envs:
            kw = {k: v for k, v in kwargs.items() if k in e.obs_labels}
            ob, _, done, meta = e.step(action[e.name], **kw)     # component reward discarded
            obs[e.name] = ob.copy()
            dones.append(done)
            metas[e.name] = dict(meta)
            real_power += e.real_power
        self._real_power = real_power
        rew, _ = self.step_reward()                              # recomputed post-step (:137)
        return obs, rew, any(dones), metas

    def step_reward(self, **kwargs):
        total, meta = 0.0, {}
        for e in self.envs:
            r, m = e.step_reward()
            total += r
            meta[e.name] = dict(m)
        return total, meta

    def get_obs(self, **kwargs):
        obs, meta = {}, {}
        for e in self.envs:
            kw = {k: v for k, v in kwargs.items() if k in e.obs_labels}
            obs[e.name], meta[e.name] = e.get_obs(**kw)
        return obs, meta

    @property
    def env_dict(self):
        return {e.name: e for e in self.envs}

    @property
    def obs_labels_dict(self):
        return self._obs_labels_dict
