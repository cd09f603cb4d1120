"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement of the reference's Home-Steward ("HS") composite: one house = an ordered list of
components (PV -> storage -> EV charger -> other devices in the shipped configuration) that
share the step's available solar / battery / grid power and pay a blended energy cost.
Plain-Python float64, one env at a time, same order of floating-point operations as the
reference so that the golden traces recorded from the unmodified reference
(tests/golden/make_golden_hs.py -> tests/golden/hs_*.npz) replay bit for bit.

Follows (paths relative to the reference root):
  gridworld/base_hs.py:12-199                                   HSMultiComponentEnv
  gridworld/agents/pv/pv_profile_env_hs.py:15-169               HSPVEnv
  gridworld/agents/energy_storage/energy_storage_env_hs.py:10-273   HSEnergyStorageEnv
  gridworld/agents/vehicles/ev_charging_env_hs.py:15-336        HSEVChargingEnv
  gridworld/agents/devices/devices_env_hs.py:14-205             HSDevicesEnv
  gridworld/scenarios/heterogeneous_hs.py:47-58                 make_env_config

The "meta state" dict of the reference carries a few numbers from component to component
within a step and from step to step: pv_power, es_power, grid_power (what is still available),
pv_cost, es_cost, grid_cost.  Only those keys are restated; the per-device ``step_meta``
telemetry records are not (they do not feed back into the dynamics).
"""
from __future__ import annotations

import numpy as np

from oracle.components import Box, ComponentEnv, _unit_box, to_raw, to_scaled

META_KEYS = ("grid_cost", "es_cost", "grid_power", "pv_power", "es_power", "pv_cost")


class HSPVEnv(ComponentEnv):
    """pv_profile_env_hs.py:15-169.  Action in [0.98, 1] = fraction of the available power."""

    def __init__(self, name=None, profile_csv=None, profile_path=None, profile_data=(),
                 scaling_factor=1.0, rescale_spaces=True, grid_aware=False,
                 max_episode_steps=None, minutes_per_step=5, **kwargs):
        super().__init__(name=name)
        if grid_aware:
            raise NotImplementedError("grid_aware HS PV is not restated")
        if len(profile_data) == 0:
            raise ValueError("the oracle takes the profile as data (the JSON config embeds it)")
        self.scaling_factor, self.rescale_spaces = scaling_factor, rescale_spaces
        self.data = [scaling_factor * i for i in np.array(profile_data)]          # :58-74
        self.episode_length = len(self.data)
        if max_episode_steps is not None:
            self.episode_length = min(max_episode_steps, self.episode_length)
        self._obs_labels = ["real_power"]
        self._observation_space = Box(np.array([-np.max(self.data)]), np.array([0.0]))
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self._action_space = Box(np.array([0.98]), np.array([1.0]))               # :100-101
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self.index = 0

    def get_obs(self, meta):
        raw = np.array([-self.data[self.index]])                                   # :110-126
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        meta = dict(meta)
        meta["real_power"] = -raw[0]
        meta["pv_power"] = meta["real_power"]
        return obs, meta

    def reset(self, meta):
        self.index = 0
        return self.get_obs(meta)

    def is_terminal(self):
        return self.index == self.episode_length

    def step(self, action, meta):
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        obs, m = self.get_obs(meta)                                                # :147
        self._real_power = np.float64((action * m["real_power"]).squeeze())       # :149
        self.index += 1
        m["pv_power"] = self._real_power                                           # :153
        return obs, 0, self.is_terminal(), m

    def step_reward(self, meta):
        return 0


class HSEnergyStorageEnv(ComponentEnv):
    """energy_storage_env_hs.py:10-273."""

    def __init__(self, name=None, storage_range=(3.0, 50.0), initial_storage_mean=30.0,
                 initial_storage_std=5.0, charge_efficiency=0.95, discharge_efficiency=0.9,
                 max_power=15.0, max_episode_steps=288, control_timedelta=None,
                 rescale_spaces=True, initial_storage_cost=0.0, max_storage_cost=0.55, **kwargs):
        super().__init__(name=name)
        self.current_cost = initial_storage_cost                                   # :39 (never reset)
        self.storage_range = storage_range
        self.initial_storage_mean, self.initial_storage_std = initial_storage_mean, initial_storage_std
        self.charge_efficiency, self.discharge_efficiency = charge_efficiency, discharge_efficiency
        self.max_power, self.rescale_spaces = max_power, rescale_spaces
        self.max_storage_cost = max_storage_cost
        self.max_episode_steps = max_episode_steps
        seconds = 300 if control_timedelta is None else control_timedelta.seconds
        self.control_interval_in_hr = seconds / 3600.0
        self._obs_labels = ["stage_of_charge", "cost"]
        self._observation_space = Box(np.array([storage_range[0], 0.00]),
                                      np.array([storage_range[1], max_storage_cost]))
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self._action_space = Box(np.array([-1.0]), np.array([1.0]))
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self.current_storage = None
        self.delta_cost = 0.0
        self.simulation_step = 0

    def reset(self, meta, init_storage=None):
        self.simulation_step = 0                                                   # :82-107
        if init_storage is None:
            from scipy.stats import truncnorm
            self.current_storage = float(truncnorm(-1, 1).rvs() * self.initial_storage_std
                                         + self.initial_storage_mean)
        else:
            self.current_storage = np.clip(float(init_storage), self.storage_range[0],
                                           self.storage_range[1])
        return self.get_obs(meta)

    def validate_power(self, power):
        st_min, st_max = self.storage_range[0], self.storage_range[1]             # :111-143
        if power > 0:
            delta = power * self.control_interval_in_hr / self.discharge_efficiency
            if self.current_storage <= st_min:
                power = 0.0
            elif self.current_storage - delta < st_min:
                delta = self.current_storage - st_min
                power = delta / self.control_interval_in_hr * self.discharge_efficiency
        elif power < 0:
            delta = -(power * self.control_interval_in_hr * self.charge_efficiency)
            if self.current_storage >= st_max:
                power = 0.0
            elif self.current_storage + delta > st_max:
                delta = st_max - self.current_storage
                power = -(delta / self.control_interval_in_hr / self.charge_efficiency)
        return power

    def get_obs(self, meta):
        raw = np.array([self.current_storage, self.current_cost])                  # :145-159
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        return obs, dict(meta)

    def step_reward(self, meta):
        step_cost = 0.0                                                            # :161-190
        if not (self._real_power < 0):
            step_cost = self.delta_cost * self.charge_efficiency * self._real_power \
                * self.control_interval_in_hr
        reward = -step_cost
        if meta["pv_power"] > 0.0 and meta["es_power"] > 0.0 \
                and self.current_storage < max(self.storage_range):
            reward -= self.max_storage_cost * (max(self.storage_range) - self.current_storage)
        return reward

    def step(self, action, meta):
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        power = self.validate_power(action[0] * self.max_power)                    # :201-203
        m = dict(meta)
        solar_capacity, solar_cost = m["pv_power"], m["pv_cost"]
        grid_cost, grid_capacity = m["grid_cost"], m["grid_power"]
        if power == 0.0:
            self.delta_cost = 0.0
            m["es_power"] = 0.0
        elif power < 0.0:                                                          # charging :221-245
            delta_storage = self.charge_efficiency * power * self.control_interval_in_hr
            solar_used = min(-power, solar_capacity)
            grid_used = min(grid_capacity, -power - solar_used)
            self.delta_cost = (solar_cost * solar_used + grid_cost * grid_used) / (solar_used + grid_used)
            self.current_cost = (self.current_storage * self.current_cost
                                 - delta_storage * self.delta_cost) / (self.current_storage - delta_storage)
            self.current_storage -= delta_storage
            self.current_storage = min(self.current_storage, self.storage_range[1])
            m["pv_power"] = max(0.0, solar_capacity - solar_used)
            m["grid_power"] = max(0.0, grid_capacity - grid_used)
            m["es_power"] = 0.0
        elif power > 0.0:                                                          # discharging :248-253
            delta_storage = power * self.control_interval_in_hr / self.discharge_efficiency
            self.current_storage = max(self.current_storage - delta_storage, self.storage_range[0])
            m["es_power"] = power
        m["es_cost"] = 0                                                           # :256
        self._real_power = -power
        obs, _ = self.get_obs(m)
        rew = self.step_reward(m)
        self.simulation_step += 1
        return obs, rew, self.simulation_step == self.max_episode_steps, m


class HSEVChargingEnv(ComponentEnv):
    """ev_charging_env_hs.py:15-336."""

    def __init__(self, num_vehicles=100, minutes_per_step=5, max_charge_rate_kw=7.0,
                 max_episode_steps=None, unserved_penalty=1.0, peak_penalty=1.0,
                 peak_threshold=10.0, reward_scale=1e5, name=None, randomize=False,
                 vehicle_csv=None, vehicle_multiplier=1, rescale_spaces=True,
                 max_charge_cost=0.55, profile_data=None, **kwargs):
        super().__init__(name=name)
        if profile_data:                           # pd.read_json(..., orient="split") (:68-69)
            cols = profile_data["columns"]
            vehicles = {c: [row[cols.index(c)] for row in profile_data["data"]]
                        for c in ("start_time_min", "end_time_park_min", "energy_required_kwh")}
        elif vehicle_csv is None:                  # vehicles_hs.csv next to the module (:72-73)
            from oracle.components import assets
            vehicles = {c: assets()[f"hs/vehicles/{c}"]
                        for c in ("start_time_min", "end_time_park_min", "energy_required_kwh")}
        else:
            raise ValueError("the oracle reads the packaged vehicle table only")
        self.num_vehicles, self.max_charge_rate_kw = num_vehicles, max_charge_rate_kw
        self.minutes_per_step, self.vehicle_multiplier = minutes_per_step, vehicle_multiplier
        self.rescale_spaces, self.unserved_penalty = rescale_spaces, unserved_penalty
        steps = max_episode_steps if max_episode_steps is not None else np.inf
        self.max_episode_steps = min(steps, 24 * 60 / minutes_per_step)           # :54-55
        self.simulation_times = np.arange(
            0, (self.max_episode_steps + 1) * minutes_per_step, minutes_per_step)
        self._energy0 = np.asarray(vehicles["energy_required_kwh"], dtype=np.float64) * vehicle_multiplier
        rnd = lambda x: x - x % minutes_per_step                                   # :333-335
        self._start = rnd(np.asarray(vehicles["start_time_min"], dtype=np.float64))
        self._end = rnd(np.asarray(vehicles["end_time_park_min"], dtype=np.float64))
        emax = self._energy0.max()
        high = np.array([self.simulation_times[-1], num_vehicles, num_vehicles * max_charge_rate_kw,
                         num_vehicles * emax, emax / (minutes_per_step / 60.), emax,
                         max_charge_cost], dtype=np.float64)                        # :87-103
        self._observation_space = Box(np.zeros(7), high)
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self._action_space = Box(np.array([0.0]), np.array([1.0]))
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self._obs_labels = ["time", "num_active_vehicles", "real_power_consumed", "real_power_demand",
                            "mean_charge_rate_deficit", "real_power_unserved", "current_cost"]
        self.state = [None] * 7

    def reset(self, meta):
        self.time_index = 0                                                        # :134-157
        self.time = self.simulation_times[0]
        self.charging_vehicles = []
        self.energy = self._energy0.copy()
        self._real_power = 0.0
        self.current_cost = getattr(self, "current_cost", None)
        self.step(None, meta)                      # hidden step; its meta updates are discarded
        return self.get_obs(meta)

    def get_obs(self, meta):
        raw = np.array(self.state, dtype=np.float64)
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        return obs.copy(), dict(meta)

    def step_reward(self, meta):
        step_cost = self.current_cost * self._real_power                           # :178-191
        return -(step_cost + self.unserved_penalty * self.state[5] ** 2)

    def step(self, action, meta):
        action = action if action is not None else self._action_space.low         # :199-201
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        action_kwh = (action[0] * self.max_charge_rate_kw) * (self.minutes_per_step / 60.)
        n = len(self.energy)
        charging = [i for i in range(n)
                    if self.time >= np.floor(self._start[i]) and self.time <= np.floor(self._end[i])
                    and self.energy[i] > 0.]                                       # :207-213
        departed = sorted(set(self.charging_vehicles) - set(charging))
        consumed, demand, deficits = 0., 0., []
        for i in charging:                                                         # :226-254
            need = self.energy[i]
            demand += need
            time_left_h = (self._end[i] - self.time) / 60.
            if time_left_h <= 0:
                continue
            deficits.append(max(0, self.max_charge_rate_kw - need / time_left_h))
            d = min(action_kwh, need)
            self.energy[i] -= d
            consumed += d
        self.time = self.simulation_times[self.time_index]                         # :260 (no +1)
        self.charging_vehicles = charging
        unserved = 0.
        for i in departed:
            unserved += self.energy[i]
        self.state[5] = unserved
        self.state[0] = self.time
        self.state[1] = self.vehicle_multiplier * len(charging)
        self.state[2] = self.vehicle_multiplier * consumed
        self.state[3] = self.vehicle_multiplier * demand
        self.state[4] = 0 if len(deficits) == 0 else np.mean(deficits)
        self._real_power = self.vehicle_multiplier * consumed
        power = self._real_power * (60.0 / self.minutes_per_step)                  # :285
        m = dict(meta)
        solar_capacity, battery_capacity, grid_capacity = m["pv_power"], m["es_power"], m["grid_power"]
        if power == 0.0 or action[0] == 0.0:
            self.current_cost = 0.0
        else:                                                                      # :295-318
            solar_cost, battery_cost, grid_cost = m["pv_cost"], m["es_cost"], m["grid_cost"]
            solar_used = min(power, solar_capacity)
            battery_used = grid_used = 0
            if battery_cost < grid_cost:
                battery_used = min(battery_capacity, power - solar_used)
                grid_used = min(grid_capacity, power - solar_used - battery_used)
            elif battery_cost >= grid_cost:
                grid_used = min(grid_capacity, power - solar_used)
                battery_used = min(battery_capacity, power - solar_used - grid_used)
            if solar_used + grid_used + battery_used > 0:
                self.current_cost = (solar_cost * solar_used + grid_cost * grid_used
                                     + battery_cost * battery_used) / (solar_used + grid_used + battery_used)
            m["pv_power"] = max(0.0, solar_capacity - solar_used)
            m["es_power"] = max(0.0, battery_capacity - battery_used)
            m["grid_power"] = max(0.0, grid_capacity - grid_used)
        self.state[6] = self.current_cost
        obs, _ = self.get_obs(m)
        rew = self.step_reward(m)
        done = self.time_index == self.max_episode_steps
        self.time_index += 1
        return obs, rew, done, m


class HSDevicesEnv(ComponentEnv):
    """devices_env_hs.py:14-205.  ``profile_data``: {label: [values per step]}."""

    def __init__(self, name=None, profile_csv=None, profile_path=None, profile_data=None,
                 scaling_factor=1.0, rescale_spaces=True, max_episode_steps=None,
                 minutes_per_step=5, **kwargs):
        super().__init__(name=name)
        if not profile_data:
            raise ValueError("the oracle takes the profile as data (the JSON config embeds it)")
        self.rescale_spaces, self.minutes_per_step = rescale_spaces, minutes_per_step
        self._obs_labels = list(profile_data.keys())
        cols = np.array([v for v in profile_data.values()], dtype=np.float64).T   # :55-57
        self.bounds_high = cols.max(axis=0)                  # bounds from the unscaled frame (:80-82)
        # `data` aliases the frame's values and is scaled in place (:70-71); pandas 3
        # copy-on-write makes .values a copy, the reference then reads the frame: unscaled
        self.data = cols * scaling_factor
        self.frame = cols
        self.episode_length = len(self.data)
        if max_episode_steps is not None:
            self.episode_length = min(max_episode_steps, self.episode_length)
        self._observation_space = Box(np.zeros(cols.shape[1]), self.bounds_high)
        self.observation_space = _unit_box(self._observation_space, rescale_spaces)
        self._action_space = Box(np.array([0.99]), np.array([1.0]))
        self.action_space = _unit_box(self._action_space, rescale_spaces)
        self.index = 0
        self.current_cost = 0.0

    def get_obs(self, meta):
        raw = np.array(self.data[self.index])                                      # :104-120
        obs = to_scaled(raw, self._observation_space.low, self._observation_space.high) \
            if self.rescale_spaces else raw
        return obs, dict(meta)

    def reset(self, meta):
        self.index = 0
        return self.get_obs(meta)

    def step_reward(self, meta):
        return -(self.current_cost * self._real_power * (self.minutes_per_step / 60.0))   # :128-140

    def step(self, action, meta):
        if self.rescale_spaces:
            action = to_raw(action, self._action_space.low, self._action_space.high)
        obs, m = self.get_obs(meta)
        returned = dict(m)     # obs_meta = kwargs.copy() (:163) is taken BEFORE the allocation
        #                        below, so what the devices consume never reaches the meta state
        total = sum([self.frame[self.index, j] for j in range(self.frame.shape[1])])       # :165
        self._real_power = np.float64((action * total).squeeze())
        solar_capacity, battery_capacity, grid_capacity = m["pv_power"], m["es_power"], m["grid_power"]
        if round(self._real_power, 3) == 0.0:                                      # :174-175
            self.current_cost = 0.0
        else:
            solar_used = min(self._real_power, solar_capacity)
            battery_used = min(battery_capacity, self._real_power - solar_used)
            grid_used = min(grid_capacity, self._real_power - solar_used - battery_used)
            self.current_cost = (m["pv_cost"] * solar_used + m["grid_cost"] * grid_used
                                 + m["es_cost"] * battery_used) / (solar_used + grid_used + battery_used)
            m["pv_power"] = max(0.0, solar_capacity - solar_used)
            m["es_power"] = max(0.0, battery_capacity - battery_used)
            m["grid_power"] = max(0.0, grid_capacity - grid_used)
        rew = self.step_reward(m)
        self.index += 1
        return obs, rew, self.index == self.episode_length, returned


class HSMultiComponentEnv:
    """base_hs.py:12-199."""

    def __init__(self, name=None, components=None, start_time="", end_time="",
                 control_timedelta=None, max_grid_power=48, max_episode_steps=None,
                 rescale_spaces=True, grid_cost=None, timestamps=None, **kwargs):
        self.name, self.max_grid_power = name, max_grid_power
        self.envs = [c["cls"](name=c["name"], **c["config"]) for c in components]
        self.observation_space = {e.name: e.observation_space for e in self.envs}
        self.action_space = {e.name: e.action_space for e in self.envs}
        self._grid_cost_data = grid_cost
        self.meta_state = {"grid_cost": None, "es_cost": 0.0, "grid_power": max_grid_power,
                           "pv_power": None, "es_power": 0.0, "pv_cost": 0.0}      # :53-61
        self._real_power = 0

    @property
    def real_power(self):
        return self._real_power

    def reset(self, init_storage=None):
        self.time_index = 0                                                        # :67-92
        self.meta_state["grid_cost"] = self._grid_cost_data[0]
        self.meta_state["grid_power"] = self.max_grid_power
        m = dict(self.meta_state)
        obs = {}
        for e in self.envs:
            if isinstance(e, HSEnergyStorageEnv):
                obs[e.name], m = e.reset(m, init_storage=init_storage)
            else:
                obs[e.name], m = e.reset(m)
        for e in self.envs:                        # get_obs of every component (:94-118)
            obs[e.name], _ = e.get_obs(m)
        return obs

    def step(self, action):
        real_power, obs, dones = 0, {}, []                                         # :120-178
        self.meta_state["grid_cost"] = self._grid_cost_data[self.time_index]
        self.meta_state["grid_power"] = self.max_grid_power
        for e in self.envs:
            o, _, d, m = e.step(action[e.name], dict(self.meta_state))
            obs[e.name] = o.copy()
            dones.append(d)
            real_power += e.real_power
            for k in META_KEYS:                    # self.meta_state.update(subcomp_meta) (:162)
                self.meta_state[k] = m[k]
        self._real_power = real_power
        reward = 0.
        for e in self.envs:                        # post-step rewards on the final meta (:176, :184-199)
            reward += e.step_reward(self.meta_state)
        self.time_index += 1
        return obs, reward, any(dones), dict(self.meta_state)
