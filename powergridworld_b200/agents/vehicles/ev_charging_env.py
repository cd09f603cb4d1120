"""EV charging station description (gridworld/agents/vehicles/ev_charging_env.py:17-275).
Dynamics: csrc/component_math.cuh ev_advance / ev_step / ev_reset.

The vehicle roster (arrival, departure, energy) is shared by every env, so the
"who is parked at minute t" half of the reference's charging set is a static
per-step list compiled here; only the "still needs energy" half is per-env state.

``randomize=True`` (:154-157): every reset draws ``num_vehicles`` rows of the table without
replacement -- NumPy's global RNG, the draw ``DataFrame.sample`` makes -- and the station's
parameter and event columns are rebuilt and pushed to the device (pgw_update_tables).  The
drawn roster is shared by all envs of the batch.
"""
from collections import OrderedDict

import numpy as np

from powergridworld_b200 import _native as N
from powergridworld_b200 import assets, spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space


def _popcount(word):
    """Set bits of a 32-bit state word: a NumPy scalar (one env) or an int32 tensor [E]."""
    if isinstance(word, (int, np.integer)):
        return np.float64(bin(int(word) & 0xFFFFFFFF).count("1"))
    x = word.long() & 0xFFFFFFFF
    x = x - ((x >> 1) & 0x55555555)
    x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F
    return ((x * 0x01010101) & 0xFFFFFFFF).__rshift__(24).double()


def _read_vehicle_csv(path):
    import pandas as pd
    df = pd.read_csv(path)
    return {c: df[c].values.astype(np.float64)
            for c in ("start_time_min", "end_time_park_min", "energy_required_kwh")}


class EVChargingEnv(ComponentEnv):

    def __init__(self, num_vehicles: int = 100, minutes_per_step: int = 5,
                 max_charge_rate_kw: float = 7.0, max_episode_steps: int = None,
                 unserved_penalty: float = 1., peak_penalty: float = 1.,
                 peak_threshold: float = 10., reward_scale: float = 1e5, name: str = None,
                 randomize: bool = False, vehicle_csv: str = None, vehicle_multiplier: int = 1,
                 rescale_spaces: bool = True, **kwargs):
        super().__init__(name=name)
        self.num_vehicles = num_vehicles
        self.max_charge_rate_kw = max_charge_rate_kw
        self.minutes_per_step = minutes_per_step
        self.randomize = randomize
        self.vehicle_multiplier = vehicle_multiplier
        self.rescale_spaces = rescale_spaces
        self.unserved_penalty = unserved_penalty
        self.peak_penalty = peak_penalty
        self.peak_threshold = peak_threshold
        self.reward_scale = reward_scale
        mes = max_episode_steps if max_episode_steps is not None else np.inf
        self.max_episode_steps = min(mes, 24 * 60 / minutes_per_step)
        self.simulation_times = np.arange(
            0, self.max_episode_steps * minutes_per_step, minutes_per_step)
        cols = _read_vehicle_csv(vehicle_csv) if vehicle_csv else {
            c: assets.array(f"vehicles/{c}")
            for c in ("start_time_min", "end_time_park_min", "energy_required_kwh")}
        energy = cols["energy_required_kwh"] * self.vehicle_multiplier            # :72
        rnd = lambda x: x - x % self.minutes_per_step                             # :273-275
        self._roster_start = rnd(cols["start_time_min"])
        self._roster_end = rnd(cols["end_time_park_min"])
        self._roster_energy = energy
        emax = energy.max()
        obs_bounds = OrderedDict({
            "time": (0, self.simulation_times[-1]),
            "num_active_vehicles": (0, self.num_vehicles),
            "real_power_consumed": (0, self.num_vehicles * self.max_charge_rate_kw),
            "real_power_demand": (0, self.num_vehicles * emax),
            "mean_charge_rate_deficit": (0, emax / (self.minutes_per_step / 60.)),
            "real_power_unserved": (0, emax)})
        self._observation_space = spaces.Box(
            low=np.array([x[0] for x in obs_bounds.values()], dtype=np.float64),
            high=np.array([x[1] for x in obs_bounds.values()], dtype=np.float64),
            shape=(len(obs_bounds),), dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(low=0., high=1., shape=(1,), dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)
        self.state = OrderedDict({k: None for k in obs_bounds.keys()})
        self._obs_labels = list(self.state.keys())

    def _reset_result(self, obs):
        return obs, {}                              # (obs, {}) like the reference (:168)

    def _terminal_after(self):
        # reset leaves time_index == 1 (hidden step, :163); terminal at max_episode_steps - 1
        return self.max_episode_steps - 2

    def _meta(self, ctx) -> dict:
        """The station's state dict (get_obs returns it as meta, :121-128; step_reward adds
        nothing, :135-142).  Raw values: the delivered observation, un-scaled when the spaces are
        rescaled (the inverse map reproduces the raw value to an ulp or two)."""
        o0 = self._slot["obs"][0]
        hi = self._observation_space.high
        out = OrderedDict()
        for j, key in enumerate(self._obs_labels):
            v = ctx.obs(o0 + j)
            out[key] = (v + 1.0) * (0.5 * hi[j]) if self.rescale_spaces else v
        # the state dict is NOT clipped to the observation bounds (:121-128 copies self.state), the
        # scaled observation is: with vehicle_multiplier > 1 the vehicle count exceeds its bound
        # num_vehicles (:81, :247), so it is counted from the charging-set words instead
        s0, words = self._slot["si"][0], (self.num_vehicles + 31) // 32
        count = None
        for w in range(words):
            word = ctx.si(s0 + w)
            c = _popcount(word)
            count = c if count is None else count + c
        out["num_active_vehicles"] = self.vehicle_multiplier * count
        # meta.update(rew_meta) (:259-262): the unserved-energy REWARD TERM replaces the state
        # entry of the same name, and the peak term is added
        over = out["real_power_consumed"] - self.peak_threshold
        over = over * (over > 0)
        out["peak_reward"] = -self.peak_penalty * over ** 2
        out["real_power_unserved"] = -self.unserved_penalty * out["real_power_unserved"] ** 2
        return out

    _rows = None
    _num_envs = 1            # set by the MultiAgentEnv that compiles this station

    @property
    def _per_env(self) -> bool:
        """A randomised station in a batch: every env instance samples its own roster, like every
        instance of the reference's class does in its own reset (:154-157)."""
        return bool(self.randomize) and self._num_envs > 1

    def _draw_roster(self):
        """df.sample(n) (:155): np.random.choice(len(df), n, replace=False), rows kept in drawn
        order (the order fixes the index each vehicle gets after reset_index, :157).  In a batch:
        one draw per env instance, env 0 first -- ``_rows`` is [num_envs, n]."""
        draw = lambda: np.random.choice(len(self._roster_energy), size=self.num_vehicles, replace=False)
        self._rows = np.stack([draw() for _ in range(self._num_envs)]) if self._per_env else draw()

    def _per_env_rows(self):
        """(window words uint32 [n, E], initial energies float64 [n, E]) of the current per-env
        draw, as PGW_F_EV_PER_ENV lays them out: floor(start) << 16 | floor(end) in minutes."""
        rows = self._rows                                   # [E, n]
        start = np.floor(self._roster_start[rows])
        end = np.floor(self._roster_end[rows])
        if not (np.array_equal(end, self._roster_end[rows]) and start.min() >= 0 and end.min() >= 0
                and max(start.max(), end.max()) < 65535):
            raise NotImplementedError("per-env rosters need parking times that are whole minutes in "
                                      "[0, 65535) after rounding to the step (ev_charging_env.py:75-76)")
        words = (start.astype(np.uint32) << np.uint32(16)) | end.astype(np.uint32)
        return np.ascontiguousarray(words.T), np.ascontiguousarray(self._roster_energy[rows].T)

    def _static_dpar(self):
        hi = self._observation_space.high
        dpar = [self.max_charge_rate_kw, self.minutes_per_step / 60., float(self.vehicle_multiplier),
                self.unserved_penalty, self.peak_penalty, self.peak_threshold, self.reward_scale]
        return dpar + list(hi) + list(1.0 / hi) + [1.0 / self.reward_scale, 1.0 / 60.0]

    def _retable(self):
        """(dpar, dtab_fn, itab_fn) of the current roster; widths do not depend on the draw."""
        n = self.num_vehicles
        if self._per_env:                                   # the roster lives in per-env state rows
            times = self.simulation_times
            n_ev = len(times) - 1
            return (self._static_dpar(),
                    lambda r: [times[min(r, n_ev - 1)], times[min(r, n_ev - 1) + 1]], None)
        rows = self._rows if (self.randomize and self._rows is not None) else np.arange(n)
        start = np.floor(self._roster_start[rows])
        end = np.floor(self._roster_end[rows])
        times = self.simulation_times
        idx = np.arange(n)

        def window(k):
            t = times[k]
            return idx[(t >= start) & (t <= end)]          # :186-190, ascending index

        n_ev = len(times) - 1                               # events 0 .. len-2 have a "next" time
        wins = [window(k) for k in range(n_ev)]
        lefts = [np.array([], dtype=int)] + [np.setdiff1d(wins[k - 1], wins[k]) for k in range(1, n_ev)]
        # a randomised station sizes its lists for the worst draw (everyone parked at once)
        cap = n if self.randomize else \
            max(1, max(len(w) for w in wins), max(len(l) for l in lefts))
        self._cap = cap

        end_raw = self._roster_end[rows]

        def dtab_fn(r):
            k = min(r, n_ev - 1)
            left = np.zeros(cap)
            left[:len(wins[k])] = (end_raw[wins[k]] - times[k]) / 60.             # :216
            with np.errstate(divide="ignore"):
                inv = np.where(left > 0, 1.0 / left, 0.0)
            return [times[k], times[k + 1]] + list(left) + list(inv)

        def itab_fn(r):
            k = min(r, n_ev - 1)
            row = np.zeros(2 + 2 * cap, dtype=np.int32)
            row[0], row[1] = len(wins[k]), len(lefts[k])
            row[2:2 + len(wins[k])] = wins[k]
            row[2 + cap:2 + cap + len(lefts[k])] = lefts[k]
            return row

        dpar = self._static_dpar()
        dpar += list(self._roster_end[rows]) + list(self._roster_energy[rows])
        return dpar, dtab_fn, itab_fn

    def _emit(self, b, agent_index, standalone):
        n = self.num_vehicles
        dpar, dtab_fn, itab_fn = self._retable()
        if self._per_env:
            if n > 256:
                raise NotImplementedError("per-env rosters serve stations of up to 256 vehicles")
            words = (n + 31) // 32
            b.add_component(self, N.EV, agent_index,
                            flags=(N.F_RESCALE if self.rescale_spaces else 0) | N.F_EV_PER_ENV,
                            dpar=dpar, ipar=[n, words, 0], sd_rows=n, si_rows=words + n,
                            dtab_width=2, dtab_fn=dtab_fn)
            return
        cap, words = self._cap, (n + 31) // 32
        b.add_component(self, N.EV, agent_index,
                        flags=N.F_RESCALE if self.rescale_spaces else 0,
                        dpar=dpar, ipar=[n, words, cap], sd_rows=n, si_rows=words,
                        dtab_width=2 + 2 * cap, dtab_fn=dtab_fn, itab_width=2 + 2 * cap, itab_fn=itab_fn)
