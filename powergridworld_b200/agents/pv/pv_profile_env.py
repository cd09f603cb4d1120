"""PV description (gridworld/agents/pv/pv_profile_env.py:15-148).
Dynamics: csrc/component_math.cuh pv_step / pv_obs."""
import os

import numpy as np

from powergridworld_b200 import _native as N
from powergridworld_b200 import assets, spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space


def read_profile(profile_csv, profile_path=None) -> np.ndarray:
    """First CSV column with the first line consumed as a header, i.e. what
    ``pd.read_csv(f).values[:, 0]`` yields in the reference (:62-68)."""
    if profile_path is None and assets.has(f"pv/{profile_csv}"):
        return assets.array(f"pv/{profile_csv}")
    path = profile_path if profile_path is not None else profile_csv
    if not os.path.isfile(path):
        raise FileNotFoundError(f"PV profile {path!r} not found (packaged: constant.csv, "
                                "off-peak.csv, pv_profile.csv)")
    vals = []
    with open(path) as fh:
        next(fh)
        for line in fh:
            line = line.strip()
            if line:
                vals.append(float(line.split(",")[0]))
    return np.asarray(vals, dtype=np.float64)


class PVEnv(ComponentEnv):
    _volt_reward = False

    def __init__(self, name: str, profile_csv: str = None, profile_path: str = None,
                 scaling_factor: float = 1., rescale_spaces: bool = True,
                 grid_aware: bool = False, max_episode_steps: int = None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.scaling_factor = scaling_factor
        self.rescale_spaces = rescale_spaces
        self.grid_aware = grid_aware
        self.profile_csv = profile_path if profile_path is not None else profile_csv
        self.data = read_profile(profile_csv, profile_path)
        self.data *= self.scaling_factor
        self.episode_length = len(self.data)
        if max_episode_steps is not None:
            self.episode_length = min(max_episode_steps, self.episode_length)
        self._obs_labels = ["real_power"] + (["min_voltage"] if grid_aware else [])
        bounds = {"real_power": (-np.max(self.data), 0.), "min_voltage": (0.9, 1.1)}
        self._observation_space = spaces.Box(
            shape=(len(self.obs_labels),),
            low=np.array([bounds[k][0] for k in self.obs_labels]),
            high=np.array([bounds[k][1] for k in self.obs_labels]), dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(shape=(1,), low=0., high=1., dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)
        self.index = None

    def _reset_result(self, obs):
        return None                                 # PVEnv.reset returns nothing (:127-130)

    def _terminal_after(self):
        return self.episode_length - 1              # index == episode_length - 1 (:117-119)

    def _meta(self, ctx) -> dict:
        # raw available power of the row the step acted on (obs_meta of get_obs, :113, :143)
        return {"real_power": ctx.scalar(-ctx.dtab(self._slot["dtab"][0]))}

    def _emit(self, b, agent_index, standalone):
        flags = (N.F_RESCALE if self.rescale_spaces else 0) | \
                (N.F_GRID_AWARE if self.grid_aware else 0)
        if self._volt_reward:
            if not standalone:
                raise ValueError("a voltage-rewarded PV must be a top-level agent: inside a "
                                 "MultiComponentEnv the reference calls step_reward() without "
                                 "the grid variables (gridworld/base.py:151)")
            flags |= N.F_PV_VOLT_REWARD
        lo, hi = self._observation_space.low, self._observation_space.high
        with np.errstate(divide="ignore"):
            dpar = [lo[0], hi[0], 0.9, 1.1, 1.0 / (hi[0] - lo[0]), 1.0 / (1.1 - 0.9)]
        data = self.data
        # event 0 (reset) shows row 0; step t acts on row t *before* advancing (:143-145)
        b.add_component(self, N.PV, agent_index, flags=flags, dpar=dpar,
                        dtab_width=1, dtab_fn=lambda r: [data[max(r - 1, 0)]],
                        needs_grid=self.grid_aware or self._volt_reward)


class GridAwarePVEnv(PVEnv):
    """``ThisPVEnv`` of gridworld/scenarios/heterogeneous.py:46-52: reward
    ``-(1000 (min(0, v-0.95) + min(0, 1.05-v)))^2`` on the lagged feeder-minimum voltage."""
    _volt_reward = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if "min_voltage" not in self._obs_labels:
            # the reward needs the grid variable delivered to the agent (multiagent_env.py:112)
            raise ValueError("GridAwarePVEnv requires grid_aware=True")
