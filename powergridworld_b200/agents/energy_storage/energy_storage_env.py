"""Battery description (gridworld/agents/energy_storage/energy_storage_env.py:11-181).
Dynamics: csrc/component_math.cuh storage_step / storage_reset."""
import numpy as np
import pandas as pd

from powergridworld_b200 import _native as N
from powergridworld_b200 import spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space


class EnergyStorageEnv(ComponentEnv):

    def __init__(self, name: str = None, storage_range: tuple = (3.0, 50.0),
                 initial_storage_mean: float = 30.0, initial_storage_std: float = 5.0,
                 charge_efficiency: float = 0.95, discharge_efficiency: float = 0.9,
                 max_power: float = 15.0, max_episode_steps: int = 288,
                 control_timedelta: pd.Timedelta = pd.Timedelta(300, "s"),
                 rescale_spaces: bool = True, **kwargs):
        super().__init__(name=name)
        self.storage_range = storage_range
        self.initial_storage_mean = initial_storage_mean
        self.initial_storage_std = initial_storage_std
        self.charge_efficiency = charge_efficiency
        self.discharge_efficiency = discharge_efficiency
        self.max_power = max_power
        self.current_storage = None
        self.rescale_spaces = rescale_spaces
        self.simulation_step = 0
        self.max_episode_steps = max_episode_steps
        self.control_interval_in_hr = control_timedelta.seconds / 3600.0
        self._obs_labels = ["stage_of_charge"]      # sic (reference :51)
        self._observation_space = spaces.Box(shape=(1,), low=storage_range[0],
                                             high=storage_range[1], dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(shape=(1,), low=-1.0, high=1.0, dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)

    def draw_initial_storage(self, size=None):
        """The reference's host-side draw (:82-84): truncnorm(-1, 1) on SciPy's global RNG."""
        from scipy.stats import truncnorm
        return truncnorm(-1, 1).rvs(size=size) * self.initial_storage_std + self.initial_storage_mean

    def _reset_result(self, obs):
        return obs, {"state_of_charge": None}      # (obs, meta) like the reference (:97)

    def _terminal_after(self):
        return self.max_episode_steps - 1           # simulation_step + 1 == max (:155-157, :181)

    def _meta(self, ctx) -> dict:
        soc = ctx.sd(self._slot["sd"][0])
        return {"state_of_charge": ctx.vector([soc])}         # raw_obs, a 1-vector (:178)

    def _emit(self, b, agent_index, standalone):
        dpar = [self.storage_range[0], self.storage_range[1], self.charge_efficiency,
                self.discharge_efficiency, self.max_power, self.control_interval_in_hr,
                self.initial_storage_mean,
                1.0 / (self.storage_range[1] - self.storage_range[0]),
                1.0 / self.discharge_efficiency, 1.0 / self.control_interval_in_hr]
        b.add_component(self, N.STORAGE, agent_index,
                        flags=N.F_RESCALE if self.rescale_spaces else 0,
                        dpar=dpar, ipar=[b.next_storage_ordinal(self)], sd_rows=1)
