"""Home-Steward battery description
(gridworld/agents/energy_storage/energy_storage_env_hs.py:10-273).
Dynamics: csrc/component_math.cuh hs_storage_step / hs_storage_reset."""
import numpy as np
import pandas as pd

from powergridworld_b200 import _native as N
from powergridworld_b200 import spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space


class HSEnergyStorageEnv(ComponentEnv):

    def __init__(self, name: str = None, storage_range: tuple = (3.0, 50.0),
                 initial_storage_mean: float = 30.0, initial_storage_std: float = 5.0,
                 charge_efficiency: float = 0.95, discharge_efficiency: float = 0.9,
                 max_power: float = 15.0, max_episode_steps: int = 288,
                 control_timedelta: pd.Timedelta = pd.Timedelta(300, "s"),
                 rescale_spaces: bool = True, initial_storage_cost: float = 0.0,
                 max_storage_cost: float = 0.55, **kwargs):
        super().__init__(name=name)
        self.initial_storage_cost = initial_storage_cost
        self.storage_range = storage_range
        self.initial_storage_mean = initial_storage_mean
        self.initial_storage_std = initial_storage_std
        self.charge_efficiency = charge_efficiency
        self.discharge_efficiency = discharge_efficiency
        self.max_power = max_power
        self.rescale_spaces = rescale_spaces
        self.max_storage_cost = max_storage_cost
        self.max_episode_steps = max_episode_steps
        self.control_interval_in_hr = control_timedelta.seconds / 3600.0
        self._obs_labels = ["stage_of_charge", "cost"]
        self._observation_space = spaces.Box(
            shape=(2,), low=np.array([storage_range[0], 0.00]),
            high=np.array([storage_range[1], max_storage_cost]), dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(shape=(1,), low=-1.0, high=1.0, dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)

    def draw_initial_storage(self, size=None):
        """The reference's host-side draw (:88-91): truncnorm(-1, 1) on SciPy's global RNG."""
        from scipy.stats import truncnorm
        return truncnorm(-1, 1).rvs(size=size) * self.initial_storage_std + self.initial_storage_mean

    def _terminal_after(self):
        return self.max_episode_steps               # simulation_step == max after the step (:270)

    def _emit(self, b, agent_index, standalone):
        if standalone:
            raise NotImplementedError("HS components are stepped inside an HSMultiComponentEnv")
        dpar = [self.storage_range[0], self.storage_range[1], self.charge_efficiency,
                self.discharge_efficiency, self.max_power, self.control_interval_in_hr,
                self.initial_storage_mean, self.initial_storage_cost, self.max_storage_cost,
                1.0 / (self.storage_range[1] - self.storage_range[0]), 1.0 / self.max_storage_cost,
                1.0 / self.discharge_efficiency, 1.0 / self.control_interval_in_hr,
                1.0 / self.charge_efficiency]
        b.add_component(self, N.HS_STORAGE, agent_index,
                        flags=(N.F_RESCALE if self.rescale_spaces else 0)
                        | (N.F_TELEMETRY if getattr(self, "_telemetry", False) else 0),
                        dpar=dpar, ipar=[b.next_storage_ordinal(self)],
                        sd_rows=2 + (N.HS_TEL_ROWS if getattr(self, "_telemetry", False) else 0))
