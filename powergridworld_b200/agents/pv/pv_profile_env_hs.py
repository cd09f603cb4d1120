"""Home-Steward PV description (gridworld/agents/pv/pv_profile_env_hs.py:15-169).
Dynamics: csrc/component_math.cuh hs_pv_step."""
import json
import os

import numpy as np

from powergridworld_b200 import _native as N
from powergridworld_b200 import spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space

_HS_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "data",
                       "hs_env_config.json")


def packaged_hs_config() -> dict:
    """The numbers of gridworld/scenarios/data/env_config.json (tools/import_reference_data.py)."""
    with open(_HS_CFG) as fh:
        return json.load(fh)


class HSPVEnv(ComponentEnv):

    def __init__(self, name: str, profile_csv: str = None, profile_path: str = None,
                 profile_data: list = [], scaling_factor: float = 1., rescale_spaces: bool = True,
                 grid_aware: bool = False, max_episode_steps: int = None,
                 minutes_per_step: int = 5, **kwargs):
        super().__init__(name=name, **kwargs)
        if grid_aware:
            raise NotImplementedError("grid_aware HSPVEnv: the house env has no feeder")
        self.scaling_factor, self.rescale_spaces = scaling_factor, rescale_spaces
        self.grid_aware, self.minutes_per_step = grid_aware, minutes_per_step
        if len(profile_data):
            data = np.array(profile_data)
        elif profile_path is not None:
            import pandas as pd
            data = pd.read_csv(profile_path).values[:, 0].squeeze()
        else:                                   # profiles/pv_profile_hs.csv = the packaged house
            cfg = [c for c in packaged_hs_config()["components"] if c["cls"] == "HSPVEnv"][0]
            data = np.array(cfg["config"]["profile_data"])
        self.data = [self.scaling_factor * i for i in data]                       # :74
        self.episode_length = len(self.data)
        if max_episode_steps is not None:
            self.episode_length = min(max_episode_steps, self.episode_length)
        self._obs_labels = ["real_power"]
        self._observation_space = spaces.Box(shape=(1,), low=np.array([-np.max(self.data)]),
                                             high=np.array([0.]), dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(shape=(1,), low=0.98, high=1., dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)

    def _terminal_after(self):
        return self.episode_length                  # index == episode_length after the step (:134)

    def _emit(self, b, agent_index, standalone):
        if standalone:
            raise NotImplementedError("HS components are stepped inside an HSMultiComponentEnv")
        lo, hi = self._observation_space.low, self._observation_space.high
        data = self.data
        last = len(data) - 1
        # event 0 (reset) shows row 0; step t acts on row t
        b.add_component(self, N.HS_PV, agent_index,
                        flags=(N.F_RESCALE if self.rescale_spaces else 0)
                        | (N.F_TELEMETRY if getattr(self, "_telemetry", False) else 0),
                        dpar=[lo[0], hi[0], 1.0 / (hi[0] - lo[0])],
                        dtab_width=1, dtab_fn=lambda r: [data[min(max(r - 1, 0), last)]],
                        sd_rows=N.HS_TEL_ROWS if getattr(self, "_telemetry", False) else 0)
