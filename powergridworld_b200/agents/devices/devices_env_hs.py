"""Home-Steward "other devices" description (gridworld/agents/devices/devices_env_hs.py:14-205).
Dynamics: csrc/component_math.cuh hs_devices_step."""
import numpy as np

from powergridworld_b200 import _native as N
from powergridworld_b200 import spaces
from powergridworld_b200.base import ComponentEnv
from powergridworld_b200.utils import maybe_rescale_box_space


class HSDevicesEnv(ComponentEnv):

    def __init__(self, name: str, profile_csv: str = None, profile_path: str = None,
                 profile_data: dict = {}, scaling_factor: float = 1., rescale_spaces: bool = True,
                 max_episode_steps: int = None, minutes_per_step: int = 5, **kwargs):
        super().__init__(name=name, **kwargs)
        self.scaling_factor, self.rescale_spaces = scaling_factor, rescale_spaces
        self.minutes_per_step = minutes_per_step
        if profile_data:
            labels = list(profile_data.keys())
            frame = np.array([v for v in profile_data.values()], dtype=np.float64).T   # :55-57
        elif profile_path is not None:
            import pandas as pd
            df = pd.read_csv(profile_path)
            labels, frame = list(df.columns), df.values.astype(np.float64)
        else:                                   # data/devices_profile_hs.csv = the packaged house
            from powergridworld_b200.agents.pv.pv_profile_env_hs import packaged_hs_config
            cfg = [c for c in packaged_hs_config()["components"] if c["cls"] == "HSDevicesEnv"][0]
            pdict = cfg["config"]["profile_data"]
            labels = list(pdict.keys())
            frame = np.array([v for v in pdict.values()], dtype=np.float64).T
        if frame.ndim != 2 or frame.shape[1] < 2:
            raise NotImplementedError("single-column device tables (the reference squeezes them "
                                      "into a vector, :70) are not supported")
        # observation = the scaled copy; demand = the frame itself (:70-71, :118, :165: under
        # pandas copy-on-write `.values` is a copy, so the frame stays unscaled)
        self._frame = frame
        self.data = frame * scaling_factor
        self.episode_length = len(self.data)
        if max_episode_steps is not None:
            self.episode_length = min(max_episode_steps, self.episode_length)
        self._obs_labels = labels
        high = frame.max(axis=0)                                                   # :80-82
        self._observation_space = spaces.Box(shape=(len(labels),), low=np.zeros(len(labels)),
                                             high=high, dtype=np.float64)
        self.observation_space = maybe_rescale_box_space(self._observation_space, rescale_spaces)
        self._action_space = spaces.Box(shape=(1,), low=0.99, high=1., dtype=np.float64)
        self.action_space = maybe_rescale_box_space(self._action_space, rescale_spaces)

    def _terminal_after(self):
        return self.episode_length                  # index == episode_length after the step (:125)

    def _emit(self, b, agent_index, standalone):
        if standalone:
            raise NotImplementedError("HS components are stepped inside an HSMultiComponentEnv")
        k = self.data.shape[1]
        data, frame, last = self.data, self._frame, len(self.data) - 1
        row = lambda r: min(max(r - 1, 0), last)
        b.add_component(self, N.HS_DEVICES, agent_index,
                        flags=(N.F_RESCALE if self.rescale_spaces else 0)
                        | (N.F_TELEMETRY if getattr(self, "_telemetry", False) else 0),
                        dpar=[self.minutes_per_step / 60.0] + list(self._observation_space.high)
                        + list(1.0 / self._observation_space.high),
                        ipar=[k], dtab_width=2 * k,
                        dtab_fn=lambda r: list(data[row(r)]) + list(frame[row(r)]),
                        sd_rows=N.HS_TEL_ROWS if getattr(self, "_telemetry", False) else 0)
