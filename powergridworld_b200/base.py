"""Agent plugin protocol (mirrors gridworld/base.py:12-182).

The classes keep the reference's names, constructor keywords and attributes
(``name``, ``observation_space``, ``action_space``, ``_observation_space``,
``_action_space``, ``rescale_spaces``, ``obs_labels``, ``envs``, ``env_dict`` ...),
but they do not compute anything themselves: a component is a *description* that
``MultiAgentEnv`` compiles into the structure-of-arrays tables of the CUDA
kernels (``_emit``).  There is deliberately no Python implementation of the
dynamics in this package -- no CPU fallback.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, List

from powergridworld_b200 import spaces

_GPU_ONLY = ("{cls} is stepped on the GPU by powergridworld_b200.MultiAgentEnv "
             "(wrap it as an agent; pf_config=None gives a component-only env); "
             "this package has no CPU implementation of the dynamics.")


class ComponentEnv(ABC):
    """gridworld/base.py:12-71."""

    def __init__(self, name: str = None, **kwargs):
        self.name = name
        self._real_power = 0.
        self._reactive_power = 0.
        self._obs_labels: List[str] = []

    # ---- the reference's per-object protocol (gridworld/base.py:29-49), as used by its
    # tests/agents/*.py and notebooks: a component stepped on its own.  Served by a private
    # one-env, one-agent device handle without a feeder -- still the CUDA path, there is no
    # Python implementation of the dynamics.  Grid variables cannot be injected this way.
    _standalone = None

    def _runner(self):
        if self._standalone is None:
            import pandas as pd
            from powergridworld_b200.multiagent_env import MultiAgentEnv
            if any(k in self.obs_labels for k in ("bus_voltage", "min_voltage", "max_voltage")):
                raise NotImplementedError(_GPU_ONLY.format(cls=type(self).__name__))
            me = self
            self.name = self.name if self.name is not None else type(self).__name__
            self._standalone = MultiAgentEnv(
                common_config={"start_time": "01-01-2021 00:00:00",
                               "end_time": "01-01-2031 00:00:00",
                               "control_timedelta": pd.Timedelta(300, "s")},
                pf_config=None,
                agents=[{"name": self.name, "bus": None, "cls": lambda name, **kw: me,
                         "config": {}}])
            self._last = None
        return self._standalone

    def reset(self, **kwargs):
        r = self._runner()
        obs = r.reset(init_storage=[kwargs["init_storage"]]
                      if kwargs.get("init_storage") is not None and r.num_storage == 1 else None)
        self._last = (obs[self.name], 0.0, False, {})
        return self._reset_result(obs[self.name])

    def _reset_result(self, obs):
        return obs

    def step(self, action, **kwargs):
        r = self._runner()
        if r._needs_reset:
            raise RuntimeError("call reset before step")
        obs, rew, dones, meta = r.step({self.name: action})
        self._last = (obs[self.name], rew[self.name], dones["__all__"], meta[self.name])
        return self._last

    def step_reward(self, **kwargs):
        if self._last is None:
            raise RuntimeError("no step has been taken")
        return self._last[1], {}

    def get_obs(self, **kwargs):
        if self._last is None:
            raise RuntimeError("call reset first")
        return self._last[0], {}

    @property
    def real_power(self) -> float:
        """kW, + load / - generation; refreshed by MultiAgentEnv after each step (E == 1)."""
        return self._real_power

    @property
    def reactive_power(self) -> float:
        return self._reactive_power          # always 0 in the reference (base.py:24)

    @property
    def obs_labels(self) -> list:
        return self._obs_labels

    # ---- meta of a step, rebuilt from the device state (the 4th return value of
    # gridworld's step(); ctx: see MultiAgentEnv._meta_context)
    def _meta(self, ctx) -> dict:
        return {}

    # ---- spec compiler hooks
    @abstractmethod
    def _emit(self, builder, agent_index: int, standalone: bool) -> None:
        """Append this component's descriptor / parameters / event blocks to ``builder``."""

    @abstractmethod
    def _terminal_after(self) -> float:
        """Number of env steps after which the component's is_terminal() turns true."""

    @property
    def _act_dim(self) -> int:
        return int(self.action_space.shape[0])

    @property
    def _obs_dim(self) -> int:
        return int(self.observation_space.shape[0])


class MultiComponentEnv(ComponentEnv):
    """gridworld/base.py:74-182: one agent made of an ordered list of components."""

    def __init__(self, name: str = None, components: List[dict] = None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.envs = [c["cls"](name=c["name"], **c["config"]) for c in components]
        for e in self.envs:
            if isinstance(e, MultiComponentEnv):
                raise TypeError("nested MultiComponentEnv is not supported")
        self.observation_space = spaces.Dict({e.name: e.observation_space for e in self.envs})
        self.action_space = spaces.Dict({e.name: e.action_space for e in self.envs})
        self._obs_labels_dict = {e.name: e.obs_labels for e in self.envs}
        labels = []
        for e in self.envs:
            labels += e.obs_labels
        self._obs_labels = list(set(labels))

    def _reset_result(self, obs):
        return obs, {e.name: {} for e in self.envs}          # (obs, meta), base.py:108-111

    def _emit(self, builder, agent_index, standalone):
        for e in self.envs:
            e._emit(builder, agent_index, standalone=False)

    def _terminal_after(self):
        return min(e._terminal_after() for e in self.envs)

    def _meta(self, ctx) -> dict:
        return {e.name: e._meta(ctx) for e in self.envs}     # base.py:127-130

    @property
    def obs_labels_dict(self) -> Dict[str, list]:
        return self._obs_labels_dict

    @property
    def env_dict(self) -> Dict[str, ComponentEnv]:
        return {e.name: e for e in self.envs}
