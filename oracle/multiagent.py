"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the multi-agent orchestrator.

Restates gridworld/multiagent_env.py:20-230 (MultiAgentEnv) and the two
grid-level reward hooks that sit on the hot path:
  * examples/marl/openai/train.py:37-88  CoordinatedMultiBuildingControlEnv
  * gridworld/scenarios/heterogeneous.py:46-52 (see oracle.components.GridAwarePVEnv)

One object = one env instance.  The power-flow plugin is whatever
``pf_config["cls"]`` names (oracle.powerflow.ZBusOracleSolver in this repo).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class PowerFlowSolver:
    """gridworld/distribution_system/powerflow.py:7-51 (plugin protocol)."""

    def __init__(self, config=None, **kwargs):
        pass

    def calculate_power_flow(self, p_controllable_consumed=None,
                             q_controllable_consumed=None, **kwargs):
        raise NotImplementedError

    def get_bus_voltages(self):
        raise NotImplementedError

    def get_bus_voltage_by_name(self, name):
        raise NotImplementedError


class MultiAgentEnv:
    """multiagent_env.py:20-230."""

    def __init__(self, common_config=None, pf_config=None, agents=None,
                 max_episode_steps=None, rescale_spaces=True, **kwargs):
        assert agents is not None and len(agents) > 0, "need at least one agent!"
        self.common_config = common_config
        self.rescale_spaces = rescale_spaces
        self.start_time = pd.Timestamp(common_config["start_time"])
        self.end_time = pd.Timestamp(common_config["end_time"])
        self.control_timedelta = common_config["control_timedelta"]
        self.max_episode_steps = max_episode_steps if max_episode_steps is not None else np.inf
        self.episode_step = None
        self.time = None
        self.history = None
        self.voltages = None
        self.agents = []
        for a in agents:
            cfg = {k: v for k, v in a["config"].items() if k != "name"}
            self.agents.append(a["cls"](name=a["name"], **cfg, **common_config))   # :69
        self.agent_name_bus_map = {a["name"]: a["bus"] for a in agents}
        names = [a.name for a in self.agents]
        assert len(set(names)) == len(agents), "all agents need unique names"
        self.agent_names = names
        self.pf_solver = pf_config["cls"](**pf_config["config"])
        self.observation_space = {a.name: a.observation_space for a in self.agents}
        self.action_space = {a.name: a.action_space for a in self.agents}

    # :90-115
    def get_external_obs_vars(self, agent):
        kw = {}
        if "bus_voltage" in agent.obs_labels:
            kw["bus_voltage"] = self.pf_solver.get_bus_voltage_by_name(
                self.agent_name_bus_map[agent.name])
        if "max_voltage" in agent.obs_labels:
            kw["max_voltage"] = max(list(self.voltages.values()))
        if "min_voltage" in agent.obs_labels:
            kw["min_voltage"] = min(list(self.voltages.values()))
        return kw

    # :125-148.  ``init_storage`` is an oracle-only hook: {agent: soc} or
    # {agent: {component: soc}} overriding the host RNG draw so that parity runs
    # can feed identical initial SOC to every implementation.
    def reset(self, init_storage=None):
        self.episode_step = 0
        self.time = self.start_time
        self.history = {"timestamp": [], "voltage": [], "agent_power_p": []}
        self.pf_solver.calculate_power_flow(current_time=self.time)
        self.voltages = self.pf_solver.get_bus_voltages()
        for agent in self.agents:
            kw = self.get_external_obs_vars(agent)
            agent.reset(**kw)
            if init_storage is not None and agent.name in init_storage:
                _override_soc(agent, init_storage[agent.name])
        return self.get_obs()

    def get_obs(self):
        obs = {}
        for agent in self.agents:
            kw = self.get_external_obs_vars(agent)
            obs[agent.name], _ = agent.get_obs(**kw)
        return obs

    # :151-212
    def step(self, action):
        self.episode_step += 1
        self.time += self.control_timedelta
        obs, rew, done, meta = {}, {}, {}, {}
        load_p, load_q, agent_power_p = {}, {}, []
        for agent in self.agents:
            kw = self.get_external_obs_vars(agent)          # voltages of the PREVIOUS solve
            obs[agent.name], rew[agent.name], done[agent.name], meta[agent.name] = \
                agent.step(action=action[agent.name], **kw)
            bus = self.agent_name_bus_map[agent.name]
            p, q = agent.real_power, agent.reactive_power
            agent_power_p.append(p)
            if bus in load_p:
                load_p[bus] += p
                load_q[bus] += q
            else:
                load_p[bus] = p
                load_q[bus] = q
        self.pf_solver.calculate_power_flow(
            current_time=self.time, p_controllable_consumed=load_p,
            q_controllable_consumed=load_q)
        self.voltages = self.pf_solver.get_bus_voltages()
        self.history["timestamp"].append(self.time)
        self.history["voltage"].append(dict(self.voltages))
        self.history["agent_power_p"].append(agent_power_p)
        any_done = bool(np.any(list(done.values())))
        finished = any_done or (self.episode_step == self.max_episode_steps - 1) \
            or (self.time >= self.end_time)
        dones = {a.name: finished for a in self.agents}
        dones["__all__"] = finished
        return obs, self.reward_transform(rew), dones, self.meta_transform(meta)

    def reward_transform(self, rew):
        return rew

    def meta_transform(self, meta):
        return meta

    @property
    def agent_dict(self):
        return {a.name: a for a in self.agents}


def _override_soc(agent, value):
    from oracle.components import EnergyStorageEnv
    if isinstance(agent, EnergyStorageEnv):
        agent.soc = float(np.clip(float(value), agent.lo, agent.hi))
        return
    for e in getattr(agent, "envs", []):
        if isinstance(e, EnergyStorageEnv):
            v = value[e.name] if isinstance(value, dict) else value
            e.soc = float(np.clip(float(v), e.lo, e.hi))


class CoordinatedMultiBuildingControlEnv(MultiAgentEnv):
    """examples/marl/openai/train.py:37-88: shared voltage-violation penalty
    computed from the FRESH solve at the common load node."""

    VOLTAGE_LIMITS = [0.95, 1.05]
    VV_UNIT_PENALTY = 1e4

    def reward_transform(self, rew):
        penalty = self.get_voltage_violation() * self.VV_UNIT_PENALTY
        n = len(rew)
        for k in rew.keys():
            rew[k] -= penalty / n
        return rew

    def meta_transform(self, meta):
        meta.update({"voltage_violation": self.get_voltage_violation()})
        return meta

    def get_voltage_violation(self):
        buses = set(self.agent_name_bus_map.values())
        assert len(buses) == 1, "all buildings should be on the same bus"
        v = self.pf_solver.get_bus_voltage_by_name(list(buses)[0])
        return max([0.0, self.VOLTAGE_LIMITS[0] - v, v - self.VOLTAGE_LIMITS[1]])
