"""Dict <-> list adapter (mirrors gridworld/multiagent_list_interface_env.py:8-111), the
interface the reference's MADDPG example trains through
(examples/marl/openai/train.py:165-188).

For ``num_envs == 1`` the methods return per-agent lists of 1-D arrays exactly like the
reference.  For batches, ``reset_batch`` / ``step_batch`` return per-agent *views* of the
env's device tensors (``[agent_obs_dim, E]``), no copies: the flat row layout of the
device buffers already is the list interface's concatenation order.
"""
from collections import OrderedDict

import numpy as np

from powergridworld_b200 import spaces


class MultiAgentListInterfaceEnv:

    def __init__(self, multi_agent_env_cls, env_config, **env_kwargs):
        self.ma_env = multi_agent_env_cls(**env_config, **env_kwargs)
        self.n = len(self.ma_env.agents)
        self.nested_sequence = self.get_nested_sequence(env_config['agents'])
        self.observation_space, self.action_space = [], []
        self._obs_rows, self._act_rows = [], []
        for k, v in self.nested_sequence.items():
            obs_len = sum(self.ma_env.observation_space[k][c].shape[0] for c in v)
            act_len = sum(self.ma_env.action_space[k][c].shape[0] for c in v)
            self.observation_space.append(
                spaces.Box(shape=(obs_len,), low=-1.0, high=1.0, dtype=np.float64))
            self.action_space.append(
                spaces.Box(shape=(act_len,), low=-1.0, high=1.0, dtype=np.float64))
            o = self.ma_env.obs_slices[k]
            a = self.ma_env.act_slices[k]
            first_o = min(s[0] for s in o.values())
            first_a = min(s[0] for s in a.values())
            self._obs_rows.append((first_o, obs_len))
            self._act_rows.append((first_a, act_len))

    @staticmethod
    def get_nested_sequence(agent_config):
        seq = OrderedDict()
        for item in agent_config:
            seq[item['name']] = [x['name'] for x in item['config']['components']]
        return seq

    # ---- reference API (num_envs == 1)
    def reset(self, **kw):
        return self.convert_to_list_obs(self.ma_env.reset(**kw))

    def step(self, action):
        next_obs, reward, done, info = self.ma_env.step(self.convert_from_list_act(action))
        return (self.convert_to_list_obs(next_obs),
                [reward[k] for k in self.nested_sequence.keys()],
                [done[k] for k in self.nested_sequence.keys()], info)

    def convert_to_list_obs(self, obs):
        return [np.concatenate([obs[k][x] for x in v]) for k, v in self.nested_sequence.items()]

    def convert_from_list_act(self, action):
        converted = {}
        for idx, (k, v) in enumerate(self.nested_sequence.items()):
            agent_action, start = {}, 0
            for component in v:
                n = self.ma_env.action_space[k][component].shape[0]
                agent_action[component] = action[idx][start:start + n]
                start += n
            converted[k] = agent_action
        return converted

    # ---- batched, zero-copy
    def reset_batch(self, init_storage=None):
        obs = self.ma_env.reset_batch(init_storage)
        return [obs[o:o + n] for o, n in self._obs_rows]

    def step_batch(self, actions):
        """``actions``: the env's flat ``[act_dim, E]`` tensor (per-agent views of it can be
        obtained once with ``action_views``).  Returns (obs list, reward list, done, all_done)."""
        obs, rew, done, all_done = self.ma_env.step_batch(actions)
        return ([obs[o:o + n] for o, n in self._obs_rows], [rew[i] for i in range(self.n)],
                done, all_done)

    def action_views(self, actions):
        return [actions[o:o + n] for o, n in self._act_rows]
