"""powergridworld_b200 -- B200-native batched simulator for PowerGridworld's
``MultiAgentEnv.step`` hot path.  Same plugin surface as the reference
(``gridworld``): ComponentEnv / MultiComponentEnv agents, a PowerFlowSolver plugin,
scenario config dicts, per-agent obs / action / reward dicts -- executed by
hand-written sm_100a CUDA kernels behind the C ABI of ``include/pgw.h``."""
__version__ = "0.1.0"

from .base import ComponentEnv, MultiComponentEnv
from .base_hs import HSMultiComponentEnv
from .multiagent_env import CoordinatedMultiBuildingControlEnv, MultiAgentEnv
from .multiagent_list_interface_env import MultiAgentListInterfaceEnv
