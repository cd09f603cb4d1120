from .pv_profile_env import PVEnv, GridAwarePVEnv
from .pv_profile_env_hs import HSPVEnv
