from .pv_profile_env import PVEnv, GridAwarePVEnv
