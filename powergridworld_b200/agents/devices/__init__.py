from .devices_env_hs import HSDevicesEnv
