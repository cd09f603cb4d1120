from .ev_charging_env import EVChargingEnv
from .ev_charging_env_hs import HSEVChargingEnv
