from .ev_charging_env import EVChargingEnv
