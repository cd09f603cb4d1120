# Default observation configuration of the building (gridworld/agents/buildings/defaults.py:2-10)
obs_config = {
    "zone_upper_viol": (-10., 10.),
    "zone_lower_viol": (-10., 10.),
    "comfort_lower": (20., 25.),
    "comfort_upper": (25., 30),
    "outdoor_temp": (0., 56.),
    "p_consumed": (0., 100.),
    "time_of_day": (0., 1.)
}
