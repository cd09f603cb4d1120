"""Scenario catalog: the reference's shipped scenarios and BASELINE.json's benchmark
configurations as config-dict builders (used by bench.py, __graft_entry__.smoke, the golden
generator and the parity tests).

Each builder takes a *namespace* ``ns`` providing the plugin classes
(``MultiComponentEnv, FiveZoneROMThermalEnergyEnv, PVEnv, GridAwarePVEnv,
EnergyStorageEnv, EVChargingEnv``) so the very same scenario can be built
against the reference (tests/golden/make_golden.py, authoring container only),
the CPU oracle and the product.  They restate the reference's shipped scenarios:

  buildings_scenario      gridworld/scenarios/buildings.py:11-72 as configured by
                          examples/marl/openai/train.py:165-188  (BASELINE C0/C1)
  heterogeneous_scenario  gridworld/scenarios/heterogeneous.py:13-112
  ev_pv_storage_scenario  BASELINE C2 (component-only)
  randomized_ev_scenario  charging stations that re-draw their roster on every reset
  test_* fixtures         tests/conftest.py:99-148, tests/agents/conftest.py
"""
import pandas as pd

IEEE13 = {
    "feeder_file": "ieee_13_dss/IEEE13Nodeckt.dss",
    "loadshape_file": "ieee_13_dss/annual_hourly_load_profile.csv",
}


def _pf(pf_cls, rescale):
    return {"cls": pf_cls, "config": dict(IEEE13, system_load_rescale_factor=rescale)}


def buildings_scenario(ns, pf_cls, system_load_rescale_factor=1.2, num_buildings=3):
    components = [
        {"name": "building", "cls": ns.FiveZoneROMThermalEnergyEnv, "config": {}},
        {"name": "pv", "cls": ns.PVEnv,
         "config": {"profile_csv": "pv_profile.csv", "scaling_factor": 40.}},
        {"name": "storage", "cls": ns.EnergyStorageEnv,
         "config": {"max_power": 15., "storage_range": (3., 50.)}},
    ]
    return {
        "common_config": {"start_time": "08-12-2021 00:00:00",
                          "end_time": "08-13-2021 00:00:00",
                          "control_timedelta": pd.Timedelta(300, "s")},
        "pf_config": _pf(pf_cls, system_load_rescale_factor),
        "agents": [{"name": f"building-{i}", "bus": "675c", "cls": ns.MultiComponentEnv,
                    "config": {"components": components}} for i in range(num_buildings)],
    }


def heterogeneous_scenario(ns, pf_cls, system_load_rescale_factor=0.65, rescale_spaces=True,
                           max_episode_steps=None):
    bcomp = [
        {"name": "building", "cls": ns.FiveZoneROMThermalEnergyEnv,
         "config": {"reward_structure": {"alpha": 0.0}, "rescale_spaces": rescale_spaces}},
        {"name": "pv", "cls": ns.PVEnv,
         "config": {"profile_csv": "off-peak.csv", "scaling_factor": 40.,
                    "rescale_spaces": rescale_spaces}},
        {"name": "storage", "cls": ns.EnergyStorageEnv,
         "config": {"max_power": 20., "storage_range": (3., 250.),
                    "rescale_spaces": rescale_spaces}},
    ]
    cfg = {
        "common_config": {"start_time": "08-12-2020 00:00:00",
                          "end_time": "08-13-2020 00:00:00",
                          "control_timedelta": pd.Timedelta(300, "s")},
        "pf_config": _pf(pf_cls, system_load_rescale_factor),
        "agents": [
            {"name": "building", "bus": "675c", "cls": ns.MultiComponentEnv,
             "config": {"components": bcomp}},
            {"name": "pv", "bus": "675c", "cls": ns.GridAwarePVEnv,
             "config": {"profile_csv": "constant.csv", "scaling_factor": 400.,
                        "rescale_spaces": rescale_spaces, "grid_aware": True}},
            {"name": "ev-charging", "bus": "675c", "cls": ns.EVChargingEnv,
             "config": {"num_vehicles": 25, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
                        "peak_threshold": 200., "vehicle_multiplier": 40.,
                        "rescale_spaces": rescale_spaces}},
        ],
    }
    if max_episode_steps is not None:
        cfg["max_episode_steps"] = max_episode_steps
    return cfg


def ev_pv_storage_scenario(ns, pf_cls=None, rescale_spaces=True):
    """BASELINE C2: one composite-free env = {EV(100), PV x10, storage}; no feeder."""
    cfg = {
        "common_config": {"start_time": "08-12-2020 00:00:00",
                          "end_time": "08-13-2020 00:00:00",
                          "control_timedelta": pd.Timedelta(300, "s")},
        "pf_config": {"cls": pf_cls, "config": {}},
        "agents": [
            {"name": "ev", "bus": "675c", "cls": ns.EVChargingEnv,
             "config": {"num_vehicles": 100, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
                        "peak_threshold": 250., "vehicle_multiplier": 5.,
                        "rescale_spaces": rescale_spaces}},
            {"name": "pv", "bus": "675c", "cls": ns.PVEnv,
             "config": {"profile_csv": "pv_profile.csv", "scaling_factor": 10.,
                        "rescale_spaces": rescale_spaces}},
            {"name": "storage", "bus": "675c", "cls": ns.EnergyStorageEnv,
             "config": {"rescale_spaces": rescale_spaces}},
        ],
    }
    return cfg


TIME_BASE_VARIANTS = {
    # control_timedelta [s], start, end, max_episode_steps: the same heterogeneous scenario on
    # other clocks (the storage's dt, the feeder's hourly load shape, the episode length and who
    # ends the episode all depend on them)
    "dt600_day": (600, "08-12-2020 06:00:00", "08-12-2020 18:00:00", None),
    "dt60_night": (60, "08-12-2020 00:00:00", "08-12-2020 03:00:00", None),
    "dt900_max50": (900, "08-12-2020 00:00:00", "08-13-2020 00:00:00", 50),
    "dt300_from_noon": (300, "08-12-2020 12:00:00", "08-13-2020 12:00:00", None),
}


def time_base_scenario(ns, pf_cls, variant):
    dt, start, end, mes = TIME_BASE_VARIANTS[variant]
    cfg = heterogeneous_scenario(ns, pf_cls, 0.65)
    cfg["common_config"] = {"start_time": start, "end_time": end,
                            "control_timedelta": pd.Timedelta(dt, "s")}
    if mes:
        cfg["max_episode_steps"] = mes
    return cfg


def randomized_ev_scenario(ns, pf_cls):
    """EVChargingEnv(randomize=True) (ev_charging_env.py:154-157) standalone and inside a
    MultiComponentEnv, between storages so that the order of the host RNG draws of a reset
    (storage SOC, roster, storage SOC, roster) is part of what is compared."""
    ev = lambda n, mult, rescale: {"num_vehicles": n, "minutes_per_step": 5,
                                   "max_charge_rate_kw": 7., "peak_threshold": 60.,
                                   "vehicle_multiplier": mult, "rescale_spaces": rescale,
                                   "randomize": True}
    depot = [
        {"name": "storage", "cls": ns.EnergyStorageEnv, "config": {"max_power": 20.}},
        {"name": "chargers", "cls": ns.EVChargingEnv, "config": ev(20, 3., True)},
    ]
    return {
        "common_config": {"start_time": "08-12-2020 00:00:00",
                          "end_time": "08-13-2020 00:00:00",
                          "control_timedelta": pd.Timedelta(300, "s")},
        "pf_config": _pf(pf_cls, 0.7),
        "agents": [
            {"name": "storage", "bus": "675c", "cls": ns.EnergyStorageEnv, "config": {}},
            {"name": "depot", "bus": "634a", "cls": ns.MultiComponentEnv,
             "config": {"components": depot}},
            {"name": "ev", "bus": "675a", "cls": ns.EVChargingEnv, "config": ev(45, 2., False)},
        ],
    }


def test_multicomponent_components(ns):
    """tests/conftest.py:113-148 (unscaled spaces, 6-dim building obs)."""
    return [
        {"name": "building", "cls": ns.FiveZoneROMThermalEnergyEnv,
         "config": {"start_time": "08-12-2020 00:00:00", "end_time": "08-13-2020 00:00:00",
                    "rescale_spaces": False,
                    "obs_config": {"zone_temp": (18, 34), "p_consumed": (-100, 100)}}},
        {"name": "pv", "cls": ns.PVEnv,
         "config": {"profile_csv": "pv_profile.csv", "scaling_factor": 10.,
                    "rescale_spaces": False}},
        {"name": "storage", "cls": ns.EnergyStorageEnv, "config": {"rescale_spaces": False}},
    ]


def test_heterogeneous_scenario(ns, pf_cls):
    """tests/test_multiagent_env.py:66-107 (load factor 0.7)."""
    ev_cfg = {"num_vehicles": 100, "minutes_per_step": 5, "max_charge_rate_kw": 7.,
              "peak_threshold": 250., "vehicle_multiplier": 5., "rescale_spaces": False}
    return {
        "common_config": {"start_time": "08-12-2020 00:00:00",
                          "end_time": "08-13-2020 00:00:00",
                          "control_timedelta": pd.Timedelta(300, "s")},
        "pf_config": _pf(pf_cls, 0.7),
        "agents": [
            {"name": "building", "bus": "675c", "cls": ns.MultiComponentEnv,
             "config": {"components": test_multicomponent_components(ns)}},
            {"name": "ev-charging", "bus": "675c", "cls": ns.EVChargingEnv, "config": ev_cfg},
            {"name": "pv", "bus": "675c", "cls": ns.PVEnv,
             "config": {"profile_csv": "pv_profile.csv", "scaling_factor": 400.}},
        ],
    }


# keep pytest from collecting the builders above as tests
test_multicomponent_components.__test__ = False
test_heterogeneous_scenario.__test__ = False


def synthetic123_load_names():
    """Load names of the authored 123-bus-class feeder, in definition order."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "feeders",
                        "synthetic123.dss")
    with open(path) as fh:
        return [m.group(1).lower() for m in re.finditer(r"^New Load\.(\S+)", fh.read(), re.M)]


def der123_scenario(ns, pf_cls, n_agents=100, system_load_rescale_factor=0.9):
    """BASELINE C3: 123-bus-class feeder with ~100 heterogeneous DER agents, one per load
    (PV 40 %, storage 30 %, EV station 10 %, building+PV+storage composite 20 %)."""
    loads = synthetic123_load_names()
    agents = []
    for i in range(n_agents):
        kind, bus = i % 10, loads[i % len(loads)]
        if kind <= 3:
            agents.append({"name": f"pv-{i}", "bus": bus, "cls": ns.PVEnv,
                           "config": {"profile_csv": ["pv_profile.csv", "off-peak.csv"][i % 2],
                                      "scaling_factor": 20. + i}})
        elif kind <= 6:
            agents.append({"name": f"storage-{i}", "bus": bus, "cls": ns.EnergyStorageEnv,
                           "config": {"max_power": 10. + i % 7, "storage_range": (3., 60. + i)}})
        elif kind == 7:
            agents.append({"name": f"ev-{i}", "bus": bus, "cls": ns.EVChargingEnv,
                           "config": {"num_vehicles": 25, "max_charge_rate_kw": 7.,
                                      "peak_threshold": 60., "vehicle_multiplier": 2.}})
        else:
            comps = [
                {"name": "building", "cls": ns.FiveZoneROMThermalEnergyEnv, "config": {}},
                {"name": "pv", "cls": ns.PVEnv,
                 "config": {"profile_csv": "pv_profile.csv", "scaling_factor": 30.}},
                {"name": "storage", "cls": ns.EnergyStorageEnv,
                 "config": {"max_power": 15., "storage_range": (3., 50.)}}]
            agents.append({"name": f"house-{i}", "bus": bus, "cls": ns.MultiComponentEnv,
                           "config": {"components": comps}})
    return {
        "common_config": {"start_time": "08-12-2021 00:00:00", "end_time": "08-13-2021 00:00:00",
                          "control_timedelta": pd.Timedelta(300, "s")},
        "pf_config": {"cls": pf_cls,
                      "config": {"feeder_file": "synthetic123.dss",
                                 "loadshape_file": "ieee_13_dss/annual_hourly_load_profile.csv",
                                 "system_load_rescale_factor": system_load_rescale_factor}},
        "agents": agents}
