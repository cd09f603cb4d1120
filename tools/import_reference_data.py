#!/usr/bin/env python
"""Pack the reference's *input data files* (not source code) into one binary
asset bundle, ``powergridworld_b200/data/assets.npz``.

Run in the authoring container only (``/root/reference`` is absent on the GPU
box).  The bundle is what lets the product and the oracle run the reference's
shipped scenarios without the reference tree:

  pv/<file>            first CSV column with the first line consumed as header,
                       exactly as ``pd.read_csv(f).values[:, 0]`` sees it
                       (gridworld/agents/pv/pv_profile_env.py:68)
  vehicles/<column>    columns of gridworld/agents/vehicles/vehicles.csv
                       (ev_charging_env.py:70-76)
  loadshape/<file>     np.genfromtxt of annual_hourly_load_profile.csv
                       (gridworld/distribution_system/opendss.py:43-44)
  dss/<file>           raw bytes of the IEEE-13 OpenDSS scripts
                       (gridworld/distribution_system/data/ieee_13_dss/)
  building/<key>       five-zone state-space model, from state_space_model.p
                       (gridworld/agents/buildings/five_zone_rom_env.py:49-50)
  hs/vehicles/<column> columns of gridworld/agents/vehicles/vehicles_hs.csv
                       (ev_charging_env_hs.py:71-74)

and copies the Home-Steward scenario data (numbers only: grid cost, timestamps, the PV /
device / vehicle profiles and the component parameters of
gridworld/scenarios/data/env_config.json) to ``powergridworld_b200/data/hs_env_config.json``.
"""
import os
import pickle
import sys

import numpy as np
import pandas as pd

REF = os.environ.get("PGW_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "powergridworld_b200", "data", "assets.npz")


def main():
    out = {}
    pv_dir = os.path.join(REF, "gridworld/agents/pv/profiles")
    for f in sorted(os.listdir(pv_dir)):
        if f.endswith(".csv") and not f.endswith("_hs.csv"):
            out[f"pv/{f}"] = np.ascontiguousarray(
                pd.read_csv(os.path.join(pv_dir, f)).values[:, 0].astype(np.float64))
    veh = pd.read_csv(os.path.join(REF, "gridworld/agents/vehicles/vehicles.csv"))
    for c in ["start_time_min", "end_time_park_min", "energy_required_kwh"]:
        out[f"vehicles/{c}"] = veh[c].values.astype(np.float64)
    dss_dir = os.path.join(REF, "gridworld/distribution_system/data/ieee_13_dss")
    out["loadshape/ieee_13_dss/annual_hourly_load_profile.csv"] = np.genfromtxt(
        os.path.join(dss_dir, "annual_hourly_load_profile.csv"))
    for f in ["IEEE13Nodeckt.dss", "IEEELineCodes.dss"]:
        with open(os.path.join(dss_dir, f), "rb") as fh:
            out[f"dss/ieee_13_dss/{f}"] = np.frombuffer(fh.read(), dtype=np.uint8)
    with open(os.path.join(REF, "gridworld/agents/buildings/data/state_space_model.p"), "rb") as fh:
        models = pickle.load(fh)
    out["building/ss_A"] = np.array([m["ss_A"].squeeze() for m in models], dtype=np.float64)
    out["building/ss_B"] = np.array([np.asarray(m["ss_B"]).reshape(-1) for m in models], dtype=np.float64)
    out["building/ss_C"] = np.array([np.asarray(m["ss_C"]).squeeze() for m in models], dtype=np.float64)
    out["building/ss_K"] = np.array([m["ss_K"].squeeze() for m in models], dtype=np.float64)
    out["building/mean_output"] = np.array([m["mean_output"].squeeze() for m in models], dtype=np.float64)
    out["building/input_sel_list"] = np.array([np.asarray(m["input_sel_list"]).reshape(-1) for m in models], dtype=np.int64)
    out["building/neighbors"] = np.array([m["neighbors"] for m in models], dtype=np.int64)
    out["building/x_k0"] = np.array([m["x_k"].squeeze() for m in models], dtype=np.float64)
    veh_hs = pd.read_csv(os.path.join(REF, "gridworld/agents/vehicles/vehicles_hs.csv"))
    for c in ["start_time_min", "end_time_park_min", "energy_required_kwh"]:
        out[f"hs/vehicles/{c}"] = veh_hs[c].values.astype(np.float64)
    import json
    with open(os.path.join(REF, "gridworld/scenarios/data/env_config.json")) as fh:
        hs_cfg = json.load(fh)
    with open(os.path.join(os.path.dirname(OUT), "hs_env_config.json"), "w") as fh:
        json.dump(hs_cfg, fh, separators=(",", ":"))
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(f"{k:60s} {v.dtype} {v.shape}")
    print("wrote", os.path.abspath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
