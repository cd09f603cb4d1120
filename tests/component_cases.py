"""Constructor kwargs of tests/golden/component_configs.npz from their JSON form (shared by the
recording script and the replaying tests)."""
import json
import os

import numpy as np
import pandas as pd

GOLD = os.path.join(os.path.dirname(__file__), "golden", "component_configs.npz")


def build_component(ns_cls, cfg):
    kw = dict(cfg)
    if "control_timedelta_s" in kw:
        kw["control_timedelta"] = pd.Timedelta(kw.pop("control_timedelta_s"), "s")
    if "storage_range" in kw:
        kw["storage_range"] = tuple(kw["storage_range"])
    if "obs_config" in kw:
        kw["obs_config"] = {k: tuple(v) for k, v in kw["obs_config"].items()}
    return ns_cls(**kw)


def load_cases():
    g = np.load(GOLD)
    return g, json.loads(str(g["meta"]))
