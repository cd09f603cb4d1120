"""Access to the packaged input data (data/assets.npz, produced from the reference's
data files by tools/import_reference_data.py): PV profiles, the vehicle roster,
the IEEE-13 OpenDSS scripts and load shape, and the five-zone building model."""
import os
from functools import lru_cache

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "assets.npz")


@lru_cache(maxsize=None)
def _bundle():
    with np.load(_PATH) as z:
        return {k: z[k] for k in z.files}


def array(key: str) -> np.ndarray:
    b = _bundle()
    if key not in b:
        raise FileNotFoundError(f"no packaged asset {key!r}")
    return b[key].copy()


def has(key: str) -> bool:
    return key in _bundle()


def dss_text(rel_path: str) -> str:
    want = ("dss/" + os.path.normpath(rel_path).replace("\\", "/")).lower()
    for k, v in _bundle().items():
        if k.lower() == want:
            return v.tobytes().decode("utf-8", errors="replace")
    raise FileNotFoundError(f"no packaged DSS script {rel_path!r}")


def dss_text_by_basename(name: str) -> str:
    for k, v in _bundle().items():
        if k.startswith("dss/") and k.lower().endswith("/" + name.lower()):
            return v.tobytes().decode("utf-8", errors="replace")
    raise FileNotFoundError(f"no packaged DSS script named {name!r}")
