"""Home-Steward scenario catalog, re-exported (powergridworld_b200/scenarios/catalog_hs.py)."""
from powergridworld_b200.scenarios.catalog_hs import *  # noqa: F401,F403
from powergridworld_b200.scenarios.catalog_hs import VARIANTS  # noqa: F401
