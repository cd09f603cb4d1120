"""Flat <-> dict conversion helpers (live in oracle/flatten.py)."""
from oracle.flatten import *  # noqa: F401,F403
from oracle.flatten import action_layout, flat_obs, unflatten_action  # noqa: F401
