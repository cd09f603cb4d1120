"""The scenario catalog lives in the package (powergridworld_b200/scenarios/catalog.py); the
tests and the golden generators keep importing it under this name."""
from powergridworld_b200.scenarios.catalog import *  # noqa: F401,F403
from powergridworld_b200.scenarios.catalog import (  # noqa: F401
    TIME_BASE_VARIANTS, test_heterogeneous_scenario, test_multicomponent_components)
