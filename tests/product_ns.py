"""Product plugin namespace for the scenario builders (lives in the package)."""
from powergridworld_b200.scenarios.namespace import PRODUCT_NS  # noqa: F401
