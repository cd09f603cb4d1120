"""Product plugin namespace for the Home-Steward builders (lives in the package)."""
from powergridworld_b200.scenarios.namespace import PRODUCT_HS_NS  # noqa: F401
