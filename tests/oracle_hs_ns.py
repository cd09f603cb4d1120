"""Oracle Home-Steward namespace (lives in oracle/namespace.py)."""
from oracle.namespace import ORACLE_HS_NS  # noqa: F401
