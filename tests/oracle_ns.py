"""Oracle plugin namespace (lives in oracle/namespace.py)."""
from oracle.namespace import ORACLE_NS, storage_socs_to_dict  # noqa: F401
