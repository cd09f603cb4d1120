"""CPU tier: the C-ABI library loads and exports every symbol include/pgw.h declares."""
import ctypes
import os
import re

from powergridworld_b200 import _native as N

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgw.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(pgw_[a-z_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(N.LIB_PATH), "run __graft_entry__.build() first"
    h = ctypes.CDLL(N.LIB_PATH)
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(h, name), f"{name} declared in pgw.h but not exported"
    assert declared == set(N.SYMBOLS), (declared ^ set(N.SYMBOLS))


def test_binding_loads_and_versions_agree():
    lib = N.lib()
    assert lib.pgw_abi_version() == N.ABI_VERSION
    header = open(os.path.join(ROOT, "include", "pgw.h")).read()
    assert f"#define PGW_ABI_VERSION {N.ABI_VERSION}" in header
    assert f"#define PGW_NUM_STATS {N.NUM_STATS}" in header


def test_struct_layouts_match_header_sizes():
    # 12 / 4 int32 fields, no padding
    assert ctypes.sizeof(N.Component) == 48
    assert ctypes.sizeof(N.Agent) == 16


def test_no_cpu_fallback_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tests import scenarios as S
    from tests.product_ns import PRODUCT_NS as NS
    with pytest.raises(N.NativeError):
        NS.MultiAgentEnv(**S.ev_pv_storage_scenario(NS))
