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


def test_constants_of_the_binding_match_the_header():
    """Every PGW_* enumerator / #define the Python binding mirrors has the header's value."""
    text = open(os.path.join(ROOT, "include", "pgw.h")).read()
    defs = {m.group(1): int(m.group(2).rstrip("u"), 0)
            for m in re.finditer(r"#define\s+PGW_([A-Z0-9_]+)\s+(\d+u?)\b", text)}
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\bPGW_([A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    known = {**defs, **enums}
    checked = 0
    for name in ("STORAGE", "PV", "EV", "BUILDING", "HS_BEGIN", "HS_PV", "HS_STORAGE", "HS_EV", "HS_DEVICES",
                 "HS_MAX_COMPONENTS", "F_TELEMETRY", "HS_TEL_ROWS", "F_RESCALE", "F_GRID_AWARE",
                 "F_PV_VOLT_REWARD", "F_STALE_REWARD", "F_BUILDING_FAST", "OPT_PF_KERNEL", "OPT_WARM_START",
                 "OPT_GRAPHS", "OPT_PDL", "OPT_CLIP_INIT_SOC", "OPT_PF_POLISH", "OPT_PF_TC_TOL_NANO",
                 "OPT_FUSED", "OPT_HOST_CHUNKS", "OPT_HOST_ZERO_COPY", "ABI_VERSION", "NUM_STATS"):
        assert name in known, f"PGW_{name} not found in pgw.h"
        assert getattr(N, name) == known[name], (name, getattr(N, name), known[name])
        checked += 1
    assert checked == 29
