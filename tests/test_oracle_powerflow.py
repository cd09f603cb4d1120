"""Pins for the power-flow oracle (oracle/powerflow.py).

The reference's engine (OpenDSS) is unavailable and the reference pins no voltage,
so the oracle is anchored on (a) the published IEEE 13-node solution and (b) physics."""
import json
import os

import numpy as np

from oracle.powerflow import (OracleOpenDSSSolver, compile_feeder, rpn, solve_snapshot,
                              tri_matrix)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
IEEE13 = "ieee_13_dss/IEEE13Nodeckt.dss"


def test_published_ieee13_voltage_profile():
    """Original feeder (regulator taps fixed, capacitors, PQ/Z/I wye+delta loads):
    every published node magnitude within 2e-3 p.u. (the published table has four
    decimals and places the distributed load differently from the 670 lumping)."""
    ckt = compile_feeder(os.path.join(GOLD, "ieee13_original.dss"))
    v = solve_snapshot(ckt, [l["kw"] for l in ckt.loads], [l["kvar"] for l in ckt.loads])
    pu = dict(zip(ckt.node_names, np.abs(v) / ckt.vbase))
    pub = json.load(open(os.path.join(GOLD, "ieee13_published_voltages.json")))
    n = 0
    for bus, vals in pub.items():
        if bus.startswith("_"):
            continue
        for ph, want in enumerate(vals, 1):
            if want is not None:
                assert abs(pu[f"{bus}.{ph}"] - want) < 2e-3, (bus, ph, pu[f"{bus}.{ph}"], want)
                n += 1
    assert n == 35


def test_reference_feeder_structure():
    ckt = compile_feeder(IEEE13)
    assert ckt.n == 38
    assert [l["name"] for l in ckt.loads] == ["671", "634a", "634b", "634c", "645", "675a",
                                              "675b", "675c", "670a", "670b", "670c", "684c"]
    assert sum(len(l["branches"]) for l in ckt.loads) == 14
    assert abs(sum(l["kw"] for l in ckt.loads) - 3038) < 1e-9
    assert np.abs(ckt.Ynet - ckt.Ynet.T).max() < 1e-9    # reciprocal network
    v0 = np.linalg.solve(ckt.Ynet, ckt.Isrc)
    pu0 = np.abs(v0) / ckt.vbase
    assert np.all(np.abs(pu0 - 1.0001) < 5e-5)           # no-load profile = source p.u.


def test_power_balance_and_tolerance_sweep():
    ckt = compile_feeder(IEEE13)
    for scale in (0.25, 0.7, 1.2):
        kw = [l["kw"] * scale for l in ckt.loads]
        kvar = [l["kvar"] * scale for l in ckt.loads]
        v, it = solve_snapshot(ckt, kw, kvar, tol=1e-12, return_iters=True)
        assert it < 40
        # nodal balance: Ynet V + I_load(V) = I_src  (residual relative to source current)
        resid = ckt.Ynet @ v - ckt.Isrc
        from oracle.powerflow import load_branch_current
        for ld, p, q in zip(ckt.loads, kw, kvar):
            s_ph = complex(p, q) * 1000.0 / len(ld["branches"])
            for a, b in ld["branches"]:
                vb = (v[a] if a >= 0 else 0) - (v[b] if b >= 0 else 0)
                i = load_branch_current(ld["model"], s_ph, vb, ld["vbase"], 0.95, 1.05)
                if a >= 0:
                    resid[a] += i
                if b >= 0:
                    resid[b] -= i
        assert np.abs(resid).max() < 1e-6 * np.abs(ckt.Isrc).max()
        v4 = solve_snapshot(ckt, kw, kvar, tol=1e-4)     # the engine's default tolerance
        assert np.abs(np.abs(v4) - np.abs(v)).max() / ckt.vbase.min() < 1e-3
        assert (np.abs(np.abs(v4) - np.abs(v)) / ckt.vbase).max() < 2e-5


def test_plugin_surface_matches_opendss_py():
    s = OracleOpenDSSSolver(IEEE13, "ieee_13_dss/annual_hourly_load_profile.csv", 0.7)
    s.calculate_power_flow(current_time="01-01-2021 05:00:00")   # tests/.../test_opendss.py:12
    v = s.get_bus_voltages()
    assert len(v) == 38 and "675.3" in v and "sourcebus.1" in v
    assert s.get_bus_voltage_by_name("675c") == v["675.3"]
    assert s.get_bus_voltage_by_name("671") == [v["671.1"], v["671.2"], v["671.3"]]
    base = dict(v)
    s.calculate_power_flow(current_time="01-01-2021 05:00:00",
                           p_controllable_consumed={"675c": 500.0},
                           q_controllable_consumed={"675c": 0.0})
    assert s.get_bus_voltages()["675.3"] < base["675.3"] - 1e-3


def test_dss_value_syntax():
    assert rpn("8 1000 /") == 0.008
    m = tri_matrix("0.3465 | 0.1560 0.3375 | 0.1580 0.1535 0.3414", 3)
    assert m[0, 2] == 0.1580 and m[2, 0] == 0.1580 and m[1, 1] == 0.3375
