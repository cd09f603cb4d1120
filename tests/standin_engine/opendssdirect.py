"""NOT the OpenDSS engine.  A stand-in with the handful of `opendssdirect` calls the calibration
hook makes (tests/test_opendss_calibration.py), backed by the oracle's own power flow, so that the
hook's harness -- script export, load bookkeeping, node-name matching -- can be exercised where the
real engine is absent (tests/test_calibration_harness.py puts this directory on PYTHONPATH for a
child pytest).  A pass through this module proves nothing about parity with OpenDSS."""
import numpy as np

from oracle.powerflow import compile_feeder, solve_snapshot

STANDIN = True
_st = {}


def run_command(cmd: str):
    if cmd.startswith("Redirect"):
        ckt = compile_feeder(cmd.split(" ", 1)[1])
        _st.update(ckt=ckt, kw=[l["kw"] for l in ckt.loads], kvar=[l["kvar"] for l in ckt.loads])
    elif cmd.startswith("Solve"):
        ckt = _st["ckt"]
        _st["v"] = np.abs(solve_snapshot(ckt, _st["kw"], _st["kvar"], tol=1e-12)) / ckt.vbase


class Loads:
    i = -1

    @staticmethod
    def First():
        Loads.i = 0
        return 1

    @staticmethod
    def Next():
        Loads.i += 1
        return 0 if Loads.i >= len(_st["ckt"].loads) else Loads.i + 1

    @staticmethod
    def Model():
        return _st["ckt"].loads[Loads.i]["model"]

    @staticmethod
    def Name():
        return _st["ckt"].loads[Loads.i]["name"]

    @staticmethod
    def kW(x):
        _st["kw"][Loads.i] = x

    @staticmethod
    def kvar(x):
        _st["kvar"][Loads.i] = x


class Circuit:
    @staticmethod
    def AllNodeNames():
        return list(_st["ckt"].node_names)

    @staticmethod
    def AllBusMagPu():
        return list(_st["v"])
