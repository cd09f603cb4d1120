"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the OpenDSS snapshot power flow.

PARITY UNPINNED against the real engine: the reference delegates the solve to
the third-party, un-vendored ``OpenDSSDirect.py==0.6.1`` (requirements.txt:6;
-> dss_python -> DSS C-API, the Pascal OpenDSS engine + KLUSolve), which is
neither under /root/reference nor installable here (no network), and none of
the reference's tests or notebooks pins a voltage (tests/distribution_system/
test_opendss.py:7-16 discards the result).  What pins this file instead:
  * the published IEEE 13-node test-feeder solution (Kersting / IEEE PES
    DSASC), reproduced in tests/test_oracle_powerflow.py from an authored
    script of the original feeder (regulator taps fixed, capacitors, all six
    load models) -- this validates the element models and the solver;
  * physics self-checks (power balance, no-load profile, symmetry of Y).

It restates, for the call sites in gridworld/distribution_system/opendss.py:
  :36-51   compile the feeder script, read the 8760-row load shape
  :54-77   snapshot base kW/kvar of all Model=1 loads, in definition order
  :80-135  hour-of-year -> load-shape coefficient -> per-load kW/kvar
           (+ controllable P/Q matched by *load name*), ``Solve mode=snap``
  :156-186 node-name keyed p.u. magnitudes; phase-letter -> node lookup
and the engine's documented element models / "Normal" algorithm (OpenDSS
manual; element sources Vsource, Transformer, Line/LineCode, Load, Capacitor):
fixed-point current injection  V <- Ysys^-1 (I_source + I_comp(V))  where Ysys
contains each load's nominal admittance and I_comp is the difference between
that linear model and the load's actual characteristic.

Modelling assumptions (each one a knob of ``Circuit``):
  * Vsource: MVAsc3/MVAsc1 with X1R1=4, X0R0=3 defaults -> Zs, Zm; grounded
    Norton equivalent.
  * Transformer: per-phase 2-winding units, short-circuit impedance
    (sum of %R) + j XHL on the per-phase kVA base, wye = grounded wye,
    delta winding k across nodes (k, k+1); taps multiply the winding voltage;
    no magnetising branch (defaults %imag=%noloadloss=0).
  * Line: series R+jX per length plus half the shunt capacitance at each end;
    length converted from the line's units to the line code's units; a line
    code given only r/x matrices keeps the engine's default capacitance
    (C1=3.4 nF, C0=1.6 nF per length unit); Switch=y -> 0.001-long section.
  * Load: model 1 (constant PQ) inside [Vminpu, Vmaxpu] = [0.95, 1.05] of the
    load's own base, constant impedance outside (continuous at the band edge);
    model 2 constant Z; model 5 constant current magnitude.  Delta loads act on
    the line-to-line voltages.
  * Capacitor: fixed shunt susceptance from kvar at rated kV.
  * Per-unit: each bus takes the ``Set Voltagebases`` entry nearest to its
    no-load voltage (``calcv``).
"""
from __future__ import annotations

import math
import os
import re
from datetime import datetime

import numpy as np
import pandas as pd

from oracle.multiagent import PowerFlowSolver

SQRT3 = math.sqrt(3.0)
TWO_PI = 2.0 * math.pi
UNIT_IN_M = {"none": None, "mi": 1609.344, "kft": 304.8, "km": 1000.0, "m": 1.0,
             "ft": 0.3048, "in": 0.0254, "cm": 0.01}

_ASSET_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                           "powergridworld_b200", "data", "assets.npz")


# ================================================================== DSS script reader
def _strip_comments(text: str):
    """Yield logical lines: block comments, '!' and '//' comments removed."""
    in_block = False
    for raw in text.splitlines():
        line = raw.strip()
        if in_block:
            if "*/" in line:
                in_block = False
            continue
        if line.startswith("/*"):
            if "*/" not in line:
                in_block = True
            continue
        for mark in ("!", "//"):
            k = line.find(mark)
            if k >= 0:
                line = line[:k]
        line = line.strip()
        if line:
            yield line


_TOKEN = re.compile(r"""\(([^)]*)\)|\[([^\]]*)\]|"([^"]*)"|'([^']*)'|([^\s,=]+)|(=)""")


def _tokens(line: str):
    """Split into (text, quoted?) tokens; '=' is its own token."""
    out = []
    for m in _TOKEN.finditer(line):
        if m.group(6):
            out.append(("=", False))
        elif m.group(5) is not None:
            out.append((m.group(5), False))
        else:
            body = next(g for g in m.groups()[:4] if g is not None)
            out.append((body, True))
    return out


def _pairs(tokens):
    """[(name|None, value)] from a token list with optional name = value syntax."""
    out, i = [], 0
    while i < len(tokens):
        if i + 2 < len(tokens) and tokens[i + 1] == ("=", False):
            out.append((tokens[i][0].lower(), tokens[i + 2][0]))
            i += 3
        else:
            out.append((None, tokens[i][0]))
            i += 1
    return out


def rpn(value: str) -> float:
    """A numeric property; several items are evaluated as inline RPN, e.g. ``8 1000 /``."""
    items = value.replace(",", " ").split()
    if len(items) == 1:
        return float(items[0])
    stack = []
    for it in items:
        if it in "+-*/":
            b, a = stack.pop(), stack.pop()
            stack.append({"+": a + b, "-": a - b, "*": a * b, "/": a / b}[it])
        else:
            stack.append(float(it))
    return stack[-1]


def numbers(value: str):
    return [float(x) for x in value.replace(",", " ").replace("|", " ").split()]


def tri_matrix(value: str, n: int) -> np.ndarray:
    """Lower-triangular '|'-separated rows (or a full matrix) -> symmetric n x n."""
    rows = [[float(x) for x in r.replace(",", " ").split()] for r in value.split("|")]
    rows = [r for r in rows if r]
    m = np.zeros((n, n))
    if len(rows) == 1 and len(rows[0]) == n * n:
        return np.array(rows[0]).reshape(n, n)
    for i, r in enumerate(rows):
        for j, v in enumerate(r):
            m[i, j] = v
            m[j, i] = v
    return m


def bus_spec(spec: str, nph: int):
    """'632.3.2' -> ('632', [3, 2]); nodes default to 1..nph."""
    parts = spec.lower().split(".")
    nodes = [int(p) for p in parts[1:]] or list(range(1, nph + 1))
    return parts[0], nodes


class Script:
    """The element records of a DSS script (subset used by the IEEE test feeders)."""

    def __init__(self):
        self.source = {}
        self.transformers = []
        self.linecodes = {}
        self.lines = []
        self.loads = []
        self.capacitors = []
        self.voltage_bases = []
        self.frequency = 60.0
        self._records = []

    def read(self, text: str, opener=None):
        last = None
        for line in _strip_comments(text):
            toks = _tokens(line)
            head = toks[0][0].lower()
            if head == "~" or head.startswith("~"):
                rest = toks[1:] if head == "~" else [(toks[0][0][1:], False)] + toks[1:]
                if last is not None:
                    last["props"].extend(_pairs(rest))
                continue
            if head == "more":
                if last is not None:
                    last["props"].extend(_pairs(toks[1:]))
                continue
            if head == "new":
                pr = _pairs(toks[1:])
                cls, _, name = pr[0][1].partition(".")      # "Class.Name" (or object=Class.Name)
                last = {"cls": cls.lower(), "name": name.lower(), "props": pr[1:]}
                self._records.append(last)
                continue
            last = None
            if head == "redirect" and opener is not None:
                self.read(opener(toks[1][0]), opener)
            elif head == "set":
                for k, v in _pairs(toks[1:]):
                    if k == "voltagebases":
                        self.voltage_bases = numbers(v)
                    elif k == "defaultbasefrequency":
                        self.frequency = float(v)
            # clear / calcv / solve / buscoords / show ...: nothing to record
        return self

    def finish(self):
        for rec in self._records:
            getattr(self, "_new_" + rec["cls"], lambda r: None)(rec)
        return self

    # ---- element builders -------------------------------------------------------
    def _new_circuit(self, rec):
        s = {"bus1": "sourcebus", "basekv": 115.0, "pu": 1.0, "angle": 0.0, "phases": 3,
             "mvasc3": 2000.0, "mvasc1": 2100.0, "x1r1": 4.0, "x0r0": 3.0}
        for k, v in rec["props"]:
            if k in ("bus1",):
                s[k] = v.lower()
            elif k in s:
                s[k] = rpn(v)
        self.source = s

    def _new_transformer(self, rec):
        new_wdg = lambda: dict(bus=None, conn="wye", kv=12.47, kva=1000.0, r=0.2, tap=1.0)
        t = {"name": rec["name"], "phases": 3, "xhl": 7.0, "wdg": [new_wdg(), new_wdg()]}
        cur = 0
        for k, v in rec["props"]:
            if k == "phases":
                t["phases"] = int(rpn(v))
            elif k == "windings":
                t["wdg"] = [new_wdg() for _ in range(int(rpn(v)))]
            elif k == "xhl":
                t["xhl"] = rpn(v)
            elif k == "wdg":
                cur = int(rpn(v)) - 1
                while len(t["wdg"]) <= cur:
                    t["wdg"].append(new_wdg())
            elif k == "bus":
                t["wdg"][cur]["bus"] = v.lower()
            elif k == "conn":
                t["wdg"][cur]["conn"] = "delta" if v.lower() in ("delta", "ll", "d") else "wye"
            elif k == "kv":
                t["wdg"][cur]["kv"] = rpn(v)
            elif k == "kva":
                t["wdg"][cur]["kva"] = rpn(v)
            elif k == "%r":
                t["wdg"][cur]["r"] = rpn(v)
            elif k == "tap":
                t["wdg"][cur]["tap"] = rpn(v)
            elif k == "buses":
                for w, b in zip(t["wdg"], v.replace(",", " ").split()):
                    w["bus"] = b.lower()
            elif k == "conns":
                for w, c in zip(t["wdg"], v.replace(",", " ").split()):
                    w["conn"] = "delta" if c.lower() in ("delta", "ll", "d") else "wye"
            elif k == "kvs":
                for w, x in zip(t["wdg"], numbers(v)):
                    w["kv"] = x
            elif k == "kvas":
                for w, x in zip(t["wdg"], numbers(v)):
                    w["kva"] = x
            elif k == "%loadloss":
                for w in t["wdg"]:
                    w["r"] = rpn(v) / 2.0
            elif k == "taps":
                for w, x in zip(t["wdg"], numbers(v)):
                    w["tap"] = x
        self.transformers.append(t)

    def _new_linecode(self, rec):
        # engine defaults: R1=.058 X1=.1206 R0=.1784 X0=.4047 C1=3.4 C0=1.6 (nF), units none
        c = {"nphases": 3, "r": None, "x": None, "c": None, "units": "none"}
        for k, v in rec["props"]:
            if k == "nphases":
                c["nphases"] = int(rpn(v))
        n = c["nphases"]
        for k, v in rec["props"]:
            if k == "rmatrix":
                c["r"] = tri_matrix(v, n)
            elif k == "xmatrix":
                c["x"] = tri_matrix(v, n)
            elif k == "cmatrix":
                c["c"] = tri_matrix(v, n)
            elif k == "units":
                c["units"] = v.lower()
        if c["r"] is None:
            c["r"] = _sym(0.058, 0.1784, n)
        if c["x"] is None:
            c["x"] = _sym(0.1206, 0.4047, n)
        if c["c"] is None:
            c["c"] = _sym(3.4, 1.6, n)
        self.linecodes[rec["name"]] = c

    def _new_line(self, rec):
        ln = {"name": rec["name"], "phases": 3, "bus1": None, "bus2": None, "linecode": None,
              "length": 1.0, "units": "none", "switch": False,
              "r1": 0.058, "x1": 0.1206, "r0": 0.1784, "x0": 0.4047, "c1": 3.4, "c0": 1.6}
        for k, v in rec["props"]:
            if k == "phases":
                ln["phases"] = int(rpn(v))
            elif k in ("bus1", "bus2"):
                ln[k] = v.lower()
            elif k == "linecode":
                ln["linecode"] = v.lower()
            elif k == "length":
                ln["length"] = rpn(v)
            elif k == "units":
                ln["units"] = v.lower()
            elif k == "switch":
                if v.lower() in ("y", "yes", "true", "t"):
                    # Switch=y presets r1=x1=r0=x0=1, c1=1.1, c0=1, length 0.001, no units;
                    # later properties on the same line override these.
                    ln.update(switch=True, r1=1.0, x1=1.0, r0=1.0, x0=1.0, c1=1.1, c0=1.0,
                              length=0.001, units="none")
            elif k in ("r1", "x1", "r0", "x0", "c1", "c0"):
                ln[k] = rpn(v)
        self.lines.append(ln)

    def _new_load(self, rec):
        ld = {"name": rec["name"], "bus1": None, "phases": 3, "conn": "wye", "model": 1,
              "kv": 12.47, "kw": 10.0, "kvar": 5.0, "vminpu": 0.95, "vmaxpu": 1.05}
        for k, v in rec["props"]:
            if k == "bus1":
                ld["bus1"] = v.lower()
            elif k == "phases":
                ld["phases"] = int(rpn(v))
            elif k == "conn":
                ld["conn"] = "delta" if v.lower() in ("delta", "ll", "d") else "wye"
            elif k == "model":
                ld["model"] = int(rpn(v))
            elif k in ("kv", "kw", "kvar", "vminpu", "vmaxpu"):
                ld[k] = rpn(v)
        self.loads.append(ld)

    def _new_capacitor(self, rec):
        c = {"name": rec["name"], "bus1": None, "phases": 3, "kvar": 1200.0, "kv": 12.47,
             "conn": "wye"}
        for k, v in rec["props"]:
            if k == "bus1":
                c["bus1"] = v.lower()
            elif k == "phases":
                c["phases"] = int(rpn(v))
            elif k in ("kvar", "kv"):
                c[k] = rpn(v)
            elif k == "conn":
                c["conn"] = "delta" if v.lower() in ("delta", "ll", "d") else "wye"
        self.capacitors.append(c)


def _sym(pos, zero, n):
    """n x n matrix with self = (2*pos+zero)/3 and mutual = (zero-pos)/3."""
    s, m = (2.0 * pos + zero) / 3.0, (zero - pos) / 3.0
    return np.full((n, n), m) + np.eye(n) * (s - m)


# ================================================================== network assembly
class Circuit:
    """Nodal model: Y of the passive network, source injections, load table."""

    def __init__(self, script: Script, default_line_capacitance=True):
        self.s = script
        self.freq = script.frequency
        self.node_of = {}          # (bus, node) -> index
        self.node_names = []       # 'bus.node'
        self._stamps = []          # (rows, ymatrix)
        self.default_line_capacitance = default_line_capacitance
        self._build()

    # ---- node bookkeeping
    def _n(self, bus, node):
        if node == 0:
            return -1
        key = (bus, node)
        if key not in self.node_of:
            self.node_of[key] = len(self.node_names)
            self.node_names.append(f"{bus}.{node}")
        return self.node_of[key]

    def _stamp(self, idx, y, shunt=False):
        self._stamps.append((list(idx), np.asarray(y, dtype=np.complex128), shunt))

    def _assemble(self, series_only=False):
        n = len(self.node_names)
        Y = np.zeros((n, n), dtype=np.complex128)
        for idx, y, shunt in self._stamps:
            if series_only and shunt:
                continue
            for a, ia in enumerate(idx):
                if ia < 0:
                    continue
                for b, ib in enumerate(idx):
                    if ib >= 0:
                        Y[ia, ib] += y[a, b]
        return Y

    def _build(self):
        s = self.s
        w = TWO_PI * self.freq
        # ---- source (Thevenin -> Norton)
        src = s.source
        kv = src["basekv"]
        x1 = kv ** 2 / src["mvasc3"] / math.sqrt(1.0 + 1.0 / src["x1r1"] ** 2)
        r1 = x1 / src["x1r1"]
        isc1 = src["mvasc1"] * 1000.0 / (SQRT3 * kv)
        a = 1.0 + src["x0r0"] ** 2
        b = 4.0 * (r1 + x1 * src["x0r0"])
        c = 4.0 * (r1 * r1 + x1 * x1) - (SQRT3 * kv * 1000.0 / isc1) ** 2
        r0 = (-b + math.sqrt(b * b - 4.0 * a * c)) / (2.0 * a)
        x0 = r0 * src["x0r0"]
        z1, z0 = complex(r1, x1), complex(r0, x0)
        zs, zm = (2.0 * z1 + z0) / 3.0, (z0 - z1) / 3.0
        nph = int(src["phases"])
        Zsrc = np.full((nph, nph), zm, dtype=np.complex128) + np.eye(nph) * (zs - zm)
        Ysrc = np.linalg.inv(Zsrc)
        sbus, snodes = bus_spec(src["bus1"], nph)
        self.src_idx = [self._n(sbus, k) for k in snodes]
        vmag = kv * src["pu"] * 1000.0 / SQRT3
        self.src_v = np.array([vmag * np.exp(1j * math.radians(src["angle"] - 120.0 * k))
                               for k in range(nph)])
        self._stamp(self.src_idx, Ysrc)
        self.src_inj = Ysrc @ self.src_v

        # ---- transformers (2-winding banks of single-phase units)
        for t in s.transformers:
            w1, w2 = t["wdg"][0], t["wdg"][1]
            nph = t["phases"]
            zpu = complex((w1["r"] + w2["r"]) / 100.0, t["xhl"] / 100.0)
            va_phase = w1["kva"] * 1000.0 / nph
            y1v = va_phase / zpu                       # admittance on a 1-volt base

            def winding_volts(wd):
                if nph == 1 or wd["conn"] == "delta":
                    v = wd["kv"] * 1000.0
                else:
                    v = wd["kv"] * 1000.0 / SQRT3
                return v * wd["tap"]

            n1, n2 = winding_volts(w1), winding_volts(w2)
            b1, nodes1 = bus_spec(w1["bus"], nph)
            b2, nodes2 = bus_spec(w2["bus"], nph)

            def terminals(bus, nodes, conn, k):
                if conn == "delta":
                    return self._n(bus, nodes[k]), self._n(bus, nodes[(k + 1) % len(nodes)])
                ret = nodes[nph] if len(nodes) > nph else 0
                if nph == 1 and len(nodes) > 1:
                    ret = nodes[1]
                return self._n(bus, nodes[k]), self._n(bus, ret)

            for k in range(nph):
                p1, q1 = terminals(b1, nodes1, w1["conn"], k)
                p2, q2 = terminals(b2, nodes2, w2["conn"], k)
                # winding currents from winding voltages, then to the four terminals
                g = np.array([1.0 / n1, -1.0 / n1, -1.0 / n2, 1.0 / n2])
                self._stamp([p1, q1, p2, q2], y1v * np.outer(g, g))

        # ---- lines
        for ln in s.lines:
            nph = ln["phases"]
            if ln["linecode"] is not None:
                lc = s.linecodes[ln["linecode"]]
                R, X, C = lc["r"], lc["x"], lc["c"]
                scale = _unit_ratio(ln["units"], lc["units"])
            else:
                R = _sym(ln["r1"], ln["r0"], nph)
                X = _sym(ln["x1"], ln["x0"], nph)
                C = _sym(ln["c1"], ln["c0"], nph)
                scale = 1.0
            length = ln["length"] * scale
            Z = (R + 1j * X) * length
            Yc = 1j * w * C * 1e-9 * length
            Zi = np.linalg.inv(Z)
            b1, n1 = bus_spec(ln["bus1"], nph)
            b2, n2 = bus_spec(ln["bus2"], nph)
            i1 = [self._n(b1, k) for k in n1[:nph]]
            i2 = [self._n(b2, k) for k in n2[:nph]]
            self._stamp(i1 + i2, np.block([[Zi, -Zi], [-Zi, Zi]]))
            self._stamp(i1 + i2, np.block([[Yc / 2, 0 * Yc], [0 * Yc, Yc / 2]]), shunt=True)

        # ---- capacitors (fixed shunts)
        for cp in s.capacitors:
            nph = cp["phases"]
            bus, nodes = bus_spec(cp["bus1"], nph)
            if cp["conn"] == "delta":
                vph, pairs = cp["kv"] * 1000.0, [(nodes[k], nodes[(k + 1) % nph]) for k in range(nph)]
            else:
                vph = cp["kv"] * 1000.0 / (SQRT3 if nph > 1 else 1.0)
                pairs = [(nodes[k], 0) for k in range(nph)]
            bsus = cp["kvar"] * 1000.0 / nph / vph ** 2
            for p, q in pairs:
                self._stamp([self._n(bus, p), self._n(bus, q)],
                            1j * bsus * np.array([[1, -1], [-1, 1]]), shunt=True)

        # ---- load terminals (this also creates nodes that only loads touch)
        self.loads = []
        for ld in s.loads:
            nph = ld["phases"]
            bus, nodes = bus_spec(ld["bus1"], nph)
            if ld["conn"] == "delta":
                vbase = ld["kv"] * 1000.0
                if nph == 1:
                    pairs = [(nodes[0], nodes[1])]
                else:
                    pairs = [(nodes[k], nodes[(k + 1) % nph]) for k in range(nph)]
            else:
                vbase = ld["kv"] * 1000.0 / (SQRT3 if nph > 1 else 1.0)
                pairs = [(nodes[k], nodes[nph] if len(nodes) > nph else 0) for k in range(nph)]
            br = [(self._n(bus, p), self._n(bus, q)) for p, q in pairs]
            self.loads.append(dict(ld, vbase=vbase, branches=br))

        self.n = len(self.node_names)
        self.Ynet = self._assemble()
        I = np.zeros(self.n, dtype=np.complex128)
        I[self.src_idx] = self.src_inj
        self.Isrc = I
        self._assign_voltage_bases()

    def _assign_voltage_bases(self):
        """``Set Voltagebases`` + ``calcv``: zero-load solve on the series-only Y."""
        Yser = self._assemble(series_only=True)
        v0 = np.linalg.solve(Yser, self.Isrc)
        bases = np.array(self.s.voltage_bases or [self.s.source["basekv"]]) * 1000.0 / SQRT3
        mag = np.abs(v0)
        bus_first = {}
        for (bus, node), i in self.node_of.items():
            bus_first.setdefault(bus, i)
        self.vbase = np.zeros(self.n)
        for (bus, node), i in self.node_of.items():
            m = mag[bus_first[bus]]
            self.vbase[i] = bases[np.argmin(np.abs(1.0 - m / bases))]
        self.v_noload_series = v0


def _unit_ratio(frm, to):
    """Factor converting a length in ``frm`` units into ``to`` units."""
    a, b = UNIT_IN_M.get(frm), UNIT_IN_M.get(to)
    if a is None or b is None:
        return 1.0
    return a / b


# ================================================================== load characteristic
def load_branch_current(model, s_ph, v, vbase, vminpu, vmaxpu):
    """Current drawn by one load phase (flowing from its + to its - terminal).

    model 1: conj(S/V) inside the band, constant Z outside; 2: constant Z;
    5: constant current magnitude at the rated power factor.
    """
    yeq = np.conj(s_ph) / vbase ** 2
    vm = abs(v)
    if model == 2:
        return yeq * v
    if model == 5:
        return np.conj(s_ph / vbase) * v / vm if vm > 0 else 0.0
    if vm <= vminpu * vbase:
        return (yeq / vminpu ** 2) * v
    if vm > vmaxpu * vbase:
        return (yeq / vmaxpu ** 2) * v
    return np.conj(s_ph / v)


def solve_snapshot(ckt: Circuit, load_kw, load_kvar, v_start=None, tol=1e-10, max_iter=100,
                   return_iters=False):
    """OpenDSS-style "Normal" solve: loads' nominal admittances live in Ysys and
    the iteration injects compensation currents (see module docstring)."""
    n = ckt.n
    Ysys = ckt.Ynet.copy()
    branch = []          # (p, q, model, s_ph, vbase, vmin, vmax, yeq)
    for ld, kw, kvar in zip(ckt.loads, load_kw, load_kvar):
        nbr = len(ld["branches"])
        s_ph = complex(kw, kvar) * 1000.0 / nbr
        yeq = np.conj(s_ph) / ld["vbase"] ** 2
        for p, q in ld["branches"]:
            for a, sa in ((p, 1.0), (q, -1.0)):
                for b, sb in ((p, 1.0), (q, -1.0)):
                    if a >= 0 and b >= 0:
                        Ysys[a, b] += sa * sb * yeq
            branch.append((p, q, ld["model"], s_ph, ld["vbase"], ld["vminpu"], ld["vmaxpu"], yeq))
    import scipy.linalg as sla
    fac = sla.lu_factor(Ysys)
    Ysys_x = Ysys.astype(np.clongdouble)

    def solve(rhs):
        # LU solve with two refinement sweeps on extended-precision residuals
        # (cond(Y) ~ 2e8: a 1e-7 ohm switch and a very stiff source)
        x = sla.lu_solve(fac, rhs)
        for _ in range(2):
            r = (rhs.astype(np.clongdouble) - Ysys_x @ x.astype(np.clongdouble)).astype(np.complex128)
            x = x + sla.lu_solve(fac, r)
        return x

    v = np.linalg.solve(ckt.Ynet, ckt.Isrc) if v_start is None else v_start.copy()
    it = 0
    for it in range(1, max_iter + 1):
        inj = ckt.Isrc.copy()
        for p, q, model, s_ph, vb, vmin, vmax, yeq in branch:
            vbr = (v[p] if p >= 0 else 0.0) - (v[q] if q >= 0 else 0.0)
            comp = yeq * vbr - load_branch_current(model, s_ph, vbr, vb, vmin, vmax)
            if p >= 0:
                inj[p] += comp
            if q >= 0:
                inj[q] -= comp
        v_new = solve(inj)
        err = np.max(np.abs(v_new - v) / ckt.vbase)
        v = v_new
        if err < tol:
            break
    if return_iters:
        return v, it
    return v


# ================================================================== plugin class
def _asset_text_opener(prefix):
    with np.load(_ASSET_PATH) as z:
        files = {k: z[k] for k in z.files if k.startswith("dss/")}

    def opener(name):
        key = os.path.normpath(os.path.join("dss", prefix, name)).replace("\\", "/")
        for k, v in files.items():
            if k.lower() == key.lower():
                return v.tobytes().decode("utf-8", errors="replace")
        raise FileNotFoundError(key)
    return opener


def compile_feeder(feeder_file: str = None, text: str = None) -> Circuit:
    if text is not None:
        return Circuit(Script().read(text, lambda f: "").finish())
    packaged = os.path.join(os.path.dirname(_ASSET_PATH), "feeders", feeder_file)
    if not os.path.isfile(feeder_file) and os.path.isfile(packaged):
        feeder_file = packaged                     # e.g. the authored "synthetic123.dss"
    if os.path.isfile(feeder_file):
        base = os.path.dirname(feeder_file)

        def opener(name):
            path = os.path.join(base, name)
            if os.path.isfile(path):
                with open(path) as fh:
                    return fh.read()
            with np.load(_ASSET_PATH) as z:         # redirected IEEE data from the asset bundle
                for k in z.files:
                    if k.startswith("dss/") and k.lower().endswith("/" + name.lower()):
                        return z[k].tobytes().decode("utf-8", errors="replace")
            raise FileNotFoundError(path)
        with open(feeder_file) as fh:
            body = fh.read()
    else:
        opener = _asset_text_opener(os.path.dirname(feeder_file))
        body = opener(os.path.basename(feeder_file))
    return Circuit(Script().read(body, opener).finish())


class OracleOpenDSSSolver(PowerFlowSolver):
    """Drop-in for gridworld/distribution_system/opendss.py:15-186 (same ctor keywords,
    same dict outputs) on top of ``solve_snapshot``."""

    def __init__(self, feeder_file, loadshape_file, system_load_rescale_factor=1.0,
                 tol=1e-10, **kwargs):
        super().__init__(**kwargs)
        self.ckt = compile_feeder(feeder_file)
        self.system_load_rescale_factor = system_load_rescale_factor
        if os.path.isfile(loadshape_file):
            self.annual_hourly_load_profile = np.genfromtxt(loadshape_file)
        else:
            with np.load(_ASSET_PATH) as z:
                self.annual_hourly_load_profile = z["loadshape/" + loadshape_file]
        self.tol = tol
        self.bus_voltages = {}
        # opendss.py:54-77 -- Model==1 loads only, in definition order
        self.pq = [i for i, ld in enumerate(self.ckt.loads) if ld["model"] == 1]
        self.load_bus_name = [self.ckt.loads[i]["name"] for i in self.pq]
        self.base_load = np.array([[self.ckt.loads[i]["kw"], self.ckt.loads[i]["kvar"]]
                                   for i in self.pq])
        self.v = None
        self.last_iterations = 0

    def calculate_power_flow(self, p_controllable_consumed=None, q_controllable_consumed=None,
                             current_time=None, **kwargs):
        t = pd.Timestamp(current_time)
        hour = int((t - datetime(t.year, 1, 1)).total_seconds() // 3600)      # :98-105
        cur = self.annual_hourly_load_profile[hour] * self.base_load * \
            self.system_load_rescale_factor                                   # :106-108
        if p_controllable_consumed is not None:
            for i, name in enumerate(self.load_bus_name):                     # :115-129
                cur[i, 0] += p_controllable_consumed.get(name, 0.0)
                cur[i, 1] += (q_controllable_consumed or {}).get(name, 0.0)
        kw = np.array([ld["kw"] for ld in self.ckt.loads], dtype=np.float64)
        kvar = np.array([ld["kvar"] for ld in self.ckt.loads], dtype=np.float64)
        kw[self.pq] = cur[:, 0]
        kvar[self.pq] = cur[:, 1]
        self.v, self.last_iterations = solve_snapshot(
            self.ckt, kw, kvar, v_start=self.v, tol=self.tol, return_iters=True)
        mag = np.abs(self.v) / self.ckt.vbase                                 # AllBusMagPu
        for name, m in zip(self.ckt.node_names, mag):                         # :156-165
            self.bus_voltages[name] = float(m)

    def get_bus_voltages(self):
        return self.bus_voltages

    def get_bus_voltage_by_name(self, bus_name):
        # :173-186 (note: str.replace hits every occurrence of the phase letter)
        phase_map = {"a": ".1", "b": ".2", "c": ".3"}
        if bus_name[-1] in phase_map:
            return self.bus_voltages[bus_name.replace(bus_name[-1], phase_map[bus_name[-1]])]
        return [self.bus_voltages[bus_name + p] for p in phase_map.values()]
