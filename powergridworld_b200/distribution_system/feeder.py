"""Feeder compiler: OpenDSS script -> tables for the batched Z-bus kernels.

The reference hands the feeder script to the OpenDSS engine
(gridworld/distribution_system/opendss.py:36-39) and lets it rebuild and factor
the system admittance matrix on every step, because the engine keeps each
load's admittance inside Y.  Here the *network* (source, transformers, lines,
capacitors) is reduced once, on the host, to the small dense operators the GPU
iterates with, shared by every env instance:

    u  = u0 - Zbb i(u)        nb load branches (a wye phase or a delta leg)
    v  = w  - Znb i           nn node voltages

``Zbb = A^T Ynet^-1 A`` and ``Znb = Ynet^-1 A`` with ``A`` the node/branch
incidence matrix, obtained from one sparse LU of the network admittance matrix
(no dense inverse).  All outputs are per unit: branch k on its load's own
voltage base (line-to-neutral for wye, line-to-line for delta), node n on the
``Set Voltagebases`` entry assigned by ``calcv``, powers on 1 MVA.

Supported script subset = what the IEEE test feeders shipped with the reference
use (gridworld/distribution_system/data/ieee_13_dss/*.dss): Clear / Set /
New Circuit|Transformer|LineCode|Line|Load|Capacitor / ``~`` continuation /
Redirect / calcv / Solve / BusCoords, ``!`` ``//`` ``/* */`` comments, inline RPN
``(8 1000 /)``, ``|``-separated lower-triangular matrices, length-unit conversion.
Anything else raises -- there is no fallback engine.
"""
from __future__ import annotations

import cmath
import math
import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from powergridworld_b200 import assets

S_BASE_VA = 1.0e6
_LEN_M = {"mi": 1609.344, "kft": 304.8, "km": 1000.0, "m": 1.0, "ft": 0.3048,
          "in": 0.0254, "cm": 0.01}
_IGNORED = {"clear", "calcv", "calcvoltagebases", "solve", "buscoords", "show", "export",
            "plot", "batchedit", "compile"}


class FeederError(ValueError):
    pass


# ----------------------------------------------------------------------------- lexer
def _logical_lines(text: str):
    block = False
    for raw in text.splitlines():
        ln = raw.strip()
        if block:
            block = "*/" not in ln
            continue
        if ln.startswith("/*"):
            block = "*/" not in ln
            continue
        cut = len(ln)
        for marker in ("!", "//"):
            pos = ln.find(marker)
            if pos != -1:
                cut = min(cut, pos)
        ln = ln[:cut].strip()
        if ln:
            yield ln


_OPEN = {"(": ")", "[": "]", '"': '"', "'": "'", "{": "}"}


def _split(line: str) -> List[str]:
    """Whitespace/comma separated words; (...) [...] "..." '...' {...} groups are one word
    (returned without the delimiters); '=' is a word of its own."""
    words, i, n = [], 0, len(line)
    while i < n:
        ch = line[i]
        if ch.isspace() or ch == ",":
            i += 1
        elif ch == "=":
            words.append("=")
            i += 1
        elif ch in _OPEN:
            close = line.find(_OPEN[ch], i + 1)
            if close == -1:
                raise FeederError(f"unterminated {ch!r} in: {line}")
            words.append(line[i + 1:close])
            i = close + 1
        else:
            j = i
            while j < n and not line[j].isspace() and line[j] not in ",=":
                j += 1
            words.append(line[i:j])
            i = j
    return words


def _assignments(words: List[str]) -> List[Tuple[Optional[str], str]]:
    out, i = [], 0
    while i < len(words):
        if i + 2 < len(words) and words[i + 1] == "=":
            out.append((words[i].lower(), words[i + 2]))
            i += 3
        elif i + 1 < len(words) and words[i + 1] == "=":
            raise FeederError(f"dangling '=' after {words[i]}")
        else:
            out.append((None, words[i]))
            i += 1
    return out


def _num(text: str) -> float:
    """Scalar property; a multi-word value is inline RPN."""
    parts = text.replace(",", " ").split()
    if not parts:
        raise FeederError("empty numeric value")
    if len(parts) == 1:
        return float(parts[0])
    st: List[float] = []
    for w in parts:
        if w == "+":
            b = st.pop(); st[-1] += b
        elif w == "-":
            b = st.pop(); st[-1] -= b
        elif w == "*":
            b = st.pop(); st[-1] *= b
        elif w == "/":
            b = st.pop(); st[-1] /= b
        else:
            st.append(float(w))
    return st[-1]


def _vec(text: str) -> List[float]:
    return [float(w) for w in text.replace(",", " ").replace("|", " ").split()]


def _sym_matrix(text: str, n: int) -> np.ndarray:
    rows = [[float(w) for w in r.replace(",", " ").split()] for r in text.split("|")]
    rows = [r for r in rows if r]
    if len(rows) == 1 and len(rows[0]) == n * n:
        return np.asarray(rows[0], dtype=float).reshape(n, n)
    if len(rows) != n or any(len(r) != k + 1 for k, r in enumerate(rows)):
        raise FeederError(f"expected a lower-triangular {n}x{n} matrix, got {text!r}")
    m = np.zeros((n, n))
    for a, r in enumerate(rows):
        m[a, :a + 1] = r
        m[:a + 1, a] = r
    return m


def _seq_matrix(pos: float, zero: float, n: int) -> np.ndarray:
    return np.where(np.eye(n, dtype=bool), (2 * pos + zero) / 3.0, (zero - pos) / 3.0)


def _is_delta(word: str) -> bool:
    return word.lower() in ("delta", "ll", "d")


def _truthy(word: str) -> bool:
    return word.lower() in ("y", "yes", "true", "t")


# ----------------------------------------------------------------------------- elements
@dataclass
class Terminal:
    bus: str
    nodes: List[int]

    @staticmethod
    def parse(spec: str, nph: int) -> "Terminal":
        bits = spec.lower().split(".")
        nodes = [int(b) for b in bits[1:]]
        if not nodes:
            nodes = list(range(1, nph + 1))
        return Terminal(bits[0], nodes)


@dataclass
class LineCode:
    n: int = 3
    r: Optional[np.ndarray] = None
    x: Optional[np.ndarray] = None
    c: Optional[np.ndarray] = None
    units: str = "none"

    def finalize(self):
        # values the engine keeps when a matrix is not given (sequence defaults)
        if self.r is None:
            self.r = _seq_matrix(0.058, 0.1784, self.n)
        if self.x is None:
            self.x = _seq_matrix(0.1206, 0.4047, self.n)
        if self.c is None:
            self.c = _seq_matrix(3.4, 1.6, self.n)


@dataclass
class Winding:
    bus: str = ""
    delta: bool = False
    kv: float = 12.47
    kva: float = 1000.0
    pct_r: float = 0.2
    tap: float = 1.0


@dataclass
class LoadDef:
    name: str
    bus1: str = ""
    phases: int = 3
    delta: bool = False
    model: int = 1
    kv: float = 12.47
    kw: float = 10.0
    kvar: float = 5.0
    vminpu: float = 0.95
    vmaxpu: float = 1.05


@dataclass
class CompiledFeeder:
    node_names: List[str]
    node_vbase: np.ndarray            # volts, line-to-neutral base of each node
    load_names: List[str]
    load_kw: np.ndarray               # nominal, definition order
    load_kvar: np.ndarray
    load_model: np.ndarray
    branch_load: np.ndarray           # [nb] owning load
    branch_share: np.ndarray          # [nb]
    branch_model: np.ndarray
    branch_vmin: np.ndarray
    branch_vmax: np.ndarray
    zbb: np.ndarray                   # [nb, nb] complex, p.u.
    u0: np.ndarray                    # [nb]
    znb: np.ndarray                   # [nn, nb]
    w: np.ndarray                     # [nn]
    ynet: sp.csc_matrix = field(repr=False, default=None)

    @property
    def nb(self):
        return len(self.branch_load)

    @property
    def nn(self):
        return len(self.node_names)

    @property
    def nl(self):
        return len(self.load_names)

    def node_index(self, name: str) -> int:
        try:
            return self.node_names.index(name.lower())
        except ValueError:
            raise FeederError(f"no node named {name!r} in the feeder") from None

    def load_index(self, name: str) -> int:
        try:
            return self.load_names.index(str(name).lower())
        except ValueError:
            raise FeederError(f"no load named {name!r}: agents attach to *load names* "
                              f"({', '.join(self.load_names)})") from None


class _Network:
    """Accumulates primitive admittance stamps in COO form."""

    def __init__(self):
        self.index: Dict[Tuple[str, int], int] = {}
        self.names: List[str] = []
        self.first_of_bus: Dict[str, int] = {}
        self.rows: List[int] = []
        self.cols: List[int] = []
        self.vals: List[complex] = []
        self.shunt: List[bool] = []

    def node(self, bus: str, n: int) -> int:
        if n == 0:
            return -1
        key = (bus, n)
        k = self.index.get(key)
        if k is None:
            k = len(self.names)
            self.index[key] = k
            self.names.append(f"{bus}.{n}")
            self.first_of_bus.setdefault(bus, k)
        return k

    def add(self, idx: List[int], y: np.ndarray, shunt: bool = False):
        for a, ia in enumerate(idx):
            if ia < 0:
                continue
            for b, ib in enumerate(idx):
                if ib < 0 or y[a, b] == 0:
                    continue
                self.rows.append(ia)
                self.cols.append(ib)
                self.vals.append(complex(y[a, b]))
                self.shunt.append(shunt)

    def matrix(self, with_shunts: bool = True) -> sp.csc_matrix:
        n = len(self.names)
        keep = np.ones(len(self.vals), dtype=bool) if with_shunts else ~np.asarray(self.shunt)
        return sp.coo_matrix(
            (np.asarray(self.vals)[keep], (np.asarray(self.rows)[keep], np.asarray(self.cols)[keep])),
            shape=(n, n), dtype=np.complex128).tocsc()


def _refined_solve(lu, y: sp.csc_matrix, rhs: np.ndarray, sweeps: int = 3) -> np.ndarray:
    """LU solve + iterative refinement with extended-precision residuals.  The network
    matrix has cond ~ 2e8 (a 1e-7 ohm switch and a very stiff source next to ordinary
    lines); refinement brings the operators to ~1e-13 relative instead of ~1e-9."""
    yd = y.toarray().astype(np.clongdouble)
    x = lu.solve(rhs)
    for _ in range(sweeps):
        resid = (rhs.astype(np.clongdouble) - yd @ x.astype(np.clongdouble)).astype(np.complex128)
        x = x + lu.solve(resid)
    return x


class FeederCompiler:
    def __init__(self, reader: Callable[[str], str]):
        self._read = reader
        self.freq = 60.0
        self.vbases_kv: List[float] = []
        self.source: Dict[str, float] = {}
        self.source_bus = "sourcebus"
        self.linecodes: Dict[str, LineCode] = {}
        self.elements: List[Tuple[str, str, List[Tuple[Optional[str], str]]]] = []

    # ---- script pass
    def run(self, path: str) -> "FeederCompiler":
        base = os.path.dirname(path)
        current = None
        for line in _logical_lines(self._read(path)):
            words = _split(line)
            verb = words[0].lower()
            if verb.startswith("~") or verb == "more":
                if current is None:
                    raise FeederError("continuation line without an element")
                rest = words[1:] if verb in ("~", "more") else [words[0][1:]] + words[1:]
                current.extend(_assignments(rest))
            elif verb == "new":
                props = _assignments(words[1:])
                kind, _, name = props[0][1].partition(".")
                current = props[1:]
                self.elements.append((kind.lower(), name.lower(), current))
            elif verb == "redirect":
                current = None
                self.run(os.path.join(base, words[1]))
            elif verb == "set":
                current = None
                for key, val in _assignments(words[1:]):
                    if key == "voltagebases":
                        self.vbases_kv = _vec(val)
                    elif key == "defaultbasefrequency":
                        self.freq = float(val)
            elif verb in _IGNORED:
                current = None
            else:
                raise FeederError(f"unsupported DSS command {words[0]!r} (no fallback engine)")
        return self

    # ---- build pass
    def compile(self) -> CompiledFeeder:
        net = _Network()
        omega = 2.0 * math.pi * self.freq
        loads: List[LoadDef] = []
        src_nodes: List[int] = []
        src_inj = np.zeros(0, dtype=complex)

        for kind, name, props in self.elements:
            if kind == "circuit":
                src_nodes, src_inj = self._source(net, props)
            elif kind == "linecode":
                self._linecode(name, props)
        for kind, name, props in self.elements:
            if kind in ("circuit", "linecode"):
                continue
            if kind == "transformer":
                self._transformer(net, props)
            elif kind == "line":
                self._line(net, props, omega)
            elif kind == "capacitor":
                self._capacitor(net, props)
            elif kind == "load":
                loads.append(self._load(name, props))
            else:
                raise FeederError(f"unsupported element class {kind!r} (no fallback engine)")
        if not src_nodes:
            raise FeederError("script defines no circuit (voltage source)")

        # load branches (a load may be the only element touching a node)
        b_p, b_q, b_load, b_share, b_vb, b_model, b_lo, b_hi = ([] for _ in range(8))
        for li, ld in enumerate(loads):
            term = Terminal.parse(ld.bus1, ld.phases)
            if ld.delta:
                vb = ld.kv * 1000.0
                legs = [(term.nodes[0], term.nodes[1])] if ld.phases == 1 else \
                    [(term.nodes[k], term.nodes[(k + 1) % ld.phases]) for k in range(ld.phases)]
            else:
                vb = ld.kv * 1000.0 / (math.sqrt(3.0) if ld.phases > 1 else 1.0)
                ret = term.nodes[ld.phases] if len(term.nodes) > ld.phases else 0
                legs = [(term.nodes[k], ret) for k in range(ld.phases)]
            for p, q in legs:
                b_p.append(net.node(term.bus, p)); b_q.append(net.node(term.bus, q))
                b_load.append(li); b_share.append(1.0 / len(legs)); b_vb.append(vb)
                b_model.append(ld.model); b_lo.append(ld.vminpu); b_hi.append(ld.vmaxpu)

        nn, nb = len(net.names), len(b_p)
        ynet = net.matrix()
        inj = np.zeros(nn, dtype=complex)
        inj[src_nodes] = src_inj
        lu = spla.splu(ynet)
        solve = lambda rhs: _refined_solve(lu, ynet, rhs)
        w_volts = solve(inj)
        inc = np.zeros((nn, nb), dtype=complex)
        for k, (p, q) in enumerate(zip(b_p, b_q)):
            if p >= 0:
                inc[p, k] += 1.0
            if q >= 0:
                inc[q, k] -= 1.0
        za = solve(inc) if nb else np.zeros((nn, 0), dtype=complex)        # Ynet^-1 A
        zbb_volts = inc.T @ za
        u0_volts = inc.T @ w_volts

        # voltage bases: zero-load solve on the series-only network, nearest base per bus
        v_series = spla.splu(net.matrix(with_shunts=False)).solve(inj)
        ln_bases = np.asarray(self.vbases_kv or [self.source["basekv"]]) * 1000.0 / math.sqrt(3.0)
        vbase = np.empty(nn)
        for (bus, _), k in net.index.items():
            m = abs(v_series[net.first_of_bus[bus]])
            vbase[k] = ln_bases[np.argmin(np.abs(1.0 - m / ln_bases))]

        ub = np.asarray(b_vb, dtype=float)
        return CompiledFeeder(
            node_names=list(net.names), node_vbase=vbase,
            load_names=[ld.name for ld in loads],
            load_kw=np.array([ld.kw for ld in loads]), load_kvar=np.array([ld.kvar for ld in loads]),
            load_model=np.array([ld.model for ld in loads], dtype=np.int32),
            branch_load=np.asarray(b_load, dtype=np.int32),
            branch_share=np.asarray(b_share, dtype=float),
            branch_model=np.asarray(b_model, dtype=np.int32),
            branch_vmin=np.asarray(b_lo, dtype=float), branch_vmax=np.asarray(b_hi, dtype=float),
            zbb=zbb_volts * S_BASE_VA / np.outer(ub, ub),
            u0=u0_volts / ub,
            znb=za * (S_BASE_VA / ub)[None, :] / vbase[:, None],
            w=w_volts / vbase, ynet=ynet)

    # ---- element models
    def _source(self, net: _Network, props):
        s = dict(basekv=115.0, pu=1.0, angle=0.0, phases=3.0, mvasc3=2000.0, mvasc1=2100.0,
                 x1r1=4.0, x0r0=3.0)
        bus = "sourcebus"
        for k, v in props:
            if k == "bus1":
                bus = v.lower()
            elif k in s:
                s[k] = _num(v)
            elif k is not None and k not in ("frequency",):
                raise FeederError(f"unsupported Vsource property {k!r}")
        self.source = s
        kv, nph = s["basekv"], int(s["phases"])
        # short-circuit MVA -> sequence impedances (X/R defaults 4 and 3), then phase frame
        x1 = kv * kv / s["mvasc3"] / math.sqrt(1.0 + 1.0 / s["x1r1"] ** 2)
        r1 = x1 / s["x1r1"]
        i1 = s["mvasc1"] * 1000.0 / (math.sqrt(3.0) * kv)
        qa = 1.0 + s["x0r0"] ** 2
        qb = 4.0 * (r1 + x1 * s["x0r0"])
        qc = 4.0 * (r1 * r1 + x1 * x1) - (math.sqrt(3.0) * kv * 1000.0 / i1) ** 2
        r0 = (-qb + math.sqrt(qb * qb - 4.0 * qa * qc)) / (2.0 * qa)
        z1, z0 = complex(r1, x1), complex(r0, r0 * s["x0r0"])
        zs, zm = (2 * z1 + z0) / 3.0, (z0 - z1) / 3.0
        zsrc = np.where(np.eye(nph, dtype=bool), zs, zm)
        ysrc = np.linalg.inv(zsrc)
        term = Terminal.parse(bus, nph)
        idx = [net.node(term.bus, n) for n in term.nodes[:nph]]
        net.add(idx, ysrc)
        emf = np.array([cmath.rect(kv * s["pu"] * 1000.0 / math.sqrt(3.0),
                                   math.radians(s["angle"] - 120.0 * k)) for k in range(nph)])
        return idx, ysrc @ emf

    def _linecode(self, name: str, props):
        lc = LineCode()
        for k, v in props:
            if k == "nphases":
                lc.n = int(_num(v))
        for k, v in props:
            if k == "rmatrix":
                lc.r = _sym_matrix(v, lc.n)
            elif k == "xmatrix":
                lc.x = _sym_matrix(v, lc.n)
            elif k == "cmatrix":
                lc.c = _sym_matrix(v, lc.n)
            elif k == "units":
                lc.units = v.lower()
            elif k not in ("nphases", "basefreq"):
                raise FeederError(f"unsupported LineCode property {k!r}")
        lc.finalize()
        self.linecodes[name] = lc

    def _transformer(self, net: _Network, props):
        nph, xhl = 3, 7.0
        wdg = [Winding(), Winding()]
        cur = 0
        for k, v in props:
            if k == "phases":
                nph = int(_num(v))
            elif k == "windings":
                if int(_num(v)) != 2:
                    raise FeederError("only two-winding transformers are supported")
            elif k == "xhl":
                xhl = _num(v)
            elif k == "wdg":
                cur = int(_num(v)) - 1
            elif k == "bus":
                wdg[cur].bus = v.lower()
            elif k == "conn":
                wdg[cur].delta = _is_delta(v)
            elif k == "kv":
                wdg[cur].kv = _num(v)
            elif k == "kva":
                wdg[cur].kva = _num(v)
            elif k == "%r":
                wdg[cur].pct_r = _num(v)
            elif k == "tap":
                wdg[cur].tap = _num(v)
            elif k == "buses":
                for wd, b in zip(wdg, v.replace(",", " ").split()):
                    wd.bus = b.lower()
            elif k == "conns":
                for wd, c in zip(wdg, v.replace(",", " ").split()):
                    wd.delta = _is_delta(c)
            elif k == "kvs":
                for wd, x in zip(wdg, _vec(v)):
                    wd.kv = x
            elif k == "kvas":
                for wd, x in zip(wdg, _vec(v)):
                    wd.kva = x
            elif k == "taps":
                for wd, x in zip(wdg, _vec(v)):
                    wd.tap = x
            elif k == "%loadloss":
                for wd in wdg:
                    wd.pct_r = _num(v) / 2.0
            elif k in ("xht", "xlt"):
                pass                                   # irrelevant for two windings
            else:
                raise FeederError(f"unsupported Transformer property {k!r}")
        z_pu = complex((wdg[0].pct_r + wdg[1].pct_r) / 100.0, xhl / 100.0)
        y_1v = (wdg[0].kva * 1000.0 / nph) / z_pu       # leakage admittance on a 1 V base
        volts, terms = [], []
        for wd in wdg:
            v = wd.kv * 1000.0
            if nph > 1 and not wd.delta:
                v /= math.sqrt(3.0)
            volts.append(v * wd.tap)
            terms.append(Terminal.parse(wd.bus, nph))
        for ph in range(nph):
            idx, turns = [], []
            for wd, t, v in zip(wdg, terms, volts):
                if wd.delta:
                    a, b = t.nodes[ph], t.nodes[(ph + 1) % len(t.nodes)]
                else:
                    a = t.nodes[ph]
                    b = t.nodes[nph] if len(t.nodes) > nph else 0   # wye neutral -> ground
                idx += [net.node(t.bus, a), net.node(t.bus, b)]
                turns.append(1.0 / v)
            g = np.array([turns[0], -turns[0], -turns[1], turns[1]])
            net.add(idx, y_1v * np.outer(g, g))

    def _line(self, net: _Network, props, omega: float):
        nph, bus1, bus2, code = 3, "", "", None
        length, units = 1.0, "none"
        seq = dict(r1=0.058, x1=0.1206, r0=0.1784, x0=0.4047, c1=3.4, c0=1.6)
        for k, v in props:
            if k == "phases":
                nph = int(_num(v))
            elif k == "bus1":
                bus1 = v
            elif k == "bus2":
                bus2 = v
            elif k == "linecode":
                code = v.lower()
            elif k == "length":
                length = _num(v)
            elif k == "units":
                units = v.lower()
            elif k == "switch":
                if _truthy(v):                          # a switch is a short, fixed dummy section
                    seq.update(r1=1.0, x1=1.0, r0=1.0, x0=1.0, c1=1.1, c0=1.0)
                    length, units = 0.001, "none"
            elif k in seq:
                seq[k] = _num(v)
            else:
                raise FeederError(f"unsupported Line property {k!r}")
        if code is not None:
            if code not in self.linecodes:
                raise FeederError(f"unknown line code {code!r}")
            lc = self.linecodes[code]
            if lc.n != nph:
                raise FeederError(f"line code {code!r} has {lc.n} phases, line has {nph}")
            r, x, c = lc.r, lc.x, lc.c
            if units in _LEN_M and lc.units in _LEN_M:
                length *= _LEN_M[units] / _LEN_M[lc.units]
        else:
            r = _seq_matrix(seq["r1"], seq["r0"], nph)
            x = _seq_matrix(seq["x1"], seq["x0"], nph)
            c = _seq_matrix(seq["c1"], seq["c0"], nph)
        ys = np.linalg.inv((r + 1j * x) * length)
        yc_half = 0.5j * omega * c * 1e-9 * length
        t1, t2 = Terminal.parse(bus1, nph), Terminal.parse(bus2, nph)
        i1 = [net.node(t1.bus, n) for n in t1.nodes[:nph]]
        i2 = [net.node(t2.bus, n) for n in t2.nodes[:nph]]
        net.add(i1 + i2, np.block([[ys, -ys], [-ys, ys]]))
        zero = np.zeros_like(yc_half)
        net.add(i1 + i2, np.block([[yc_half, zero], [zero, yc_half]]), shunt=True)

    def _capacitor(self, net: _Network, props):
        nph, bus1, kvar, kv, delta = 3, "", 1200.0, 12.47, False
        for k, v in props:
            if k == "bus1":
                bus1 = v
            elif k == "phases":
                nph = int(_num(v))
            elif k == "kvar":
                kvar = _num(v)
            elif k == "kv":
                kv = _num(v)
            elif k == "conn":
                delta = _is_delta(v)
            else:
                raise FeederError(f"unsupported Capacitor property {k!r}")
        t = Terminal.parse(bus1, nph)
        if delta:
            vph = kv * 1000.0
            legs = [(t.nodes[k], t.nodes[(k + 1) % nph]) for k in range(nph)]
        else:
            vph = kv * 1000.0 / (math.sqrt(3.0) if nph > 1 else 1.0)
            legs = [(t.nodes[k], 0) for k in range(nph)]
        b = kvar * 1000.0 / nph / vph ** 2
        for p, q in legs:
            net.add([net.node(t.bus, p), net.node(t.bus, q)],
                    1j * b * np.array([[1.0, -1.0], [-1.0, 1.0]]), shunt=True)

    def _load(self, name: str, props) -> LoadDef:
        ld = LoadDef(name=name)
        for k, v in props:
            if k == "bus1":
                ld.bus1 = v
            elif k == "phases":
                ld.phases = int(_num(v))
            elif k == "conn":
                ld.delta = _is_delta(v)
            elif k == "model":
                ld.model = int(_num(v))
            elif k in ("kv", "kw", "kvar", "vminpu", "vmaxpu"):
                setattr(ld, k, _num(v))
            else:
                raise FeederError(f"unsupported Load property {k!r}")
        if ld.model not in (1, 2, 5):
            raise FeederError(f"load model {ld.model} is not implemented (1, 2, 5 are)")
        return ld


_CACHE: Dict[str, CompiledFeeder] = {}


def compile_feeder(feeder_file: str) -> CompiledFeeder:
    """``feeder_file`` as the reference takes it: a path relative to the packaged DSS data
    (e.g. ``ieee_13_dss/IEEE13Nodeckt.dss``), or any readable file path."""
    packaged = os.path.join(os.path.dirname(os.path.abspath(assets.__file__)), "data", "feeders",
                            feeder_file)
    if not os.path.isfile(feeder_file) and os.path.isfile(packaged):
        feeder_file = packaged                      # e.g. "synthetic123.dss"
    key = os.path.abspath(feeder_file) if os.path.isfile(feeder_file) else feeder_file
    if key not in _CACHE:
        if os.path.isfile(feeder_file):
            def reader(p):
                if os.path.isfile(p):
                    with open(p, "r", errors="replace") as fh:
                        return fh.read()
                # a redirected file that is not next to the script: the packaged IEEE data
                return assets.dss_text_by_basename(os.path.basename(p))
        else:
            reader = assets.dss_text
        _CACHE[key] = FeederCompiler(reader).run(feeder_file).compile()
    return _CACHE[key]
