#!/usr/bin/env python
"""Author a 123-bus-CLASS radial test feeder as an OpenDSS script.

The reference ships only the IEEE-13 circuit; `IEEELineCodes.dss` merely carries the line
codes 1-12 of the IEEE 123-node feeder.  BASELINE configs C3/C4 need a feeder of that size,
so this script *authors* one (it is NOT the IEEE 123-node data set): a deterministic
pseudo-random radial tree with the same voltage level (115/4.16 kV, 5 MVA substation), the
same conductor data (line codes 1-12, lengths in kft), 123 buses, a three-phase trunk with
one-/two-/three-phase laterals and 85 single-phase wye spot loads (20/40 kW at 0.9 pf, like
the IEEE case; a few constant-Z and constant-I ones), about 2.5 MW in total.

    python tools/make_synthetic123.py   ->  powergridworld_b200/data/feeders/synthetic123.dss
"""
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "powergridworld_b200",
                   "data", "feeders", "synthetic123.dss")
THREE = [1, 2, 3, 4, 5, 6]          # three-phase overhead configurations
TWO, ONE = [7, 8], [9, 10, 11]      # two-phase (phases 1.3 / 1.2), single-phase


def main():
    rng = np.random.default_rng(123)
    lines, loads = [], []
    bus_phases = {"150": [1, 2, 3]}              # substation secondary
    trunk = ["150"]
    nbus = 0

    def new_bus():
        nonlocal nbus
        nbus += 1
        return str(nbus)

    def add_line(a, b, phases, kft):
        code = {3: rng.choice(THREE), 2: None, 1: rng.choice(ONE)}[len(phases)]
        if len(phases) == 2:
            code = 7 if phases == [1, 3] else 8
            if phases == [2, 3]:
                code = 8
        nodes = "." + ".".join(str(p) for p in phases)
        lines.append(f"New Line.L{len(lines) + 1} Phases={len(phases)} Bus1={a}{nodes} Bus2={b}{nodes} "
                     f"LineCode={code} Length={kft:.3f} units=kft")
        bus_phases[b] = list(phases)

    # three-phase trunk with branching
    frontier = ["150"]
    while nbus < 44:
        parent = frontier[int(rng.integers(len(frontier)))]
        b = new_bus()
        add_line(parent, b, [1, 2, 3], float(rng.uniform(0.15, 0.55)))
        trunk.append(b)
        frontier.append(b)
        if len(frontier) > 6:
            frontier.pop(0)
    # laterals
    while nbus < 122:
        root = trunk[int(rng.integers(4, len(trunk)))]
        kind = rng.random()
        phases = [int(rng.integers(1, 4))] if kind < 0.7 else \
            [[1, 2], [1, 3], [2, 3]][int(rng.integers(3))]
        prev = root
        for _ in range(int(rng.integers(1, 5))):
            if nbus >= 122:
                break
            b = new_bus()
            add_line(prev, b, phases, float(rng.uniform(0.15, 0.55)))
            prev = b
    # 85 spot loads on distinct (bus, phase) pairs
    spots = [(b, p) for b, ph in bus_phases.items() if b != "150" for p in ph]
    rng.shuffle(spots)
    used_bus = set()
    for b, p in spots:
        if len(loads) >= 85:
            break
        if b in used_bus:
            continue
        used_bus.add(b)
        kw = 40.0 if rng.random() < 0.55 else 20.0
        model = 1 if len(loads) % 9 else (2 if len(loads) % 2 else 5)
        loads.append(f"New Load.{b}{'abc'[p - 1]} Bus1={b}.{p} Phases=1 Conn=Wye Model={model} "
                     f"kV=2.4 kW={kw:.1f} kvar={kw / 2:.1f}")
    total = sum(float(l.split("kW=")[1].split()[0]) for l in loads)

    text = f"""! 123-bus-class synthetic radial feeder, authored by tools/make_synthetic123.py for
! powergridworld_b200 (BASELINE configs C3/C4).  NOT the IEEE 123-node data set: the
! reference ships no such circuit, only its line codes (IEEELineCodes.dss, codes 1-12).
! {nbus + 1} buses, {len(lines)} line sections, {len(loads)} spot loads, {total:.0f} kW.
Clear
Set DefaultBaseFrequency=60

New Circuit.synthetic123 basekv=115 pu=1.0 phases=3 bus1=SourceBus Angle=30 MVAsc3=20000 MVAsc1=21000

New Transformer.Sub Phases=3 Windings=2 XHL=2
~ wdg=1 bus=SourceBus conn=delta kv=115  kva=5000 %r=0.25
~ wdg=2 bus=150       conn=wye   kv=4.16 kva=5000 %r=0.25

redirect IEEELineCodes.dss

""" + "\n".join(lines) + "\n\n" + "\n".join(loads) + """

Set Voltagebases=[115, 4.16]
calcv
Solve
"""
    with open(OUT, "w") as fh:
        fh.write(text)
    print("wrote", os.path.abspath(OUT), f"{nbus + 1} buses, {len(lines)} lines, {len(loads)} loads, {total:.0f} kW")


if __name__ == "__main__":
    main()
