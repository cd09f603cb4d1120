#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE -- installs the UNMODIFIED reference into ``baseline/_ref``
(git-ignored; it travels to the GPU box with the snapshot) so that ``bench.py --impl reference`` can
time the reference's own classes there.

    python oracle/install_reference.py [/root/reference]

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the tree>``
   (the tree is read-only, so the wheel is built from a copy under /tmp; ``--no-deps`` because
   ``gym`` / ``opendssdirect`` are not in the offline wheelhouse -- oracle/ref_harness.py stubs the
   former and plugs the power-flow port in for the latter).
2. The wheel carries no package data (the reference's setup.py lists none), so the data files
   the classes read at run time (PV profiles, vehicle table, building model, feeder scripts) are
   copied next to the installed modules, as an editable install would find them.
Nothing under baseline/_ref is tracked by git; no reference source enters the repository."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA_EXT = (".csv", ".p", ".dss", ".json", ".txt", ".pkl")


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    if not os.path.isdir(os.path.join(src, "gridworld")):
        print(f"no reference tree at {src}: nothing installed")
        return 1
    dst = os.path.join(ROOT, "baseline", "_ref")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(src, copy)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index",
                               "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
                               "--target", dst, copy])
    n = 0
    for base, _, files in os.walk(os.path.join(src, "gridworld")):
        for f in files:
            if f.endswith(DATA_EXT):
                rel = os.path.relpath(os.path.join(base, f), src)
                os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
                shutil.copy2(os.path.join(base, f), os.path.join(dst, rel))
                n += 1
    print(f"installed the reference into {dst} (+ {n} data files)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
