"""TEST INFRASTRUCTURE ONLY: drive the host build of the device arithmetic
(tests/emu/emu.cpp) with the tables of a ``_dry_run`` product env, with a NumPy
Z-bus fixed point standing in for csrc/powerflow.cu.  Lets the CPU-only test tier
check spec compiler + component arithmetic against the reference golden traces."""
import ctypes as C
import os
import subprocess

import numpy as np

from powergridworld_b200 import _native as N

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpgw_emu.so")


def build():
    src = os.path.join(HERE, "emu.cpp")
    hdr = os.path.join(HERE, "..", "..", "powergridworld_b200", "csrc", "component_math.cuh")
    if not os.path.isfile(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC",
                               "-x", "c++", "-o", SO, src])
    return C.CDLL(SO)


class EmuArgs(C.Structure):
    _fields_ = [("E", C.c_int), ("A", C.c_int)] + [(n, C.c_void_p) for n in (
        "agents", "comps", "dpar", "ipar", "drow", "irow", "actions", "obs", "rew", "agent_p",
        "sd", "si", "init_soc")] + [("clip_init_soc", C.c_int)] + [
        (n, C.c_void_p) for n in ("vmin", "vmax", "vbus")]


def zbus_solve(f, kw, kvar, tol=1e-12, max_iter=200):
    """NumPy statement of the fixed point the CUDA kernel runs (per unit)."""
    s = (kw[f.branch_load] + 1j * kvar[f.branch_load]) * f.branch_share * 1e-3
    u = f.u0.copy()
    for _ in range(max_iter):
        m = np.abs(u)
        us = np.where(u == 0, 1, u)
        i = np.where(m <= f.branch_vmin, np.conj(s) / f.branch_vmin ** 2 * u,
                     np.where(m > f.branch_vmax, np.conj(s) / f.branch_vmax ** 2 * u,
                              np.conj(s / us)))                      # model 1: PQ with Z band
        i = np.where(f.branch_model == 2, np.conj(s) * u, i)         # constant impedance
        i = np.where(f.branch_model == 5, np.conj(s) * us / np.abs(us), i)   # constant |I|
        un = f.u0 - f.zbb @ i
        d = np.abs(un - u).max() if len(u) else 0.0
        u = un
        if d < tol:
            break
    return np.abs(f.w - f.znb @ i)


class EmulatedEnv:
    def __init__(self, env):
        self.lib = build()
        self.env = env
        b = env._b
        self.E, self.A = env.num_envs, len(env.agents)
        self.comps = (N.Component * len(b.comps))(*b.comps)
        self.dpar = env._dpar          # shared: roster re-draws rewrite it in place
        self.ipar = np.asarray(b.ipar if b.ipar else [0], dtype=np.int32)
        self.sd = np.zeros((max(b.sd_rows, 1), self.E))
        self.si = np.zeros((max(b.si_rows, 1), self.E), dtype=np.uint32)
        self.obs = np.zeros((env.obs_dim, self.E))
        self.rew = np.zeros((self.A, self.E))
        self.agent_p = np.zeros((self.A, self.E))
        self.vmin = np.ones(self.E)
        self.vmax = np.ones(self.E)
        self.vbus = np.ones((self.A, self.E))
        self.vmag = None
        self.t = 0
        self.resets = 0

    def _args(self, event, actions=None, init_soc=None, drawn=False):
        a = EmuArgs()
        a.clip_init_soc = 0 if drawn else 1
        a.E, a.A = self.E, self.A
        p = lambda x: x.ctypes.data if x is not None else None
        a.agents = C.addressof(self.env._agent_recs)
        a.comps = C.addressof(self.comps)
        a.dpar, a.ipar = p(self.dpar), p(self.ipar)
        self._drow = np.ascontiguousarray(self.env._dtab[event])
        self._irow = np.ascontiguousarray(self.env._itab[event]) if self.env._istride else np.zeros(4, np.int32)
        a.drow, a.irow = p(self._drow), p(self._irow)
        a.actions, a.obs, a.rew, a.agent_p = p(actions), p(self.obs), p(self.rew), p(self.agent_p)
        a.sd, a.si, a.init_soc = p(self.sd), p(self.si), p(init_soc)
        a.vmin, a.vmax, a.vbus = p(self.vmin), p(self.vmax), p(self.vbus)
        return a

    def _powerflow(self, event, controllable):
        env = self.env
        if env.pf_solver is None:
            return
        f = env.pf_solver.feeder
        nl = f.nl
        row = env._dtab[event]
        self.vmag = np.zeros((f.nn, self.E))
        for e in range(self.E):
            kw, kvar = row[2:2 + nl].copy(), row[2 + nl:2 + 2 * nl].copy()
            if controllable:
                add = {}
                for ai in range(self.A):
                    l = env._agent_recs[ai].load_slot
                    if l < 0:
                        continue
                    add[l] = add[l] + self.agent_p[ai, e] if l in add else self.agent_p[ai, e]
                for l, v in add.items():
                    kw[l] += v
            self.vmag[:, e] = zbus_solve(f, kw, kvar)
        self.vmin, self.vmax = self.vmag.min(axis=0).copy(), self.vmag.max(axis=0).copy()
        for ai in range(self.A):
            n = env._agent_recs[ai].bus_node
            self.vbus[ai] = self.vmag[n] if n >= 0 else 1.0

    def reset(self, init_soc, drawn=False):
        """``drawn``: the SOCs are the reference's own (unclipped) draw, not an explicit
        init_storage (PGW_OPT_CLIP_INIT_SOC = 0)."""
        self._powerflow(0, controllable=False)
        for field, row0, arr in self.env._per_env_roster_rows():   # what pgw_set_rows does on the device
            dst = self.sd if field == N.FIELD_STATE_D else self.si
            dst[row0:row0 + arr.shape[0]] = arr
        soc = np.ascontiguousarray(init_soc, dtype=np.float64) if init_soc is not None else None
        self.lib.emu_set_first_reset(1 if self.resets == 0 else 0)
        self.resets += 1
        self.lib.emu_reset(C.byref(self._args(0, init_soc=soc, drawn=drawn)))
        self.t = 0
        return self.obs.copy()

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        event = self.t + 1
        self.lib.emu_step(C.byref(self._args(event, actions=actions)))
        self._powerflow(event, controllable=True)
        pen = self.env._penalty
        if pen is not None:
            node = self.env.pf_solver.node_for_bus_name(
                str(list(set(self.env.agent_name_bus_map.values()))[0]))
            v = self.vmag[node]
            viol = np.maximum(0.0, np.maximum(pen[0] - v, v - pen[1]))
            self.rew -= (viol * pen[2]) / self.A
        self.t += 1
        done = self.env._dtab[event, 0] != 0.0
        return self.obs.copy(), self.rew.copy(), done
