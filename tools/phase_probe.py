#!/usr/bin/env python
"""Phase breakdown of pf_tc2_kernel from SM-clock stamps (instrumented build of the library).

    PGW_EXTRA_FLAGS=-DPGW_PHASE_TIMERS PGW_OUT=tools/_build/libpgw_b200_phases.so \
        PGW_BUILD_DIR=tools/_build/obj bash powergridworld_b200/csrc/build.sh     # here (no GPU needed)
    gpurun -- python tools/phase_probe.py [c1|c3] [envs]

Every CTA's thread 0 stamps clock64() at the boundaries of its first tile; printed: mean and max
over CTAs of each phase, in microseconds at the SM clock nvidia-smi reports, cold L2 (256 MiB
write before each step) and warm."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from powergridworld_b200 import _native as N  # noqa: E402

N.LIB_PATH = os.path.join(ROOT, "tools", "_build", "libpgw_b200_phases.so")
from powergridworld_b200.scenarios import catalog as S  # noqa: E402
from powergridworld_b200.scenarios.namespace import PRODUCT_NS as NS  # noqa: E402

PHASES = ["clock read", "barrier init + TMA issue + prefetch + TMEM alloc", "wait tables/event row",
          "tile inputs + first currents + A", "first chain issue", "fixed-point loop",
          "branch state (u_state, derived vmag)", "expansion + node magnitudes", "rewards / bus voltages",
          "dealloc + clock publish"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
    polish = int(os.environ.get("PGW_POLISH", "-1"))
    E = int(sys.argv[2]) if len(sys.argv) > 2 else (4096 if wl == "c1" else 16384)
    import warnings
    warnings.simplefilter("ignore")
    if wl == "c1":
        env = NS.CoordinatedMultiBuildingControlEnv(
            **S.buildings_scenario(NS, NS.OpenDSSSolver, 1.2), num_envs=E, pf_kernel="tc2")
    else:
        env = NS.MultiAgentEnv(**S.der123_scenario(NS, NS.OpenDSSSolver), num_envs=E, pf_kernel="tc2")
    if polish >= 0:
        env.set_option(N.OPT_PF_POLISH, polish)
    lib = env._lib
    lib.pgw_debug_phases.restype = C.c_int
    lib.pgw_debug_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True).stdout.split()[0])
    ctas = min((E + 127) // 128, 4096)
    rng = np.random.default_rng(0)
    soc = rng.uniform(10, 45, size=(env.num_storage, E))
    acts = torch.as_tensor(rng.uniform(-1, 1, size=(env.act_dim, E))).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib.pgw_debug_comp_span.restype = C.c_int
    lib.pgw_debug_comp_span.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.pgw_debug_num_comp_ctas.restype = C.c_int
    lib.pgw_debug_num_comp_ctas.argtypes = [C.c_void_p]
    nc = lib.pgw_debug_num_comp_ctas(env._h)
    for pdl in (1, 0):
        env.set_option(N.OPT_PDL, pdl)
        env.reset_batch(soc)
        spans, cphase = [], []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev_ms = []
        for t in range(60):
            flush.fill_(t & 0xFF)
            e0.record()
            env.step_batch(acts)
            e1.record()
            if t >= 20:
                pf = np.zeros((ctas, 16), dtype=np.int64)
                cp = np.zeros((nc, 8), dtype=np.int64)
                N.check(lib.pgw_debug_phases(env._h, pf.ctypes.data_as(C.c_void_p), ctas))
                N.check(lib.pgw_debug_comp_span(env._h, cp.ctypes.data_as(C.c_void_p), nc))
                t0 = cp[:, 0].min()
                spans.append([cp[:, 1].max() - t0, pf[:, 12].min() - t0, pf[:, 13].max() - t0,
                              pf[:, 12].max() - t0])
                ev_ms.append(e0.elapsed_time(e1) * 1e3)
                cphase.append(np.diff(cp[:, 2:7], axis=1) / mhz)
        sp = np.array(spans, dtype=np.float64).mean(axis=0) / 1e3
        print(f"--- {wl} E={E} PDL={pdl} cold L2 (globaltimer, us from the first component CTA's entry): "
              f"components end {sp[0]:.2f}, power flow first entry {sp[1]:.2f} / last entry {sp[3]:.2f}, "
              f"power flow end {sp[2]:.2f}; step by CUDA events {np.mean(ev_ms):.2f} us")
        cph = np.stack(cphase).mean(axis=(0, 1))
        print("    component kernel CTA phases (mean, us): entry -> TMA issued + barrier %.2f | static slice landed %.2f | "
              "event row landed (first block's prefetches issued) %.2f | agent steps of all blocks %.2f" % tuple(cph))
    env.set_option(N.OPT_PDL, 1)
    for cold in (True, False):
        env.reset_batch(soc)
        acc = []
        for t in range(60):
            if cold:
                flush.fill_(t & 0xFF)
            env.step_batch(acts)
            if t >= 20:
                buf = np.zeros((ctas, 16), dtype=np.int64)
                N.check(lib.pgw_debug_phases(env._h, buf.ctypes.data_as(C.c_void_p), ctas))
                acc.append(buf)
        a = np.stack(acc).astype(np.float64)                 # [steps, ctas, 16]
        d = np.diff(a[:, :, :11], axis=2) / mhz               # us
        print(f"--- {wl} E={E} ctas={ctas} {'cold' if cold else 'warm'} L2, SM {mhz:.0f} MHz, "
              f"iterations of the stamped tile: mean {a[:, :, 11].mean():.2f}")
        for k, name in enumerate(PHASES):
            print(f"  {name:52s} mean {d[:, :, k].mean():7.2f} us   max-CTA mean {d[:, :, k].max(axis=1).mean():7.2f} us")
        if (a[:, :, 14] > 0).all():                              # FP64 polish stamps (inside "branch state")
            pol = np.stack([a[:, :, 14] - a[:, :, 6], a[:, :, 15] - a[:, :, 14], a[:, :, 7] - a[:, :, 15]], axis=2) / mhz
            for k, name in enumerate(["  polish: chains + u64 + currents of sweep 0", "  polish: full sweep 0 (matvec)",
                                      "  polish: remaining sweeps + magnitudes"]):
                print(f"  {name:52s} mean {pol[:, :, k].mean():7.2f} us   max-CTA mean {pol[:, :, k].max(axis=1).mean():7.2f} us")
        tot = (a[:, :, 10] - a[:, :, 0]) / mhz
        print(f"  {'entry -> exit':52s} mean {tot.mean():7.2f} us   max-CTA mean {tot.max(axis=1).mean():7.2f} us")


if __name__ == "__main__":
    main()
