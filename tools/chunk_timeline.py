"""GPU probe (instrumented build): globaltimer entry / exit of every CTA of the chunked zero-copy host step."""
import ctypes as C
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from powergridworld_b200 import _native as N  # noqa: E402

N.LIB_PATH = os.path.join(ROOT, "tools", "_build", "libpgw_b200_phases.so")
from powergridworld_b200.scenarios import bench as SB  # noqa: E402

E = 4096
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for chunks in (1, 4):
    env = SB.c1_env(num_envs=E, pf_kernel="tc2")
    env.set_option(N.OPT_HOST_CHUNKS, chunks)
    lib = env._lib
    lib.pgw_debug_phases.restype = C.c_int
    lib.pgw_debug_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    soc = np.full((env.num_storage, E), 30.0)
    acts = [torch.rand((env.act_dim, E), dtype=torch.float64).mul_(2).sub_(1).pin_memory() for _ in range(4)]
    env.reset_host(soc)
    rows = []
    for i in range(30):
        env.step_host(acts[i % 4])
        if i >= 10:
            buf = np.zeros((128, 16), dtype=np.int64)
            N.check(lib.pgw_debug_phases(env._h, buf.ctypes.data_as(C.c_void_p), 128))
            t0 = buf[:, 12].min()
            per = 128 // chunks
            rows.append([[(buf[k * per:(k + 1) * per, 12].min() - t0) / 1e3, (buf[k * per:(k + 1) * per, 12].max() - t0) / 1e3,
                          (buf[k * per:(k + 1) * per, 13].min() - t0) / 1e3, (buf[k * per:(k + 1) * per, 13].max() - t0) / 1e3]
                         for k in range(chunks)])
    r = np.array(rows).mean(axis=0)
    print(f"chunks={chunks}: per chunk [first entry, last entry, first exit, last exit] us from the first CTA's entry")
    for k in range(chunks):
        print("   chunk %d: %7.2f %7.2f %7.2f %7.2f" % (k, *r[k]))
    env.close()
