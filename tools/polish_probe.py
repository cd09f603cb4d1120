"""GPU probe: accuracy and cost of the tcgen05 solver's float64 polish (PGW_OPT_PF_POLISH) on the
C1 scenario, against the FP64 SIMT solver run to 1e-13.  Prints one JSON line per setting."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powergridworld_b200 import _native as N                                   # noqa: E402
from powergridworld_b200.scenarios import bench as SB                          # noqa: E402


def main():
    E, T = 4096, 60
    rng = np.random.default_rng(3)
    ref = SB.c1_env(num_envs=E, pf_tol=1e-13, pf_max_iter=200)
    soc = rng.uniform(5, 45, size=(ref.num_storage, E))
    acts = [torch.as_tensor(rng.uniform(-1, 1, size=(ref.act_dim, E))).cuda() for _ in range(T)]
    pn = ref.pf_solver.node_for_bus_name("675c")
    ref.reset_batch(soc)
    R, V = [], []
    for a in acts:
        _, r, _, _ = ref.step_batch(a)
        R.append(r.clone())
        V.append(ref.get_field(N.FIELD_VOLTAGES)[pn].clone())
    for polish, tol in ((0, 100), (1, 100), (2, 100), (1, 300), (2, 1000)):
        env = SB.c1_env(num_envs=E)
        env.set_option(N.OPT_PF_KERNEL, 2)
        env.set_option(N.OPT_PF_POLISH, polish)
        env.set_option(N.OPT_PF_TC_TOL_NANO, tol)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            env.reset_batch(soc)
            er = ev = 0.0
            its = 0.0
            for t, a in enumerate(acts):
                _, r, _, _ = env.step_batch(a)
                er = max(er, float((r - R[t]).abs().max()))
                ev = max(ev, float((env.get_field(N.FIELD_VOLTAGES)[pn] - V[t]).abs().max()))
                its += float(env.get_field(N.FIELD_PF_ITERS).abs().double().mean())
            # cost: graph replays, warm L2, CUDA events around 200 steps
            env.reset_batch(soc)
            for i in range(10):
                env.step_batch(acts[i % 4])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                env.step_batch(acts[i % 4])
            e1.record()
        st.synchronize()
        print(json.dumps({"polish": polish, "tc_tol": tol * 1e-9, "max_reward_err": er,
                          "max_v_err_penalty_node": ev, "mean_iters": its / T,
                          "us_per_step_warm": 1e3 * e0.elapsed_time(e1) / 200}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
