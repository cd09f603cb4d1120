"""GPU probe: the fused step kernel (PGW_OPT_FUSED) against the two-kernel path and the FP64 solver on
the C1 scenario: observations bit for bit, rewards / voltages within the float64 bounds, step time."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from powergridworld_b200 import _native as N                                   # noqa: E402
from powergridworld_b200.scenarios import bench as SB                          # noqa: E402


def run(E, T, fused, polish=None, timing=True):
    rng = np.random.default_rng(3)
    env = SB.c1_env(num_envs=E)
    env.set_option(N.OPT_PF_KERNEL, 2)
    env.set_option(N.OPT_FUSED, fused)
    if polish is not None:
        env.set_option(N.OPT_PF_POLISH, polish)
    soc = rng.uniform(5, 45, size=(env.num_storage, E))
    acts = [torch.as_tensor(rng.uniform(-1, 1, size=(env.act_dim, E))).cuda() for _ in range(T)]
    out = {"obs": [], "rew": [], "v": [], "it": []}
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        out["obs0"] = env.reset_batch(soc).clone()
        for a in acts:
            o, r, d, _ = env.step_batch(a)
            out["obs"].append(o.clone()); out["rew"].append(r.clone())
            out["v"].append(env.get_field(N.FIELD_VOLTAGES).clone())
            out["it"].append(env.get_field(N.FIELD_PF_ITERS).clone())
        out["ep"] = env.get_field(N.FIELD_EP_RETURN).clone()
        out["sd"] = env.get_field(N.FIELD_STATE_D).clone()
        out["ap"] = env.get_field(N.FIELD_AGENT_P).clone()
        us = None
        if timing:
            env.reset_batch(soc)
            for i in range(10):
                env.step_batch(acts[i % 4])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                env.step_batch(acts[i % 4])
            e1.record()
            st.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / 200
    st.synchronize()
    env.close()
    return out, us


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = 40
    rng = np.random.default_rng(3)
    ref = SB.c1_env(num_envs=E, pf_tol=1e-13, pf_max_iter=200)
    soc = rng.uniform(5, 45, size=(ref.num_storage, E))
    acts = [torch.as_tensor(rng.uniform(-1, 1, size=(ref.act_dim, E))).cuda() for _ in range(T)]
    ref.reset_batch(soc)
    R, V = [], []
    for a in acts:
        _, r, _, _ = ref.step_batch(a)
        R.append(r.clone()); V.append(ref.get_field(N.FIELD_VOLTAGES).clone())
    base, us0 = run(E, T, 0)
    for fused, polish in ((2, None), (2, 0), (2, 2)):
        got, us = run(E, T, fused, polish)
        rec = {"E": E, "fused": fused, "polish": polish, "us_per_step_warm": us, "two_kernel_us": us0,
               "obs_bit_identical": all(bool((a == b).all()) for a, b in zip(got["obs"], base["obs"])),
               "state_bit_identical": bool((got["sd"] == base["sd"]).all()),
               "agent_p_bit_identical": bool((got["ap"] == base["ap"]).all()),
               "max_rew_err_vs_fp64": max(float((a - b).abs().max()) for a, b in zip(got["rew"], R)),
               "max_v_err_vs_fp64": max(float((a - b).abs().max()) for a, b in zip(got["v"], V)),
               "max_rew_diff_vs_two_kernel": max(float((a - b).abs().max()) for a, b in zip(got["rew"], base["rew"])),
               "mean_iters": float(torch.stack(got["it"]).abs().double().mean()),
               "mean_iters_two_kernel": float(torch.stack(base["it"]).abs().double().mean()),
               "min_iters": int(torch.stack(got["it"]).min())}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
