"""Distribution of power-flow iteration counts per env and per 128-env tile (C3 workload)."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from powergridworld_b200.scenarios import catalog as S
from powergridworld_b200.scenarios.namespace import PRODUCT_NS as NS
from powergridworld_b200 import _native as N
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
k = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    env = NS.MultiAgentEnv(**S.der123_scenario(NS, NS.OpenDSSSolver), num_envs=E)
env.set_option(N.OPT_PF_KERNEL, k)
rng = np.random.default_rng(0)
soc = torch.as_tensor(30 + 5 * rng.uniform(-1, 1, size=(env.num_storage, E))).cuda()
env.reset_batch(soc)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for t in range(12):
    act = torch.rand((env.act_dim, E), generator=gen, device="cuda", dtype=torch.float64) * 2 - 1
    env.step_batch(act)
    it = env.get_field(7).abs().cpu().numpy()
    tiles = it[: E // 128 * 128].reshape(-1, 128).max(axis=1)
    print(f"t={t} env mean {it.mean():.2f} max {it.max()}  tile-max mean {tiles.mean():.2f} min {tiles.min()} max {tiles.max()}  hist {np.bincount(it)[:20]}")
