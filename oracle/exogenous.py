"""TEST INFRASTRUCTURE ONLY.

Deterministic synthetic stand-in for the reference's missing
``gridworld/agents/buildings/data/exogenous_data.csv`` (listed in the
reference's ``.MISSING_LARGE_BLOBS``).  Columns follow the regex prefixes
the reference selects by (five_zone_rom_env.py:140-144): ``T_oa``,
``Q_solar*`` (5), ``Q_cool_*`` (5), ``Q_int*`` (5).  Recipe from SURVEY.md
section 8(d)/C1: T_oa = 25 + 5 sin(2 pi k / 288), Q_solar,Q_int ~ U(0,1),
Q_cool ~ -U(0,5), numpy default_rng(0); 5-minute index from 2020-08-12 00:00,
577 rows (two days inclusive).

The product ships its own copy of this recipe
(powergridworld_b200/agents/buildings/exogenous.py); tests assert equality.
"""
import numpy as np

N_ROWS = 577
START = "2020-08-12 00:00:00"
COLUMNS = (["T_oa"] + [f"Q_solar_{z}" for z in range(5)]
           + [f"Q_cool_{z}" for z in range(5)] + [f"Q_int_{z}" for z in range(5)])


def synthetic_exogenous_table() -> np.ndarray:
    """[N_ROWS, 16] float64 in COLUMNS order."""
    rng = np.random.default_rng(0)
    k = np.arange(N_ROWS, dtype=np.float64)
    t_oa = 25.0 + 5.0 * np.sin(2.0 * np.pi * k / 288.0)
    q_solar = rng.uniform(0.0, 1.0, size=(N_ROWS, 5))
    q_cool = -rng.uniform(0.0, 5.0, size=(N_ROWS, 5))
    q_int = rng.uniform(0.0, 1.0, size=(N_ROWS, 5))
    return np.concatenate([t_oa[:, None], q_solar, q_cool, q_int], axis=1)


def synthetic_exogenous_frame():
    import pandas as pd

    idx = pd.date_range(START, periods=N_ROWS, freq="5min")
    return pd.DataFrame(synthetic_exogenous_table(), index=idx, columns=COLUMNS)
