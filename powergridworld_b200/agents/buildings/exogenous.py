"""Exogenous (weather / gains) table of the five-zone building.

The reference reads ``gridworld/agents/buildings/data/exogenous_data.csv``
(five_zone_rom_env.py:35), a file that is absent from the reference checkout
(``.MISSING_LARGE_BLOBS``).  ``load_exogenous`` therefore accepts a user-supplied
CSV of the same shape (datetime index in the first column; columns selected by the
prefixes ``T_oa``, ``Q_solar``, ``Q_cool_``, ``Q_int``, :140-144) and otherwise
falls back to a deterministic synthetic table: T_oa = 25 + 5 sin(2 pi k / 288),
Q_solar, Q_int ~ U(0, 1), Q_cool ~ -U(0, 5) from ``numpy.random.default_rng(0)``,
5-minute rows from 2020-08-12 00:00, 577 rows.
"""
import os

import numpy as np
import pandas as pd

N_ROWS = 577
START = "2020-08-12 00:00:00"
ENV_VAR = "PGW_EXOGENOUS_CSV"


def synthetic_table() -> np.ndarray:
    """[N_ROWS, 16]: T_oa, Q_solar[5], Q_cool[5], Q_int[5]."""
    rng = np.random.default_rng(0)
    k = np.arange(N_ROWS, dtype=np.float64)
    t_oa = 25.0 + 5.0 * np.sin(2.0 * np.pi * k / 288.0)
    q_solar = rng.uniform(0.0, 1.0, size=(N_ROWS, 5))
    q_cool = -rng.uniform(0.0, 5.0, size=(N_ROWS, 5))
    q_int = rng.uniform(0.0, 1.0, size=(N_ROWS, 5))
    return np.concatenate([t_oa[:, None], q_solar, q_cool, q_int], axis=1)


def load_exogenous(start_time=None, end_time=None, csv_path=None):
    """Rows between the two timestamps inclusive (``df.loc[start:end]``, :38-41)."""
    csv_path = csv_path or os.environ.get(ENV_VAR)
    if csv_path:
        df = pd.read_csv(csv_path, index_col=0)
        df.index = pd.DatetimeIndex(df.index)
        pick = lambda prefix: [c for c in df.columns if c.startswith(prefix)]
        cols = pick("T_oa")[:1] + pick("Q_solar")[:5] + pick("Q_cool_")[:5] + pick("Q_int")[:5]
        if len(cols) != 16:
            raise ValueError("exogenous CSV needs T_oa, 5x Q_solar*, 5x Q_cool_*, 5x Q_int* columns")
        df = df[cols]
    else:
        df = pd.DataFrame(synthetic_table(),
                          index=pd.date_range(START, periods=N_ROWS, freq="5min"))
    lo = pd.Timestamp(start_time) if start_time else df.index[0]
    hi = pd.Timestamp(end_time) if end_time else df.index[-1]
    out = df.loc[lo:hi]
    if len(out) == 0:
        raise ValueError(
            f"start and/or end times ({lo}, {hi}) resulted in empty dataframe.  First and last "
            f"indices are ({df.index[0]}, {df.index[-1]}), choose values in this range.")
    return out.values.astype(np.float64), out.index
