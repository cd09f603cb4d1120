"""Drop-in for the reference's ``OpenDSSSolver``
(gridworld/distribution_system/opendss.py:15-186) that needs no OpenDSS engine.

Same constructor keywords and the same result accessors; the snapshot solve is
the batched Z-bus fixed point of csrc/powerflow.cu.  Inside ``MultiAgentEnv`` the
solver is not called from Python at all -- the env compiles ``self.feeder`` and
the load shape into device tables and the solve runs as part of the step.  Used
on its own (as in the reference's tests/distribution_system/test_opendss.py) it
owns a one-env device handle and goes through ``pgw_pf_solve``.
"""
from datetime import datetime
from typing import List, Union

import numpy as np
import pandas as pd

from powergridworld_b200 import assets
from powergridworld_b200.distribution_system.feeder import compile_feeder
from powergridworld_b200.distribution_system.powerflow import PowerFlowSolver


def hour_of_year(t) -> int:
    t = pd.Timestamp(t)
    return int((t - datetime(t.year, 1, 1)).total_seconds() // 3600)        # opendss.py:98-105


class ZBusSolver(PowerFlowSolver):

    def __init__(self, feeder_file: str, loadshape_file: str,
                 system_load_rescale_factor: float = 1.0, tol: float = 1e-9,
                 max_iter: int = 60, **kwargs):
        super().__init__(**kwargs)
        self.feeder_file = feeder_file
        self.feeder = compile_feeder(feeder_file)
        self.system_load_rescale_factor = system_load_rescale_factor
        import os
        if os.path.isfile(loadshape_file):
            self.annual_hourly_load_profile = np.genfromtxt(loadshape_file)
        else:
            self.annual_hourly_load_profile = assets.array("loadshape/" + loadshape_file)
        if len(self.annual_hourly_load_profile) != 8760:
            print("Warning: The provided load shape file is not annual hourly ",
                  "profile. Error might occur later")
        self.tol = tol
        self.max_iter = max_iter
        f = self.feeder
        pq = f.load_model == 1                                               # :54-77
        self.load_bus_name = [n for n, m in zip(f.load_names, pq) if m]
        self.base_load = np.stack([f.load_kw[pq], f.load_kvar[pq]], axis=1)
        self._pq_mask = pq
        self.bus_voltages = {}
        self._env = None           # the MultiAgentEnv that owns this solver (set by it), if any
        self._host_env = None      # private one-env handle behind a stand-alone calculate_power_flow

    # ---- table compilation used by MultiAgentEnv
    def base_load_at(self, current_time):
        """Total kW / kvar of every load (definition order) at ``current_time`` with no
        controllable injection (:106-108); non-PQ loads keep their nominal values."""
        f = self.feeder
        coeff = self.annual_hourly_load_profile[hour_of_year(current_time)]
        kw, kvar = f.load_kw.copy(), f.load_kvar.copy()
        scaled = coeff * self.base_load * self.system_load_rescale_factor
        kw[self._pq_mask] = scaled[:, 0]
        kvar[self._pq_mask] = scaled[:, 1]
        return kw, kvar

    # ---- PowerFlowSolver protocol
    def calculate_power_flow(self, p_controllable_consumed: dict = None,
                             q_controllable_consumed: dict = None, current_time: str = None):
        from powergridworld_b200.multiagent_env import _standalone_solver_env
        if self._env is not None:
            # the owning env solves on the device inside reset / step; a second, private solve
            # here would leave two sets of voltages behind one accessor
            raise RuntimeError("this solver belongs to a MultiAgentEnv: its power flow runs inside "
                               "env.reset() / env.step(); build a separate OpenDSSSolver for "
                               "stand-alone solves")
        if self._host_env is None:
            self._host_env = _standalone_solver_env(self)
        kw, kvar = self.base_load_at(current_time)
        if p_controllable_consumed is not None:
            for name in self.load_bus_name:                                  # :115-129
                i = self.feeder.load_index(name)
                kw[i] += p_controllable_consumed.get(name, 0.0)
                kvar[i] += (q_controllable_consumed or {}).get(name, 0.0)
        self._host_env._solve_loads(kw, kvar)
        self.bus_voltages = self._host_env._voltage_dict()

    def get_bus_voltages(self) -> dict:
        if self._env is not None:
            self.bus_voltages = self._env._voltage_dict()
        return self.bus_voltages

    def get_bus_voltage_by_name(self, bus_name: any) -> Union[float, List[float]]:
        v = self.get_bus_voltages()
        phase_map = {'a': '.1', 'b': '.2', 'c': '.3'}
        if bus_name[-1] in phase_map:                                        # :177-181
            return v[bus_name.replace(bus_name[-1], phase_map[bus_name[-1]])]
        return [v[x] for x in [bus_name + p for p in phase_map.values()]]

    def node_for_bus_name(self, bus_name: str):
        """Node index behind ``get_bus_voltage_by_name`` for a single-phase name, else None."""
        phase_map = {'a': '.1', 'b': '.2', 'c': '.3'}
        if bus_name[-1] in phase_map:
            return self.feeder.node_index(bus_name.replace(bus_name[-1], phase_map[bus_name[-1]]))
        return None


class OpenDSSSolver(ZBusSolver):
    """The reference's class name (gridworld/distribution_system/opendss.py:18), so that scenario
    configs read -- and print -- the same."""
