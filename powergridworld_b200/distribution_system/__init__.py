from .opendss import OpenDSSSolver, ZBusSolver
from .powerflow import PowerFlowSolver
