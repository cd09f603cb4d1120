"""Power-flow plugin protocol (mirrors gridworld/distribution_system/powerflow.py:7-51)."""
from abc import ABC, abstractmethod
from typing import Dict


class PowerFlowSolver(ABC):

    def __init__(self, config: dict = None, **kwargs):
        return

    @abstractmethod
    def calculate_power_flow(self, p_controllable_consumed: Dict[str, any] = None,
                             q_controllable_consumed: Dict[str, any] = None, **kwargs) -> any:
        raise NotImplementedError

    @abstractmethod
    def get_bus_voltages(self) -> Dict[str, any]:
        raise NotImplementedError

    @abstractmethod
    def get_bus_voltage_by_name(self, name: str) -> any:
        raise NotImplementedError
