"""Abstract bases of the ObjectiveFunction protocol (reference src/dualip/objectives/base.py:8-26)."""
from abc import ABC, abstractmethod
from dataclasses import dataclass

from dualip_b200.types import ObjectiveResult


@dataclass
class BaseInputArgs(ABC):
    def __post_init__(self):
        pass


class BaseObjective(ABC):
    @abstractmethod
    def calculate(self) -> ObjectiveResult:
        pass


__all__ = ["BaseInputArgs", "BaseObjective", "ObjectiveResult"]
