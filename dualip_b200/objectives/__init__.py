from .base import BaseInputArgs, BaseObjective
from .matching import (
    MatchingInputArgs,
    MatchingSolverDualObjectiveFunction,
    MatchingSolverDualObjectiveFunctionDistributed,
)

__all__ = ["BaseInputArgs", "BaseObjective", "MatchingInputArgs", "MatchingSolverDualObjectiveFunction",
           "MatchingSolverDualObjectiveFunctionDistributed"]
