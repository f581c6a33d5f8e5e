from .base import BaseInputArgs, BaseObjective
from .matching import (
    MatchingInputArgs,
    MatchingSolverDualObjectiveFunction,
    MatchingSolverDualObjectiveFunctionDistributed,
)

from .matching_fairness import MatchingFairnessDualObjectiveFunction, build_fairness_constraints
from .miplib import MIPLIB2017ObjectiveFunction, MIPLIBInputArgs

__all__ = ["BaseInputArgs", "BaseObjective", "MatchingInputArgs", "MatchingSolverDualObjectiveFunction",
           "MatchingSolverDualObjectiveFunctionDistributed", "MIPLIBInputArgs", "MIPLIB2017ObjectiveFunction",
           "MatchingFairnessDualObjectiveFunction", "build_fairness_constraints"]
