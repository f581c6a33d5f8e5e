"""Module path of the reference (projections/cone.py); the operator lives in clamp.py."""
from dualip_b200.projections.clamp import coneProjection  # noqa: F401
