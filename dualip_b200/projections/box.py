"""Module path of the reference (projections/box.py); the operator lives in clamp.py."""
from dualip_b200.projections.clamp import BoxProjection  # noqa: F401
