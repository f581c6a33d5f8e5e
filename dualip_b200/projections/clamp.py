"""Element-wise projections (reference projections/box.py:6-16 and projections/cone.py:6-28): both are one clamp
x = min(max(v, lo), hi) per entry, with an infinite bound for an open side, and share one row format of the C-ABI class
table (DUALIP_PROJ_CLAMP).  The fused kernel applies them in registers; `__call__` on a padded block goes through
dualip_project_block (ProjectionOperator.__call__)."""
import math

from dualip_b200 import _native
from dualip_b200.projections.base import ProjectionOperator, register


class _Clamp(ProjectionOperator):
    lower = None
    upper = None

    def bounds(self) -> tuple:
        """(lo, hi) with -inf / +inf for a side that is not constrained."""
        return (-math.inf if self.lower is None else float(self.lower),
                math.inf if self.upper is None else float(self.upper))

    def native_class(self) -> _native.ProjClass:
        lo, hi = self.bounds()
        return _native.ProjClass(_native.PROJ_CLAMP, lo, hi, 1.0, 1.0, 0)


@register("box")
class BoxProjection(_Clamp):
    """[lower, upper] per coordinate; defaults [0, 1] (reference box.py:11-13)."""

    def __init__(self, lower: float = 0.0, upper: float = 1.0):
        self.lower, self.upper = lower, upper


@register("cone")
class coneProjection(_Clamp):
    """[lower, +inf) or (-inf, upper] per coordinate, identity when neither is given; naming both is an error, as in the
    reference (cone.py:13-19).  The reference spells the class with a lower-case c."""

    def __init__(self, lower: float | None = None, upper: float | None = None):
        if lower is not None and upper is not None:
            raise ValueError("Only one of 'lower' or 'upper' should be specified, not both.")
        self.lower, self.upper = lower, upper
