import math

from dualip_b200 import _native
from dualip_b200.projections.base import ProjectionOperator, register


@register("cone")
class coneProjection(ProjectionOperator):
    """[lower, +inf) or (-inf, upper] per coordinate; identity when neither is given
    (reference projections/cone.py:6-28)."""

    def __init__(self, lower: float | None = None, upper: float | None = None):
        if lower is not None and upper is not None:
            raise ValueError("Only one of 'lower' or 'upper' should be specified, not both.")
        self.lower, self.upper = lower, upper

    def native_class(self) -> _native.ProjClass:
        lo = -math.inf if self.lower is None else float(self.lower)
        hi = math.inf if self.upper is None else float(self.upper)
        return _native.ProjClass(_native.PROJ_CLAMP, lo, hi, 1.0, 1.0, 0)
