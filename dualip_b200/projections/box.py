from dualip_b200 import _native
from dualip_b200.projections.base import ProjectionOperator, register


@register("box")
class BoxProjection(ProjectionOperator):
    """Projection onto [lower, upper] per coordinate (reference projections/box.py:6-16)."""

    def __init__(self, lower: float = 0.0, upper: float = 1.0):
        self.lower, self.upper = lower, upper

    def native_class(self) -> _native.ProjClass:
        return _native.ProjClass(_native.PROJ_CLAMP, float(self.lower), float(self.upper), 1.0, 1.0, 0)
