import numpy as np

from dualip_b200 import _native
from dualip_b200.projections.base import ProjectionOperator, register


def _simplex_class(kind: int, z: float) -> _native.ProjClass:
    if not z > 0:
        raise AssertionError("Simplex radius z must be positive.")  # reference simplex.py:145
    # the reference compares the column sum with the Python float z + tol, rounded to float32 (simplex.py:154)
    z_thr = float(np.float32(float(z) + 1e-6))
    return _native.ProjClass(kind, 0.0, 0.0, float(np.float32(z)), z_thr, 0)


class _SimplexBase(ProjectionOperator):
    _kind = _native.PROJ_SIMPLEX

    def __init__(self, z: float = 1.0, method: str = "duchi"):
        self.z = z
        self.proj_method = method
        if self.proj_method not in ("duchi", "bisection_search"):
            raise ValueError(f"Unsupported projection method: {self.proj_method}")
        if self.proj_method == "bisection_search":
            # reference simplex.py:6-123; not on the path of any benchmark configuration
            raise NotImplementedError("method='bisection_search' is not implemented by dualip_b200; use 'duchi'")

    def native_class(self) -> _native.ProjClass:
        return _simplex_class(self._kind, self.z)


@register("simplex")
class SimplexIneq(_SimplexBase):
    """{x >= 0, sum x <= z}: batched Duchi with pre-clamp (reference projections/simplex.py:238-255)."""

    _kind = _native.PROJ_SIMPLEX


@register("simplex_eq")
class SimplexEq(_SimplexBase):
    """{x >= 0, sum x = z} (reference projections/simplex.py:258-274).  Inside the fused matching kernel a column is
    projected at its true length; the reference pads it to its bucket's length, which changes the result when the
    column sum is below z (reference quirk #4)."""

    _kind = _native.PROJ_SIMPLEX_EQ
