import numpy as np
import torch

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

    def native_class(self):
        if self.proj_method == "bisection_search":
            # the fused kernel and dualip_project_block restate the reference's bisection (simplex.py:6-123) per column
            kind = _native.PROJ_SIMPLEX_BISECT if self._kind == _native.PROJ_SIMPLEX else _native.PROJ_SIMPLEX_EQ_BISECT
            return _simplex_class(kind, self.z)
        return _simplex_class(self._kind, self.z)


@register("simplex")
class SimplexIneq(_SimplexBase):
    """{x >= 0, sum x <= z}: batched Duchi with pre-clamp (reference projections/simplex.py:238-255)."""

    _kind = _native.PROJ_SIMPLEX


@register("simplex_eq")
class SimplexEq(_SimplexBase):
    """{x >= 0, sum x = z} (reference projections/simplex.py:258-274).  The result depends on the padded length of the
    block a column is projected in when the column sum is below z; the fused kernel gets that length
    per column class and length bucket (objectives/matching.py: _build_pad_table)."""

    _kind = _native.PROJ_SIMPLEX_EQ
