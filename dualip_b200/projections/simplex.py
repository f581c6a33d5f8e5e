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
            return None  # not in the fused kernel: the objective routes these columns through padded blocks
        return _simplex_class(self._kind, self.z)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.proj_method != "bisection_search":
            return super().__call__(x)
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("dualip_b200 projections run on CUDA tensors only (no CPU fallback)")
        if x.ndim == 1:
            x = x.unsqueeze(1)
        return _bisection_block(x, float(self.z), self._kind == _native.PROJ_SIMPLEX)


def _bisection_block(x: torch.Tensor, z: float, inequality: bool, tol: float = 1e-6, max_iter: int = 50) -> torch.Tensor:
    """The reference's bisection projection (projections/simplex.py:6-123) on a zero-padded [L x K] device block, kept for
    `method="bisection_search"`; tensor operations on the device, not part of the fused kernel (no benchmark
    configuration uses it).  Reference behaviour that is reproduced: no pre-clamp; feasibility needs every entry >= -tol
    (:40); top-2 shortcut on x/z (:52-81); the shift uses max(x/z) and the root is searched for a sum of 1 (:86-104);
    every active column halves the same interval [-1, 0], so the step count is the same for all columns (:95-118)."""
    if not z > 0:
        raise AssertionError("Simplex radius z must be positive.")
    L, K = x.shape
    w = torch.empty_like(x)
    todo = torch.ones(K, dtype=torch.bool, device=x.device)
    if inequality:
        feasible = (x.sum(dim=0) <= z + tol) & (x >= -tol).all(dim=0)
        w[:, feasible] = x[:, feasible]
        todo &= ~feasible
    if L > 1:
        cand = todo.nonzero(as_tuple=True)[0]
        if cand.numel():
            vals, pos = torch.topk(x[:, cand] / z, 2, dim=0)
            short = (vals[0] - vals[1]) > 1.0
            cols = cand[short]
            if cols.numel():
                sol = torch.zeros(L, cols.numel(), dtype=x.dtype, device=x.device)
                sol[pos[0, short], torch.arange(cols.numel(), device=x.device)] = z
                w[:, cols] = sol
                todo[cols] = False
    rest = todo.nonzero(as_tuple=True)[0]
    if rest.numel() == 0:
        return w
    sub = x[:, rest]
    shifted = sub - (sub / z).max(dim=0).values.unsqueeze(0)
    lo = torch.full((rest.numel(),), -1.0, dtype=x.dtype, device=x.device)
    hi = torch.zeros_like(lo)
    active = torch.ones_like(lo, dtype=torch.bool)
    prev = None
    for _ in range(max_iter):
        mid = (lo + hi) / 2.0
        if prev is not None and float((mid - prev).abs().max()) < tol:
            break
        high = torch.clamp(shifted - mid.unsqueeze(0), min=0.0).sum(dim=0) > 1.0
        lo = torch.where(high & active, mid, lo)
        hi = torch.where(~high & active, mid, hi)
        active = active & ~((hi - lo) < tol)
        prev = mid
        if not bool(active.any()):
            break
    nu = (lo + hi) / 2.0
    w[:, rest] = torch.clamp(shifted - nu.unsqueeze(0), min=0.0) * z
    return w


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
