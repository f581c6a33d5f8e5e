"""Generic (non-block) LP objective: same class names, constructor signature and `calculate` keywords as the reference
(src/dualip/objectives/miplib.py), with the body of `calculate` replaced by one call into the C-ABI
(include/dualip_b200.h: dualip_lp_calc).

CUDA-only: tensors must live on a CUDA device.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from dualip_b200 import _native
from dualip_b200.objectives.base import BaseInputArgs, BaseObjective, ObjectiveResult
from dualip_b200.projections.base import ProjectionEntry

_IDX = {name: i for i, name in enumerate(_native.SCALAR_FIELDS)}


@dataclass
class MIPLIBInputArgs(BaseInputArgs):
    """Input arguments of the generic-LP objective (reference miplib.py:11-25): A (m x n, dense or torch.sparse),
    c (n,), b_vec (m,), a projection map over VARIABLES (box / cone entries) and an optional equality mask over rows."""

    A: torch.Tensor
    c: torch.Tensor
    projection_map: dict[str, ProjectionEntry]
    b_vec: torch.Tensor
    equality_mask: Optional[torch.Tensor] = None


def _bounds_from_projection_map(projection_map, n: int, device):
    """Per-variable clamp bounds equivalent to applying the map's entries in order (miplib.py:80-90): box(lower, upper)
    and cone(lower) / cone(upper) / cone() are element-wise clamps (box.py:16, cone.py:21-28)."""
    lo = torch.full((n,), float("-inf"), dtype=torch.float32, device=device)
    hi = torch.full((n,), float("inf"), dtype=torch.float32, device=device)
    for key, item in projection_map.items():
        idx = torch.as_tensor(list(item.indices) if not isinstance(item.indices, torch.Tensor) else item.indices,
                              dtype=torch.long, device=device)
        if idx.numel() == 0:
            continue
        params = dict(item.proj_params or {})
        if item.proj_type == "box":
            lower, upper = params.get("lower", 0.0), params.get("upper", 1.0)
        elif item.proj_type == "cone":
            lower, upper = params.get("lower"), params.get("upper")
            if lower is not None and upper is not None:
                raise ValueError("coneProjection accepts at most one of lower/upper")  # cone.py:16-17
        else:
            raise ValueError(f"projection '{item.proj_type}' (entry {key!r}) is not an element-wise bound; the generic-LP "
                             "objective supports box and cone entries")
        # entries are applied one after the other (miplib.py:80-90); a clamp after a clamp is the clamp whose bounds are
        # the earlier bounds pushed through the later one: min(max(., l2), u2)
        l2 = float("-inf") if lower is None else float(lower)
        u2 = float("inf") if upper is None else float(upper)
        lo[idx] = lo[idx].clamp_min(l2).clamp_max(u2)
        hi[idx] = hi[idx].clamp_min(l2).clamp_max(u2)
    return lo, hi


class MIPLIB2017ObjectiveFunction(BaseObjective):
    """Dual gradient, objective and regularisation penalty of a generic LP (reference miplib.py:28-109)."""

    def __init__(self, miplib_input_args: MIPLIBInputArgs, use_jacobi_precondition: bool = False):
        args = miplib_input_args
        if not (isinstance(args.c, torch.Tensor) and args.c.is_cuda):
            raise ValueError("dualip_b200 objectives need CUDA tensors (there is no CPU fallback)")
        self.device = args.c.device
        A = args.A.to(self.device)
        self.A = A
        self.m, self.n = int(A.shape[0]), int(A.shape[1])
        # sparse inputs stay sparse (the reference keeps CSR + CSC copies, miplib.py:41-42); only strided inputs are dense
        if A.layout == torch.strided:
            src = A.to(torch.float32)
        elif A.layout == torch.sparse_coo:
            src = A.to(torch.float32).coalesce()
        else:
            src = A.to(torch.float32)
        csr = src if src.layout == torch.sparse_csr else src.to_sparse_csr()
        csc = src if src.layout == torch.sparse_csc else src.to_sparse_csc()
        self._csr = (csr.crow_indices().to(torch.int32).contiguous(), csr.col_indices().to(torch.int32).contiguous(),
                     csr.values().to(torch.float32).contiguous())
        self._csc = (csc.ccol_indices().to(torch.int32).contiguous(), csc.row_indices().to(torch.int32).contiguous(),
                     csc.values().to(torch.float32).contiguous())
        self.nnz = int(self._csr[2].numel())
        self.primal_size = self.n
        self.c = args.c.to(device=self.device, dtype=torch.float32).contiguous()
        self.b_vec = args.b_vec.to(device=self.device, dtype=torch.float32).contiguous()
        self.projection_map = args.projection_map
        self.equality_mask = args.equality_mask.to(self.device) if args.equality_mask is not None else None
        self._lo, self._hi = _bounds_from_projection_map(self.projection_map, self.n, self.device)
        self.lower, self.upper = self._construct_variable_lower_upper_bound()
        self.use_jacobi_precondition = use_jacobi_precondition
        self.gamma = None
        self.is_distributed = False
        if use_jacobi_precondition:
            if args.A.layout != torch.strided:
                raise NotImplementedError("Jacobi preconditioning is not implemented for sparse matrices")  # miplib.py:50-52
            row_norms = torch.norm(A.to(torch.float32), dim=1, keepdim=True)
            self.row_norms = torch.where(row_norms == 0, torch.ones_like(row_norms), row_norms).squeeze()
            self._row_scale = (1.0 / self.row_norms).to(torch.float32).contiguous()
        else:
            self.row_norms = None
            self._row_scale = None
        self._scratch = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._desc = _native.LpDesc(
            m=self.m, n=self.n, nnz=self.nnz,
            csr_rowptr_dev=self._csr[0].data_ptr(), csr_col_dev=self._csr[1].data_ptr(), csr_val_dev=self._csr[2].data_ptr(),
            csc_colptr_dev=self._csc[0].data_ptr(), csc_row_dev=self._csc[1].data_ptr(), csc_val_dev=self._csc[2].data_ptr(),
            c_dev=self.c.data_ptr(), b_dev=self.b_vec.data_ptr(), lo_dev=self._lo.data_ptr(), hi_dev=self._hi.data_ptr(),
            row_scale_dev=self._row_scale.data_ptr() if self._row_scale is not None else None,
            device=self.device.index if self.device.index is not None else torch.cuda.current_device())
        self._x = torch.empty(self.n, dtype=torch.float32, device=self.device)

    # -- raw launch (device pointers; used by the Maximizer's fused loop) ----------------------------------
    def launch_calc(self, lam_ptr: int, gamma: float, grad_ptr: int, scal_ptr: int, x_ptr: Optional[int] = None,
                    diag_ptr: Optional[int] = None) -> None:
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = _native.lib().dualip_lp_calc(ctypes.byref(self._desc), lam_ptr, float(gamma), x_ptr or self._x.data_ptr(),
                                          grad_ptr, scal_ptr, self._scratch.data_ptr(), stream)
        _native.check(rc, "dualip_lp_calc")

    def calculate(self, dual_val: torch.Tensor, gamma: float = None, save_primal: bool = False, **kwargs) -> ObjectiveResult:
        """Dual gradient, objective and reg penalty at `dual_val` (reference miplib.py:60-109)."""
        if gamma is None:
            gamma = self.gamma
        if gamma is None:
            raise TypeError("calculate() needs gamma")
        if not dual_val.is_cuda:
            raise ValueError("dual_val must be a CUDA tensor (there is no CPU fallback)")
        lam = dual_val.to(device=self.device, dtype=torch.float32).contiguous()
        grad = torch.empty(self.m, dtype=torch.float32, device=self.device)
        scal = torch.zeros(len(_native.SCALAR_FIELDS), dtype=torch.float64, device=self.device)
        x = torch.empty(self.n, dtype=torch.float32, device=self.device) if save_primal else self._x
        with torch.cuda.device(self.device):
            self.launch_calc(lam.data_ptr(), gamma, grad.data_ptr(), scal.data_ptr(), x.data_ptr())
        s32 = scal.to(torch.float32)
        result = ObjectiveResult(dual_gradient=grad, dual_objective=s32[_IDX["dual_objective"]],
                                 reg_penalty=s32[_IDX["reg_penalty"]])
        result.scalars64 = scal
        if save_primal:
            result.primal_var = x
            result.primal_objective = s32[_IDX["primal_objective"]]
        return result

    def invert_jacobi_precondition(self, dual_val: torch.Tensor) -> torch.Tensor:
        """Dual of the ORIGINAL rows from the dual of the row-scaled problem (the call at reference run_solver.py:136-144
        targets a method the reference never defines; this is lambda / ||A_r||, what its calculate applies at :73-74)."""
        return dual_val if self.row_norms is None else dual_val / self.row_norms.to(dual_val.device)

    # -- host-side diagnostics, plain torch on the objective's device (not on the iteration path) ----------
    def _construct_variable_lower_upper_bound(self):
        """NaN where a bound is absent (reference miplib.py:111-121; accepts both the operators' `lower`/`upper` keys
        and the `l`/`u` keys that method looks for)."""
        lower = torch.full_like(self.c, float("nan"))
        upper = torch.full_like(self.c, float("nan"))
        for _, item in self.projection_map.items():
            idx = torch.as_tensor(list(item.indices) if not isinstance(item.indices, torch.Tensor) else item.indices,
                                  dtype=torch.long, device=self.c.device)
            params = item.proj_params or {}
            for key in ("l", "lower"):
                if params.get(key) is not None:
                    lower[idx] = params[key]
            for key in ("u", "upper"):
                if params.get(key) is not None:
                    upper[idx] = params[key]
        return lower, upper

    @staticmethod
    def _clamp_x_bound_duals(x_bound_duals, l_mask_exists, u_mask_exists):
        """Projection of the bound duals onto the set Lambda of PDLP (reference miplib.py:123-154)."""
        result = x_bound_duals.clone()
        only_l = l_mask_exists & ~u_mask_exists
        only_u = ~l_mask_exists & u_mask_exists
        result[only_l] = torch.clamp(result[only_l], min=0)
        result[only_u] = torch.clamp(result[only_u], max=0)
        result[~l_mask_exists & ~u_mask_exists] = 0
        return result

    def calculate_convergence_bound(self, dual_val: torch.Tensor, x: torch.Tensor = None, optimal_primal_obj=None,
                                    tol: float = 1e-4):
        """PDLP stopping test without regularisation (reference miplib.py:156-230): relative duality gap, primal and
        dual feasibility.  Returns (gap_upperbound, gap_lower_bound, primal_feas, dual_feas, converged)."""
        dual_val = dual_val.to(self.device)
        if self.row_norms is not None:
            dual_val = 1 / self.row_norms * dual_val
        if self.A.layout == torch.strided:
            A32 = self.A.to(torch.float32)
            a_mv, at_mv = A32.mv, A32.t().mv
        else:  # sparse products on the CSR / CSC copies built at construction; never densified
            csr = torch.sparse_csr_tensor(self._csr[0], self._csr[1], self._csr[2], size=(self.m, self.n))
            csct = torch.sparse_csr_tensor(self._csc[0], self._csc[1], self._csc[2], size=(self.n, self.m))  # A^T as CSR
            a_mv = lambda v: (csr @ v.unsqueeze(1)).squeeze(1)  # noqa: E731
            at_mv = lambda v: (csct @ v.unsqueeze(1)).squeeze(1)  # noqa: E731
        r = self.c + at_mv(dual_val)
        if x is None:
            x = torch.where(r >= 0, self.lower, self.upper)
            if torch.isnan(x).any():
                raise ValueError("Unbounded x.")
        lambda_neg, lambda_pos = torch.clamp(r, max=0.0), torch.clamp(r, min=0.0)
        u_exists, l_exists = ~torch.isnan(self.upper), ~torch.isnan(self.lower)
        d = (-torch.dot(self.b_vec, dual_val) + torch.dot(lambda_neg[u_exists], self.upper[u_exists])
             + torch.dot(lambda_pos[l_exists], self.lower[l_exists]))
        p = torch.dot(self.c, x)
        gap_upperbound = torch.abs(p - d) / (1.0 + torch.abs(p) + torch.abs(d))
        if optimal_primal_obj is not None:
            gap_lower_bound = torch.abs(p - optimal_primal_obj) / (1.0 + torch.abs(p) + abs(optimal_primal_obj))
        else:
            gap_lower_bound = torch.tensor(float("nan"))
        resid = a_mv(x) - self.b_vec
        violation = torch.relu(resid) if self.equality_mask is None else torch.where(self.equality_mask, resid.abs(), torch.relu(resid))
        primal_feas = torch.linalg.vector_norm(violation) / (1.0 + torch.linalg.vector_norm(self.b_vec))
        x_bound_duals = self._clamp_x_bound_duals(-r, l_exists, u_exists)
        dual_feas = torch.linalg.vector_norm(r + x_bound_duals) / (1.0 + torch.linalg.vector_norm(self.c))
        converged = bool((gap_upperbound <= tol) and (primal_feas <= tol) and (dual_feas <= tol))
        return gap_upperbound, gap_lower_bound, primal_feas, dual_feas, converged
