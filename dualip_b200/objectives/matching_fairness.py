"""Matching objective with the two fairness rows of the reference's demo (docs/demo/matching_complex.rst).

The reference ships no class for it: the demo shows a subclass of MatchingSolverDualObjectiveFunction whose `calculate`
is composed from the public CSC operators (rst:82-168) and whose extra constraint matrix comes from
`_build_fairness_constraints` (rst:46-64): the columns are split into two groups by `group_ratio`; group 1 keeps A's
values times 1/|group 1|, group 2 gets A's values times -1/|group 2|.  With dual variables lambda in R^{m+2} and
b' = (b, delta, delta) the two extra rows enforce | mean load of group 1 - mean load of group 2 | <= delta.

Here that recipe is ONE kernel over the caller's CSC arrays plus the shared m+2-length tail (dualip_fair_calc,
csrc/ops.cu): v = a*scaled_r + scaled_m*f - scaled_{m+1}*f + c_rescaled in the demo's operation order, projection per
column with the reference's padded-block semantics, row sums, the two dense rows, c.x and ||x||^2.  It plugs into the same
Maximizer (the generic device loop: dualip_agd_step after every evaluation).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from dualip_b200 import _native
from dualip_b200.objectives.base import BaseObjective, ObjectiveResult
from dualip_b200.objectives.matching import MatchingInputArgs, _build_class_table, _build_pad_table
from dualip_b200.utils.sparse_utils import hstack_csc, split_csc_by_cols

_IDX = {name: i for i, name in enumerate(_native.SCALAR_FIELDS)}


def build_fairness_constraints(A: torch.Tensor, group_ratio: float) -> torch.Tensor:
    """A_fairness of the demo (rst:46-64): same pattern as A; first int(n * group_ratio) columns scaled by 1/|group 1|, the
    remaining ones by -1/|group 2|."""
    num_cols = A.size(1)
    g1 = max(0, min(int(num_cols * group_ratio), num_cols))
    g2 = num_cols - g1
    if g1 == 0 or g2 == 0:
        raise ValueError("group_ratio must leave both groups non-empty")
    blocks = split_csc_by_cols(A, [g1, g2])
    scaled = []
    for blk, w in zip(blocks, (1 / g1, -1 / g2)):
        scaled.append(torch.sparse_csc_tensor(blk.ccol_indices(), blk.row_indices(), blk.values() * w, size=blk.size()))
    return hstack_csc(scaled)


class MatchingFairnessDualObjectiveFunction(BaseObjective):
    """Matching LP with two dense fairness rows.  `matching_input_args.A` has m rows; `b_vec` and the dual have m + 2
    entries (b' = (b, delta, delta)).  Either `A_fairness` (CSC, A's pattern) or `group_ratio` must be given."""

    def __init__(self, matching_input_args: MatchingInputArgs, gamma: float, batching: bool = True,
                 A_fairness: Optional[torch.Tensor] = None, group_ratio: Optional[float] = None):
        A, c = matching_input_args.A, matching_input_args.c
        if A.layout != torch.sparse_csc or c.layout != torch.sparse_csc:
            raise ValueError("Both A and c must be CSC-format sparse tensors")
        if not A.is_cuda:
            raise RuntimeError("dualip_b200 objectives need CUDA tensors (no CPU fallback); move the inputs to a GPU")
        if A.values().dtype != torch.float32 or c.values().dtype != torch.float32:
            raise TypeError("dualip_b200 is float32-only")
        if A.shape != c.shape or A.values().shape != c.values().shape:
            raise ValueError("A and c must share the same sparsity pattern")
        if A_fairness is None:
            if group_ratio is None:
                raise ValueError("give A_fairness or group_ratio")
            A_fairness = build_fairness_constraints(A, group_ratio)
        if A_fairness.layout != torch.sparse_csc or A_fairness.values().shape != A.values().shape:
            raise ValueError("A_fairness must be a CSC tensor with A's sparsity pattern")
        self.A, self.c, self.A_fairness = A, c, A_fairness
        self.gamma, self.batching = gamma, batching
        self.device = A.device
        self.m, self.n = int(A.shape[0]), int(A.shape[1])
        self.b_vec = matching_input_args.b_vec
        if self.b_vec is None or self.b_vec.numel() != self.m + 2:
            raise ValueError(f"b_vec must have m + 2 = {self.m + 2} entries (b, delta, delta)")
        self.b_vec = self.b_vec.to(device=self.device, dtype=torch.float32).contiguous()
        self.equality_mask = matching_input_args.equality_mask
        self.projection_map = matching_input_args.projection_map
        self.is_distributed = False
        ccol, row = A.ccol_indices(), A.row_indices()
        if ccol.dtype != row.dtype or ccol.dtype not in (torch.int32, torch.int64):
            raise TypeError("ccol_indices and row_indices must both be int32 or both int64")
        self._a, self._c = A.values().contiguous(), c.values().contiguous()
        self._f = A_fairness.values().to(torch.float32).contiguous()
        self.nnz = int(self._a.numel())
        self.primal_size = self.nnz
        self._classes, self._n_classes, self._col_class = _build_class_table(ccol, self.n, self.projection_map, batching, self.m)
        self._pad = _build_pad_table(ccol, self.n, self.projection_map, batching, self.m, self._n_classes)
        self._desc = _native.CscDesc(
            n_cols=self.n, nnz=self.nnz, n_rows=self.m, index_bits=32 if ccol.dtype == torch.int32 else 64,
            ccol_dev=ccol.data_ptr(), row_dev=row.data_ptr() if self.nnz else 0, a_dev=self._a.data_ptr() if self.nnz else 0,
            c_dev=self._c.data_ptr() if self.nnz else 0,
            col_class_dev=self._col_class.data_ptr() if self._col_class is not None else None,
            classes=ctypes.cast(self._classes, ctypes.POINTER(_native.ProjClass)), n_classes=self._n_classes,
            device=self.device.index if self.device.index is not None else torch.cuda.current_device(),
            pad_len=ctypes.cast(self._pad, ctypes.POINTER(ctypes.c_int32)) if self._pad is not None else None)
        nbytes = int(_native.lib().dualip_fair_work_bytes(self.m, self._n_classes))
        self._work = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=self.device)  # 8-byte aligned, zeroed once
        self._xbuf = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=self.device)

    # raw launch used by the Maximizer's device loop
    def launch_calc(self, lam_ptr: int, gamma: float, grad_ptr: int, scal_ptr: int, x_ptr: Optional[int] = None,
                    diag_ptr: Optional[int] = None) -> None:
        rc = _native.lib().dualip_fair_calc(ctypes.byref(self._desc), self._f.data_ptr(), lam_ptr, self.b_vec.data_ptr(),
                                            float(gamma), grad_ptr, scal_ptr, x_ptr if x_ptr else self._xbuf.data_ptr(),
                                            self._work.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _native.check(rc, "dualip_fair_calc")

    def calculate(self, dual_val: torch.Tensor, gamma: float = None, save_primal: bool = False, **kwargs) -> ObjectiveResult:
        """Same contract as the demo's `calculate` (rst:86-167); dual_val has m + 2 entries."""
        if gamma is not None:
            self.gamma = gamma
        if dual_val.device != self.device or dual_val.dtype != torch.float32 or dual_val.numel() != self.m + 2:
            raise ValueError(f"dual_val must be a float32 tensor with {self.m + 2} entries on {self.device}")
        lam = dual_val.contiguous()
        with torch.cuda.device(self.device):
            grad = torch.empty(self.m + 2, dtype=torch.float32, device=self.device)
            scal = torch.empty(len(_native.SCALAR_FIELDS), dtype=torch.float64, device=self.device)
            x = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=self.device) if save_primal else None
            self.launch_calc(lam.data_ptr(), self.gamma, grad.data_ptr(), scal.data_ptr(), x.data_ptr() if x is not None else None)
            s32 = scal.to(torch.float32)
        res = ObjectiveResult(
            dual_gradient=grad,
            dual_objective=s32[_IDX["dual_objective"]],
            reg_penalty=s32[_IDX["reg_penalty"]],
            dual_val_times_grad=s32[_IDX["dual_val_times_grad"]],
            max_pos_slack=s32[_IDX["max_pos_slack"]],
            sum_pos_slack=s32[_IDX["sum_pos_slack"]],
        )
        if save_primal:
            res.primal_var = x[: self.nnz]
            res.primal_objective = s32[_IDX["primal_objective"]].clone()
        res.scalars64 = scal
        return res
