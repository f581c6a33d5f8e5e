"""Setup-time CSC helpers used by sharding (reference src/dualip/utils/sparse_utils.py:246-290).  The per-iteration
operators of that file (left_multiply_sparse, elementwise_csc, apply_F_to_columns, row_sums_csc) have no counterpart
here: they are fused into matching_pass_kernel (csrc/calc.cu)."""
from typing import List

import torch


def split_csc_by_cols(M: torch.Tensor, split_sizes: List[int]) -> List[torch.Tensor]:
    """Contiguous column blocks of a CSC matrix with rebased column pointers.  One host read of the W+1 boundary
    pointers instead of the reference's two `.item()` syncs per block."""
    if M.layout != torch.sparse_csc:
        raise ValueError("M must be CSC-format sparse")
    m, n = M.size()
    if sum(split_sizes) != n:
        raise ValueError(f"split_sizes must sum to {n}")
    ccol, row, vals = M.ccol_indices(), M.row_indices(), M.values()
    bounds = [0]
    for w in split_sizes:
        bounds.append(bounds[-1] + w)
    nnz_bounds = ccol[torch.tensor(bounds, device=ccol.device)].tolist()
    blocks = []
    for k, width in enumerate(split_sizes):
        c0, c1 = bounds[k], bounds[k + 1]
        e0, e1 = nnz_bounds[k], nnz_bounds[k + 1]
        blocks.append(torch.sparse_csc_tensor((ccol[c0 : c1 + 1] - e0).clone(), row[e0:e1].clone(), vals[e0:e1].clone(),
                                              size=(m, width)))
    return blocks
