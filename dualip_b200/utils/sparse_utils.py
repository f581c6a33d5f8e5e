"""CSC operators with the reference's names and signatures (src/dualip/utils/sparse_utils.py).

The stock matching objective does not call the per-iteration operators -- its chain left_multiply_sparse ->
elementwise_csc -> apply_F_to_columns -> row_sums_csc is one fused kernel (csrc/calc.cu).  They are kept as device
operators because the reference documents them as the way to extend the objective
(docs/demo/matching_complex.rst:82-168: a subclass overrides `calculate` and composes them); each one is a kernel of
csrc/ops.cu behind the C-ABI (dualip_csc_*), applied to the value arrays of CUDA `torch.sparse_csc` tensors.  CPU tensors
are rejected: there is no CPU fallback.  The setup-time helpers (split / stack) are plain index arithmetic on the device.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

from dualip_b200 import _native


def _need_csc(*tensors) -> None:
    for t in tensors:
        if t.layout != torch.sparse_csc:
            raise ValueError("Expected a CSC-format sparse tensor")


def _need_cuda_f32(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"dualip_b200 operators need CUDA tensors (no CPU fallback): {what} is on {t.device}")
    if t.values().dtype != torch.float32:
        raise TypeError(f"dualip_b200 is float32-only: {what} holds {t.values().dtype}")


def _index_bits(idx: torch.Tensor) -> int:
    if idx.dtype == torch.int32:
        return 32
    if idx.dtype == torch.int64:
        return 64
    raise TypeError("CSC indices must be int32 or int64")


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dot_product_csc(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """sum_ij A_ij B_ij of two CSC tensors with one pattern (reference sparse_utils.py:7-23)."""
    assert A.layout == torch.sparse_csc and B.layout == torch.sparse_csc, "Inputs must both be CSC sparse tensors"
    assert A.shape == B.shape, f"Expected shapes (m, n) and (m, n), got {A.shape} and {B.shape}"
    return torch.dot(A.values(), B.values())


def elementwise_csc(A: torch.Tensor, B: torch.Tensor, op, output_tensor: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`op` applied to the values of two CSC tensors with identical patterns (reference sparse_utils.py:26-51).  With
    `output_tensor` the pattern check is skipped and the result is written into its values (and returned), like the
    reference.  `op` is any callable on two value tensors (operator.add / mul / sub ...): it runs on the device."""
    if A.layout != torch.sparse_csc or B.layout != torch.sparse_csc:
        raise ValueError("Both A and B must be CSC-format sparse tensors")
    if output_tensor is None and not (
        torch.equal(A.ccol_indices(), B.ccol_indices()) and torch.equal(A.row_indices(), B.row_indices())
    ):
        raise ValueError("A and B must share the same sparsity pattern")
    new_vals = op(A.values(), B.values())
    if output_tensor is None:
        return torch.sparse_csc_tensor(A.ccol_indices(), A.row_indices(), new_vals, size=A.size())
    return output_tensor.values().copy_(new_vals)


def left_multiply_sparse(v: torch.Tensor, M: torch.Tensor, output_tensor: Optional[torch.Tensor] = None) -> torch.Tensor:
    """diag(v) @ M on the stored values (reference sparse_utils.py:54-85): one gather-multiply kernel
    (dualip_csc_left_multiply), written straight into `output_tensor`'s values when given."""
    if M.layout != torch.sparse_csc:
        raise ValueError("Expected M to be a CSC-format sparse tensor")
    _need_cuda_f32(M, "M")
    row, vals = M.row_indices(), M.values().contiguous()
    vv = v.to(device=vals.device, dtype=torch.float32).contiguous()
    if vv.numel() < M.size(0):
        raise IndexError(f"v has {vv.numel()} entries, M has {M.size(0)} rows")
    dst = output_tensor.values() if output_tensor is not None else torch.empty_like(vals)
    out = dst if dst.is_contiguous() else torch.empty_like(vals)
    with torch.cuda.device(vals.device):
        _native.check(_native.lib().dualip_csc_left_multiply(
            vals.data_ptr(), row.data_ptr(), _index_bits(row), vals.numel(), vv.data_ptr(), out.data_ptr(), vals.device.index,
            _stream(vals.device)), "dualip_csc_left_multiply")
    if out is not dst:
        dst.copy_(out)
    if output_tensor is None:
        return torch.sparse_csc_tensor(M.ccol_indices(), row, dst, size=M.size())
    return dst


def right_multiply_sparse(M: torch.Tensor, v: torch.Tensor, output_tensor: Optional[torch.Tensor] = None) -> torch.Tensor:
    """M @ diag(v) (reference sparse_utils.py:88-130).  The column of every stored value comes from one repeat_interleave
    over the column lengths instead of the reference's Python loop with two host reads per column."""
    if M.layout != torch.sparse_csc:
        raise ValueError("Expected M to be a CSC-format sparse tensor")
    ccol, row, vals = M.ccol_indices(), M.row_indices(), M.values()
    col_of = torch.repeat_interleave(torch.arange(M.size(1), device=vals.device), (ccol[1:] - ccol[:-1]).to(torch.int64))
    new_vals = vals * v.to(vals.device)[col_of]
    if output_tensor is None:
        return torch.sparse_csc_tensor(ccol, row, new_vals, size=M.size())
    return output_tensor.values().copy_(new_vals)


def apply_F_to_columns(M: torch.Tensor, F_batch: Callable[[torch.Tensor], torch.Tensor], buckets: List[torch.Tensor],
                       output_tensor: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F applied column-wise through zero-padded [L x K] blocks, one block per bucket of column indices (reference
    sparse_utils.py:133-220): the block is built and written back by two kernels (dualip_csc_gather_block /
    dualip_csc_scatter_block) instead of three length-nnz index vectors; `F_batch` is any callable on the block -- the
    package's own projection operators run their native kernel on it (dualip_project_block).

    Columns that no bucket names keep M's values (the reference leaves them as `torch.empty_like` allocated them, which
    corrupts maps with several entries: sparse_utils.py:177,220 -- not reproduced)."""
    assert M.layout == torch.sparse_csc, "M must be a CSC sparse tensor"
    _need_cuda_f32(M, "M")
    ccol, rowi, vals = M.ccol_indices(), M.row_indices(), M.values().contiguous()
    device = vals.device
    new_vals = vals.clone()
    lib, bits = _native.lib(), _index_bits(ccol)
    with torch.cuda.device(device):
        for cols in buckets:
            K = int(cols.numel())
            if K == 0:
                continue
            cols64 = cols.to(device=device, dtype=torch.int64).contiguous()
            lengths = ccol[cols64 + 1] - ccol[cols64]
            L = int(lengths.max().item())  # the one host read per bucket (the reference has two, sparse_utils.py:189,197)
            if L == 0:
                continue
            block = torch.empty((L, K), dtype=torch.float32, device=device)
            _native.check(lib.dualip_csc_gather_block(ccol.data_ptr(), bits, vals.data_ptr(), cols64.data_ptr(), K, L,
                                                      block.data_ptr(), device.index, _stream(device)), "dualip_csc_gather_block")
            proj = F_batch(block)
            if proj.shape != block.shape:
                raise ValueError(f"F_batch returned shape {tuple(proj.shape)}, expected {tuple(block.shape)}")
            proj = proj.to(torch.float32).contiguous()
            _native.check(lib.dualip_csc_scatter_block(ccol.data_ptr(), bits, proj.data_ptr(), cols64.data_ptr(), K, L,
                                                       new_vals.data_ptr(), device.index, _stream(device)),
                          "dualip_csc_scatter_block")
    if output_tensor is None:
        return torch.sparse_csc_tensor(ccol, rowi, new_vals, size=M.size())
    return output_tensor.values().copy_(new_vals)


def row_sums_csc(A: torch.Tensor) -> torch.Tensor:
    """Dense vector of row sums (reference sparse_utils.py:223-243): per-CTA shared-memory sums, then one add per touched
    row (dualip_csc_row_sums).  fp32 atomics like the reference's CUDA scatter_add_: the order of additions is not fixed."""
    _need_csc(A)
    _need_cuda_f32(A, "A")
    row, vals = A.row_indices(), A.values().contiguous()
    out = torch.empty(A.size(0), dtype=torch.float32, device=vals.device)
    with torch.cuda.device(vals.device):
        _native.check(_native.lib().dualip_csc_row_sums(vals.data_ptr(), row.data_ptr(), _index_bits(row), vals.numel(),
                                                        A.size(0), out.data_ptr(), vals.device.index, _stream(vals.device)),
                      "dualip_csc_row_sums")
    return out


def row_norms_csc(A: torch.Tensor) -> torch.Tensor:
    """L2 norm of every row (reference sparse_utils.py:429-450)."""
    _need_csc(A)
    sq = torch.sparse_csc_tensor(A.ccol_indices(), A.row_indices(), A.values().pow(2), size=A.size())
    return row_sums_csc(sq).pow(0.5)


def split_csc_by_cols(M: torch.Tensor, split_sizes: List[int]) -> List[torch.Tensor]:
    """Contiguous column blocks of a CSC matrix with rebased column pointers (reference sparse_utils.py:246-290).  One host
    read of the W+1 boundary pointers instead of the reference's two `.item()` syncs per block."""
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


def hstack_csc(tensors: List[torch.Tensor]) -> torch.Tensor:
    """Column-wise concatenation of CSC tensors (reference sparse_utils.py:293-349)."""
    n_rows, dtype, device = tensors[0].size(0), tensors[0].dtype, tensors[0].device
    for i, t in enumerate(tensors):
        if t.size(0) != n_rows:
            raise ValueError(f"tensor {i} has {t.size(0)} rows, expected {n_rows}")
        if t.dtype != dtype:
            raise TypeError("all tensors must share the same dtype")
        if t.device != device:
            raise TypeError("all tensors must be on the same device")
    chunks, nnz_prefix, total_cols = [], 0, 0
    for k, t in enumerate(tensors):
        ptr = t.ccol_indices()
        chunks.append(ptr if k == 0 else ptr[1:] + nnz_prefix)
        nnz_prefix += t.values().shape[0]
        total_cols += t.size(1)
    return torch.sparse_csc_tensor(torch.cat(chunks), torch.cat([t.row_indices() for t in tensors]),
                                   torch.cat([t.values() for t in tensors]), size=(n_rows, total_cols), dtype=dtype, device=device)


def vstack_csc(tensors: List[torch.Tensor]) -> torch.Tensor:
    """Row-wise stacking of CSC tensors (reference sparse_utils.py:352-426).  The reference walks the columns in Python with
    two host reads per column and tensor; here every stored entry gets the key (column, source tensor, position) and ONE
    stable sort by column interleaves the inputs."""
    if not tensors:
        raise ValueError("Cannot stack empty list of tensors")
    n_cols, dtype, device = tensors[0].size(1), tensors[0].dtype, tensors[0].device
    for i, t in enumerate(tensors):
        if t.layout != torch.sparse_csc:
            raise ValueError(f"tensor {i} must be CSC-format sparse")
        if t.size(1) != n_cols:
            raise ValueError(f"tensor {i} has {t.size(1)} columns, expected {n_cols}")
        if t.dtype != dtype:
            raise TypeError("all tensors must share the same dtype")
        if t.device != device:
            raise TypeError("all tensors must be on the same device")
    cols, rows, vals, counts, row_offset = [], [], [], torch.zeros(n_cols, dtype=torch.int64, device=device), 0
    for t in tensors:
        lengths = (t.ccol_indices()[1:] - t.ccol_indices()[:-1]).to(torch.int64)
        cols.append(torch.repeat_interleave(torch.arange(n_cols, device=device), lengths))
        rows.append(t.row_indices().to(torch.int64) + row_offset)
        vals.append(t.values())
        counts += lengths
        row_offset += t.size(0)
    col_all = torch.cat(cols)
    order = torch.sort(col_all, stable=True).indices  # inputs are concatenated in tensor order: stable keeps it per column
    new_ccol = torch.zeros(n_cols + 1, dtype=torch.int64, device=device)
    new_ccol[1:] = counts.cumsum(0)
    return torch.sparse_csc_tensor(new_ccol, torch.cat(rows)[order], torch.cat(vals)[order], size=(row_offset, n_cols),
                                   dtype=dtype, device=device)
