"""Jacobi (row-norm) preconditioning, in place, with the reference's function names
(src/dualip/preprocessing/precondition.py:8-60).  CUDA tensors only; two streaming kernels in csrc/setup.cu."""
from pathlib import Path

import torch

from dualip_b200 import _native


def jacobi_precondition(A: torch.Tensor, b: torch.Tensor, norms_save_path: str = None, sharded: bool = False):
    """A <- diag(1/||A_r||_2) A and b <- b/||A_r||_2 in place; returns the row norms.

    sharded=True: A is this rank's column shard; squared row norms are summed over the process group first, so
    every rank scales with the norms of the full matrix (b is replicated and scaled identically everywhere)."""
    if A.layout != torch.sparse_csc:
        raise ValueError("Expected M to be a CSC-format sparse tensor")
    vals, row = A.values(), A.row_indices()
    if not vals.is_cuda:
        raise RuntimeError("dualip_b200 preprocessing runs on CUDA tensors only (no CPU fallback)")
    if vals.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError("dualip_b200 is float32-only")
    if not vals.is_contiguous() or not b.is_contiguous():
        raise ValueError("A.values() and b must be contiguous for in-place scaling")
    m = int(A.shape[0])
    if sharded:
        import torch.distributed as dist

        bits = 32 if row.dtype == torch.int32 else 64
        sq = torch.empty(m, dtype=torch.float64, device=vals.device)
        with torch.cuda.device(vals.device):
            stream = torch.cuda.current_stream().cuda_stream
            _native.check(_native.lib().dualip_row_sq_norms(vals.data_ptr(), row.data_ptr(), bits, vals.numel(), m,
                                                            sq.data_ptr(), vals.device.index, stream))
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(sq)
            norms = sq.to(torch.float32).sqrt()
            rec = 1 / norms
            _native.check(_native.lib().dualip_scale_rows(vals.data_ptr(), row.data_ptr(), bits, vals.numel(),
                                                          rec.data_ptr(), vals.device.index, stream))
            b.mul_(rec)
        vals[:0].zero_()  # the kernel wrote through a raw pointer: bump the tensor's version counter (objectives check it)
        if norms_save_path:
            torch.save(norms, Path(norms_save_path))
        return norms
    norms = torch.empty(m, dtype=torch.float32, device=vals.device)
    with torch.cuda.device(vals.device):
        rc = _native.lib().dualip_jacobi_precondition(
            vals.data_ptr(), row.data_ptr(), 32 if row.dtype == torch.int32 else 64, vals.numel(), b.data_ptr(), m,
            norms.data_ptr(), vals.device.index, torch.cuda.current_stream().cuda_stream)
    _native.check(rc, "dualip_jacobi_precondition")
    vals[:0].zero_()  # in-place change made through a raw pointer: bump the version counter (objectives check it)
    b[:0].zero_()
    if norms_save_path:
        torch.save(norms, Path(norms_save_path))
    return norms


def jacobi_invert_precondition(dual_val: torch.Tensor, norms_path_or_tensor):
    """lambda_original = lambda_preconditioned / row_norms (reference precondition.py:32-60)."""
    if isinstance(norms_path_or_tensor, str):
        row_norms = torch.load(Path(norms_path_or_tensor), map_location=dual_val.device)
    elif isinstance(norms_path_or_tensor, torch.Tensor):
        row_norms = norms_path_or_tensor.to(dual_val.device)
    else:
        raise TypeError("norms_path_or_tensor must be a path or a tensor")
    return (1 / row_norms) * dual_val
