"""Input checks for LP data with the reference's function names, exception type and messages
(src/dualip/preprocessing/input_validation.py:4-103), vectorised: the reference walks the columns of a CSC matrix in a
Python loop with one device synchronisation per column (:66-74), which is not usable at 10^8 entities; here the row ordering
of all columns is checked with one pass over the row indices.  Setup-time code, runs on whatever device the tensor is on."""
import torch


class InputValidationError(ValueError):
    """Raised when any of the checks below fails."""


def _csc_parts(t: torch.Tensor):
    return t.ccol_indices(), t.row_indices(), t.values()


def check_no_zero_row_or_col(input_tensor: torch.Tensor) -> None:
    """No all-zero row (both layouts) and, for dense tensors, no all-zero column (reference :8-33)."""
    if input_tensor.layout is torch.strided:
        nonzero = input_tensor != 0
        if bool((~nonzero.any(dim=0)).any()):
            raise InputValidationError("There is an all-zero column in the input tensor")
        if bool((~nonzero.any(dim=1)).any()):
            raise InputValidationError("There is an all-zero row in the input tensor")
        return
    _, row, _ = _csc_parts(input_tensor)
    seen = torch.zeros(input_tensor.shape[0], dtype=torch.bool, device=row.device)
    seen[row.to(torch.int64)] = True
    if not bool(seen.all()):
        raise InputValidationError("There is an all-zero row in the input tensor")


def check_nan_or_inf(input_tensor: torch.Tensor) -> None:
    """No NaN / +-Inf among the (stored) values (reference :36-49)."""
    vals = input_tensor.values() if input_tensor.layout is torch.sparse_csc else input_tensor
    if not bool(torch.isfinite(vals).all()):
        raise InputValidationError("The input tensor has nan or infinite values")


def check_correct_csc_construction(input_tensor: torch.Tensor) -> None:
    """Column pointers non-decreasing, row indices strictly increasing inside every column, no explicit zeros
    (reference :52-79)."""
    assert input_tensor.layout is torch.sparse_csc
    ccol, row, vals = _csc_parts(input_tensor)
    if bool((ccol[:-1] > ccol[1:]).any()):
        raise InputValidationError("ccol_indices must be non-decreasing")
    nnz = row.numel()
    if nnz > 1:
        not_increasing = row[1:] <= row[:-1]  # entry e+1 against entry e
        # ... which is fine when entry e+1 opens a new column
        opens = torch.zeros(nnz + 1, dtype=torch.bool, device=row.device)
        opens[ccol.to(torch.int64)] = True
        bad = not_increasing & ~opens[1:nnz]
        if bool(bad.any()):
            pos = int(torch.nonzero(bad)[0]) + 1
            col = int(torch.searchsorted(ccol.to(torch.int64), torch.tensor(pos, device=ccol.device), right=True)) - 1
            raise InputValidationError(f"row indices in column {col} are not strictly increasing")
    if bool((vals == 0).any()):
        raise InputValidationError("No zeroes are allowed in CSC values component")


def check_projection_map():
    raise NotImplementedError("Checking the projection map is not yet implemented")  # as in the reference (:82-85)


def run_all_checks(input_tensor: torch.Tensor) -> None:
    """The standard checks for an LP matrix in strided or CSC layout (reference :88-103)."""
    assert input_tensor.layout is torch.strided or input_tensor.layout is torch.sparse_csc
    if input_tensor.layout is torch.sparse_csc:
        check_correct_csc_construction(input_tensor)
    check_no_zero_row_or_col(input_tensor)
    check_nan_or_inf(input_tensor)
