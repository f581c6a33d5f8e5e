"""Argument and result records of the solve API.

The reference defines five dataclasses (src/dualip/types.py:7-50); user code constructs them positionally and by keyword and
reads their attributes, so names, order and defaults of the reference's fields are kept exactly and are pinned by
tests/test_host_logic.py::test_dataclass_defaults_match_reference and by the reference's own tests
(tests/test_reference_suite_dropin.py).  Fields this package adds come after the reference's and default to None.
"""
from dataclasses import dataclass
from typing import Any, Dict, Literal, Optional

import torch


@dataclass
class SolverArgs:
    """Maximizer settings (reference types.py:7-16).

    gamma_decay_type "step": every gamma_decay_params["decay_steps"] iterations gamma is multiplied by
    gamma_decay_params["decay_factor"] and the step cap becomes step * factor (optimizers/agd.py:102-109).
    initial_dual_path: a torch.save'd dual vector to warm-start from (run_solver.py:121-126)."""

    max_iter: int = 10000
    initial_step_size: float = 1e-5
    gamma: float = 1e-3
    max_step_size: float = 0.1
    initial_dual_path: Optional[str] = None
    gamma_decay_type: Optional[Literal["step"]] = None
    gamma_decay_params: Optional[dict] = None
    save_primal: bool = False


@dataclass
class ComputeArgs:
    """Where to solve (reference types.py:19-22).  host_device must name a CUDA device here; compute_device_num > 1 means one
    process per GPU under torchrun, each taking its contiguous entity shard."""

    host_device: str
    compute_device_num: int = 1


@dataclass
class ObjectiveArgs:
    """Which objective to build (reference types.py:25-29): "matching" (fused slab kernel) or "miplib2017" (generic LP)."""

    objective_type: Literal["miplib2017", "matching"]
    use_jacobi_precondition: bool = False
    objective_kwargs: Optional[Dict[str, Any]] = None


@dataclass
class ObjectiveResult:
    """One evaluation of the dual (reference types.py:32-41); 0-dim float32 tensors for the scalars, like the reference."""

    dual_gradient: torch.Tensor                          # A x*(lambda) - b, m floats (local-shard mode: the raw partial sums)
    dual_objective: torch.Tensor                         # c.x + reg_penalty + lambda.(Ax - b)   (local-shard mode: c.x)
    reg_penalty: Optional[torch.Tensor] = None           # gamma/2 * ||x||^2
    primal_objective: Optional[torch.Tensor] = None      # c.x, set with save_primal
    primal_var: Optional[torch.Tensor] = None            # x in CSC value order, set with save_primal (a fresh buffer)
    dual_val_times_grad: Optional[torch.Tensor] = None   # lambda.(Ax - b)
    max_pos_slack: Optional[torch.Tensor] = None         # max(max(Ax - b), 0); always a tensor here
    sum_pos_slack: Optional[torch.Tensor] = None         # sum relu(Ax - b)
    # ---- additions of this package ----
    scalars64: Optional[torch.Tensor] = None        # the eight float64 scalars of dualip_scalars (include/dualip_b200.h)
    projection_diag: Optional[torch.Tensor] = None  # calculate(..., diagnostics=True): branch | rho << 2 per simplex column


@dataclass
class SolverResult:
    """What maximize / run_solver return (reference types.py:44-50)."""

    dual_val: torch.Tensor               # last projected iterate y (optimizers/agd.py:208-227), on every rank here
    dual_objective: float                # objective of the last evaluation
    objective_result: ObjectiveResult    # last evaluation (with primal_var if save_primal)
    dual_objective_log: list[float]      # one entry per iteration
    step_size_log: list[float]           # the Lipschitz-rule step of every iteration (optimizers/agd_utils.py:44-89)
