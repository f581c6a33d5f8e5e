"""Solve API: `run_solver(input_args, solver_args, compute_args, objective_args, mlflow_config)` with the reference's
signature and types (src/dualip/run_solver.py:17-146).

`compute_device_num > 1` works under torchrun (one process per GPU): the reference's own multi-device branch passes
keyword arguments its distributed class does not accept (run_solver.py:62-67 vs matching.py:218-225)."""
from dataclasses import fields
from typing import Optional

import torch
import torch.distributed as dist

from dualip_b200.objectives.base import BaseInputArgs
from dualip_b200.objectives.matching import (
    MatchingInputArgs,
    MatchingSolverDualObjectiveFunction,
    MatchingSolverDualObjectiveFunctionDistributed,
)
from dualip_b200.optimizers.agd import AcceleratedGradientDescent
from dualip_b200.types import ComputeArgs, ObjectiveArgs, SolverArgs, SolverResult
from dualip_b200.utils.dist_utils import global_to_local_projection_map, shard_sizes
from dualip_b200.utils.mlflow_utils import MLflowConfig
from dualip_b200.utils.sparse_utils import split_csc_by_cols


def transfer_tensors_to_device(input_args: BaseInputArgs, device: str):
    """New instance of input_args with every tensor field moved to `device` (reference run_solver.py:17-41)."""
    moved = {}
    for f in fields(input_args):
        value = getattr(input_args, f.name)
        moved[f.name] = value.to(device) if isinstance(value, torch.Tensor) else value
    return type(input_args)(**moved)


def _shard_matching_args(input_args: MatchingInputArgs, rank: int, world: int, device) -> MatchingInputArgs:
    """This rank's contiguous column shard (sizes as reference utils/dist_utils.py:53-57), moved to `device`."""
    n = input_args.A.size(1)
    sizes = shard_sizes(n, world)
    start = sum(sizes[:rank])
    a_local = split_csc_by_cols(input_args.A, sizes)[rank].to(device)
    c_local = split_csc_by_cols(input_args.c, sizes)[rank].to(device)
    pm_local = global_to_local_projection_map(input_args.projection_map, range(start, start + sizes[rank]))
    mask = input_args.equality_mask.to(device) if input_args.equality_mask is not None else None
    return MatchingInputArgs(A=a_local, c=c_local, projection_map=pm_local, b_vec=None, equality_mask=mask)


def build_objective(input_args: BaseInputArgs, solver_args: SolverArgs, compute_args: ComputeArgs,
                    objective_args: ObjectiveArgs):
    objective_type = objective_args.objective_type
    kwargs = objective_args.objective_kwargs or {}
    if objective_type == "matching":
        if compute_args.compute_device_num == 1:
            return MatchingSolverDualObjectiveFunction(matching_input_args=input_args, gamma=solver_args.gamma, **kwargs)
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("compute_device_num > 1 needs torch.distributed (launch one process per GPU with torchrun)")
        world, rank = dist.get_world_size(), dist.get_rank()
        if world != compute_args.compute_device_num:
            raise ValueError(f"compute_device_num={compute_args.compute_device_num} but world size is {world}")
        device = torch.device("cuda", torch.cuda.current_device())
        local_args = _shard_matching_args(input_args, rank, world, device)
        return MatchingSolverDualObjectiveFunctionDistributed(
            local_matching_input_args=local_args, b_vec=input_args.b_vec, gamma=solver_args.gamma,
            host_device=compute_args.host_device, **kwargs)
    if objective_type == "miplib2017":
        from dualip_b200.objectives.miplib import MIPLIB2017ObjectiveFunction

        return MIPLIB2017ObjectiveFunction(miplib_input_args=input_args, **kwargs)
    raise ValueError(f"Objective type {objective_type} not supported")


def run_solver(input_args: BaseInputArgs, solver_args: SolverArgs, compute_args: ComputeArgs,
               objective_args: ObjectiveArgs, mlflow_config: Optional[MLflowConfig] = None) -> SolverResult:
    """Run the LP solver with the given configuration (reference run_solver.py:74-146)."""
    if mlflow_config is not None and mlflow_config.enabled:
        raise NotImplementedError("MLflow logging is out of scope for dualip_b200; pass mlflow_config=None")
    sharded = objective_args.objective_type == "matching" and compute_args.compute_device_num > 1
    rank = dist.get_rank() if (sharded and dist.is_initialized()) else 0
    if sharded:
        # keep the full problem where the caller put it; only this rank's shard goes to its GPU
        device = torch.device("cuda", torch.cuda.current_device())
    else:
        device = torch.device(compute_args.host_device)
        input_args = transfer_tensors_to_device(input_args, compute_args.host_device)
    objective = build_objective(input_args, solver_args, compute_args, objective_args)
    solver = AcceleratedGradientDescent(
        initial_step_size=solver_args.initial_step_size,
        max_iter=solver_args.max_iter,
        max_step_size=solver_args.max_step_size,
        gamma=solver_args.gamma,
        gamma_decay_type=solver_args.gamma_decay_type,
        gamma_decay_params=solver_args.gamma_decay_params,
        save_primal=solver_args.save_primal,
    )
    if solver_args.initial_dual_path is not None:  # warm start
        initial_dual = torch.load(solver_args.initial_dual_path, map_location="cpu")
    else:
        initial_dual = torch.zeros_like(input_args.b_vec, device="cpu")
    initial_dual = initial_dual.to(device=device, dtype=torch.float32)
    result = solver.maximize(objective, initial_dual, rank=rank)
    if getattr(objective, "use_jacobi_precondition", False) and hasattr(objective, "invert_jacobi_precondition"):
        # the reference calls a method it never defines here (run_solver.py:136-144)
        result.dual_val = objective.invert_jacobi_precondition(result.dual_val)
    return result
