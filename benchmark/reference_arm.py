"""Times the UNMODIFIED reference (oracle/_ref, populated by oracle/make_ref.py) on a workload of this benchmark.

Baseline infrastructure: nothing under dualip_b200/ imports this.  The reference's own objects are used end to end —
`MatchingInputArgs`, `create_projection_map`, `jacobi_precondition`, `MatchingSolverDualObjectiveFunction`,
`AcceleratedGradientDescent.maximize` (reference benchmark/run_matching_benchmark.py:81-107) — on `torch.sparse_csc`
inputs with int64 indices like its generator produces (generate_synthetic_data.py:135,139).  Only the input data come
from this repository's vectorised generator (benchmark/synthetic.py: same distributions; the reference's generator is a
Python loop that needs ~15 min and >40 GB at 100M entities).
"""
from __future__ import annotations

import os
import time

GAMMA = 1e-3
INITIAL_STEP, MAX_STEP = 1e-3, 1e-1  # reference benchmark/config.py:17-18


def host_threads() -> int:
    """Hardware threads this process may use (ignores OMP_NUM_THREADS, which torchrun forces to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def build_reference_problem(n_total, m, sparsity, seed, n_sample, mixed, jacobi, device, gen_device=None):
    """First `n_sample` entities of the (n_total x m) workload as the reference's MatchingInputArgs on `device`."""
    import torch

    from benchmark.synthetic import capacity_vector, generate_shard
    from oracle import make_ref

    make_ref.import_reference()
    from dualip.objectives.matching import MatchingInputArgs
    from dualip.preprocessing.precondition import jacobi_precondition
    from dualip.projections.base import create_projection_map

    gen_device = gen_device or ("cuda" if torch.cuda.is_available() else "cpu")
    shard = generate_shard(n_total, m, sparsity, seed, gen_device, 0, n_sample)
    # capacities of the full problem: scale the sample's greedy load
    b = capacity_vector(shard.greedy_load * (n_total / n_sample), m, sparsity, seed, gen_device)
    dev = torch.device(device)
    ccol, row = shard.ccol.to(dev), shard.row.to(dev)
    A = torch.sparse_csc_tensor(ccol, row, shard.a.to(dev), size=(m, n_sample))
    C = torch.sparse_csc_tensor(ccol, row, shard.c.to(dev), size=(m, n_sample))
    b = b.to(dev)
    if jacobi:
        jacobi_precondition(A, b)  # reference preprocessing/precondition.py:8-29, in place
    if mixed:
        pm = {}
        pm.update(create_projection_map("simplex", {"z": 1.0}, n_sample, indices=list(range(0, n_sample, 2))))
        pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n_sample, indices=list(range(1, n_sample, 2))))
    else:
        pm = create_projection_map("simplex", {"z": 1.0}, n_sample)
    args = MatchingInputArgs(A=A, c=C, projection_map=pm, b_vec=b, equality_mask=None)
    return args, int(row.numel())


def time_reference_maximize(input_args, steps, warmup, batching, device, sync=None):
    """W warm-up + K timed iterations of the reference's maximize(); returns (seconds for the K iterations, result)."""
    import torch

    from oracle import make_ref

    make_ref.import_reference()
    from dualip.objectives.matching import MatchingSolverDualObjectiveFunction
    from dualip.optimizers.agd import AcceleratedGradientDescent

    is_cuda = torch.device(device).type == "cuda"
    sync = sync or ((lambda: torch.cuda.synchronize()) if is_cuda else (lambda: None))
    t_obj0 = time.perf_counter()
    objective = MatchingSolverDualObjectiveFunction(matching_input_args=input_args, gamma=GAMMA, batching=batching)
    sync()
    t_obj = time.perf_counter() - t_obj0
    marks = {}

    def callback(i, result):  # the reference calls this once per iteration (agd.py:165); no printing, no .item()
        if i == warmup:
            sync()
            marks["t0"] = time.perf_counter()

    solver = AcceleratedGradientDescent(max_iter=warmup + steps, gamma=GAMMA, initial_step_size=INITIAL_STEP,
                                        max_step_size=MAX_STEP, iteration_callback=callback)
    result = solver.maximize(objective, torch.zeros_like(input_args.b_vec))
    sync()
    dt = time.perf_counter() - marks["t0"]
    return dt, t_obj, result
