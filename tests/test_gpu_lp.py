"""-m gpu: the generic-LP objective (dualip_lp_calc behind MIPLIB2017ObjectiveFunction) against outputs of the reference's
own class on the shipped MIPLIB-2017 instance and on a derived LP with equality rows, cone bounds and Jacobi scaling
(tests/golden/lp_*.npz), through calculate(), the fused Maximizer and run_solver."""
import os

import numpy as np
import pytest
import torch

from dualip_b200.objectives.miplib import MIPLIB2017ObjectiveFunction, MIPLIBInputArgs
from dualip_b200.optimizers.agd import AcceleratedGradientDescent
from dualip_b200.projections import create_projection_map
from dualip_b200.run_solver import run_solver
from dualip_b200.types import ComputeArgs, ObjectiveArgs, SolverArgs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _projection_map(d):
    """Box / cone entries equivalent to the stored per-variable bounds."""
    lower, upper = d["lower"], d["upper"]
    n = lower.size
    pm, groups = {}, {}
    for j in range(n):
        groups.setdefault((float(lower[j]), float(upper[j])), []).append(j)
    for k, ((lo, hi), idx) in enumerate(groups.items()):
        if np.isinf(lo) and np.isinf(hi):
            continue
        if np.isinf(hi):
            pm.update(create_projection_map("cone", {"lower": lo}, n, indices=idx, key_prefix=f"g{k}_"))
        elif np.isinf(lo):
            pm.update(create_projection_map("cone", {"upper": hi}, n, indices=idx, key_prefix=f"g{k}_"))
        else:
            pm.update(create_projection_map("box", {"lower": lo, "upper": hi}, n, indices=idx, key_prefix=f"g{k}_"))
    return pm


def _input_args(d, sparse):
    A = torch.from_numpy(d["A"]).to(DEV)
    eq = torch.from_numpy(d["eq_mask"]).to(DEV) if d["eq_mask"].size else None
    return MIPLIBInputArgs(A=A.to_sparse() if sparse else A, c=torch.from_numpy(d["c"]).to(DEV), projection_map=_projection_map(d),
                           b_vec=torch.from_numpy(d["b"]).to(DEV), equality_mask=eq)


@pytest.mark.parametrize("name", ["lp_miplib", "lp_eq_cone"])
def test_lp_calculate_against_reference_outputs(name):
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    jacobi = bool(d["jacobi"])
    obj = MIPLIB2017ObjectiveFunction(_input_args(d, sparse=not jacobi), use_jacobi_precondition=jacobi)
    for k, lam in enumerate(d["lams"]):
        r = obj.calculate(torch.from_numpy(lam).to(DEV), gamma=float(d["gamma"]), save_primal=True)
        x = r.primal_var.cpu().numpy()
        assert np.allclose(x, d[f"x{k}"], rtol=1e-5, atol=1e-5)
        # projection index selection: the same variables sit on their bounds
        assert np.array_equal(x == d["lower"], d[f"x{k}"] == d["lower"]) and np.array_equal(x == d["upper"], d[f"x{k}"] == d["upper"])
        scale = max(1.0, float(np.abs(d[f"grad{k}"]).max()))
        assert np.allclose(r.dual_gradient.cpu().numpy(), d[f"grad{k}"], rtol=1e-5, atol=1e-5 * scale)
        ref_obj, ref_reg, ref_primal = d[f"scal{k}"]
        s = r.scalars64.cpu().numpy()
        assert abs(s[0] - ref_obj) <= 1e-5 * max(1.0, abs(ref_obj))
        assert abs(s[2] - ref_reg) <= 1e-5 * max(1.0, abs(ref_reg)) and abs(s[1] - ref_primal) <= 1e-5 * max(1.0, abs(ref_primal))
        # evaluations do not leak state into each other
        r2 = obj.calculate(torch.from_numpy(lam).to(DEV), gamma=float(d["gamma"]))
        assert torch.equal(r2.dual_gradient, r.dual_gradient)


@pytest.mark.parametrize("name", ["lp_miplib", "lp_eq_cone"])
def test_lp_ascent_trace_against_reference_run(name):
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    jacobi = bool(d["jacobi"])
    obj = MIPLIB2017ObjectiveFunction(_input_args(d, sparse=not jacobi), use_jacobi_precondition=jacobi)
    steps, factor = int(d["decay"][0]), float(d["decay"][1])
    solver = AcceleratedGradientDescent(max_iter=int(d["iters"]), gamma=float(d["gamma"]), initial_step_size=1e-3, max_step_size=0.1,
                                        gamma_decay_type="step", gamma_decay_params={"decay_steps": steps, "decay_factor": factor},
                                        iteration_callback=lambda i, r: None)
    res = solver.maximize(obj, torch.zeros(d["b"].size, device=DEV))
    ref, got = d["obj_log"], np.array(res.dual_objective_log)
    # same tolerances as the oracle's own trace test (tests/test_lp_oracle_golden.py): the ascent amplifies rounding
    assert np.allclose(got[:12], ref[:12], rtol=1e-4, atol=1e-4 * np.abs(ref[:12]).max())
    assert np.abs(got - ref).max() <= 5e-2 * np.abs(ref).max()
    assert abs(solver.gamma - float(d["gamma"]) * factor ** (int(d["iters"]) // steps)) < 1e-12


def test_run_solver_miplib2017_objective_type():
    d = np.load(os.path.join(GOLDEN, "lp_miplib.npz"))
    args = _input_args(d, sparse=True)
    res = run_solver(args, SolverArgs(max_iter=300, initial_step_size=1e-5, gamma=1e-3), ComputeArgs(host_device=DEV),
                     ObjectiveArgs(objective_type="miplib2017"))
    assert len(res.dual_objective_log) == 300 and np.isfinite(res.dual_objective_log).all()
    assert res.dual_objective_log[-1] > res.dual_objective_log[1]  # ascent
    obj = MIPLIB2017ObjectiveFunction(args)
    gap, _, primal_feas, dual_feas, converged = obj.calculate_convergence_bound(res.dual_val)
    assert all(np.isfinite(float(v)) for v in (gap, primal_feas, dual_feas)) and isinstance(converged, bool)


def test_lp_rejects_non_elementwise_projection_and_cpu_tensors():
    d = np.load(os.path.join(GOLDEN, "lp_eq_cone.npz"))
    args = _input_args(d, sparse=False)
    args.projection_map = create_projection_map("simplex", {"z": 1.0}, d["c"].size)
    with pytest.raises(ValueError):
        MIPLIB2017ObjectiveFunction(args)
    cpu = MIPLIBInputArgs(A=torch.from_numpy(d["A"]), c=torch.from_numpy(d["c"]), projection_map={}, b_vec=torch.from_numpy(d["b"]))
    with pytest.raises(ValueError):
        MIPLIB2017ObjectiveFunction(cpu)
