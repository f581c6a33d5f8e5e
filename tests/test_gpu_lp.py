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
from oracle import dualip_oracle as O

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
    assert np.allclose(got[:12], ref[:12], rtol=1e-4, atol=1e-4 * np.abs(ref[:12]).max())
    assert abs(solver.gamma - float(d["gamma"]) * factor ** (int(d["iters"]) // steps)) < 1e-12
    # The ascent on these LPs amplifies rounding (clamp pattern switches; the Lipschitz step divides differences of
    # gradients), so two correct fp32 implementations drift apart.  The bar is therefore anchored on the fp64 oracle: the
    # CUDA trace may be no further from the fp64 trace than the reference's own fp32 run is (2x + fp32 resolution).
    A, c, b = d["A"], d["c"], d["b"]
    eq = d["eq_mask"] if d["eq_mask"].size else None
    row_norms = None
    if jacobi:
        row_norms = np.linalg.norm(A.astype(np.float32), axis=1).astype(np.float32)
        row_norms[row_norms == 0] = 1.0

    def calc64(lam, gamma):
        grad, obj, *_ = O.lp_calculate(A, c, b, d["lower"], d["upper"], lam, gamma, row_norms, dtype=np.float64)
        return grad, obj

    _, t64, _, _ = O.agd_maximize(calc64, np.zeros(b.size, dtype=np.float64), int(d["iters"]), float(d["gamma"]), 1e-3, 0.1,
                                  gamma_decay_type="step", gamma_decay_params={"decay_steps": steps, "decay_factor": factor},
                                  equality_mask=eq)
    t64 = np.array(t64)
    scale = np.abs(t64).max()
    assert np.abs(got - t64).max() <= 2.0 * np.abs(ref - t64).max() + 1e-5 * scale


@pytest.mark.parametrize("name", ["lp_miplib", "lp_eq_cone"])
def test_lp_kernel_in_lockstep_with_the_oracle_along_the_whole_ascent(name):
    """Every evaluation of a full ascent (step gamma decay included), at the oracle's own iterates.  The objective agrees to
    1e-5.  The gradient is ill-conditioned where variables sit strictly inside their bounds (x = -(A^T lambda + c)/gamma
    amplifies the rounding of the column sums by 1/gamma = 1e3), so fp32 implementations with different summation orders
    differ by more than 1e-5 there; the bar is the fp64 evaluation at the same point: the CUDA path may be no further from
    it than the reference-equivalent fp32 evaluation is (4x + fp32 resolution)."""
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    jacobi = bool(d["jacobi"])
    obj = MIPLIB2017ObjectiveFunction(_input_args(d, sparse=not jacobi), use_jacobi_precondition=jacobi)
    A, c, b = d["A"], d["c"], d["b"]
    eq = d["eq_mask"] if d["eq_mask"].size else None
    steps, factor = int(d["decay"][0]), float(d["decay"][1])
    row_norms = None
    if jacobi:
        row_norms = np.linalg.norm(A.astype(np.float32), axis=1).astype(np.float32)
        row_norms[row_norms == 0] = 1.0
    worst = {"g_cuda": 0.0, "g_fp32": 0.0, "o": 0.0, "excess": 0.0}

    def calc(lam, gamma):
        grad, dual_obj, *_ = O.lp_calculate(A, c, b, d["lower"], d["upper"], lam, gamma, row_norms)
        g64, o64, *_ = O.lp_calculate(A, c, b, d["lower"], d["upper"], lam, gamma, row_norms, dtype=np.float64)
        r = obj.calculate(torch.from_numpy(lam).to(DEV), gamma=gamma)
        g = r.dual_gradient.cpu().numpy()
        gs = max(1.0, float(np.abs(g64).max()))
        e_cuda, e_fp32 = float(np.abs(g - g64).max()) / gs, float(np.abs(grad - g64).max()) / gs
        worst["g_cuda"], worst["g_fp32"] = max(worst["g_cuda"], e_cuda), max(worst["g_fp32"], e_fp32)
        worst["excess"] = max(worst["excess"], e_cuda - (4.0 * e_fp32 + 1e-5))
        worst["o"] = max(worst["o"], abs(float(r.scalars64[0]) - o64) / max(1.0, abs(o64)))
        return grad, dual_obj

    O.agd_maximize(calc, np.zeros(b.size, dtype=np.float32), int(d["iters"]), float(d["gamma"]), 1e-3, 0.1,
                   gamma_decay_type="step", gamma_decay_params={"decay_steps": steps, "decay_factor": factor}, equality_mask=eq)
    assert worst["excess"] <= 0.0 and worst["o"] <= 1e-5, worst


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


def test_miplib_example_end_to_end_sanity_value():
    """The reference's own acceptance check for this objective (examples/miplib_2017/solve_miplib_dataset.py:53-70):
    v150d30-2hopcds through run_solver("miplib2017") with max_iter=10000, initial_step_size=1e-5, gamma=1e-3 ends within
    1 of the LP optimum 27."""
    d = np.load(os.path.join(GOLDEN, "lp_miplib.npz"))
    args = _input_args(d, sparse=True)
    res = run_solver(args, SolverArgs(max_iter=10000, initial_step_size=1e-5, gamma=1e-3), ComputeArgs(host_device=DEV),
                     ObjectiveArgs(objective_type="miplib2017"))
    assert abs(27 - res.dual_objective) < 1, res.dual_objective
    assert len(res.dual_objective_log) == 10000


def test_lp_sparse_input_is_never_densified(monkeypatch):
    """A sparse constraint matrix stays sparse (reference miplib.py:41-42 keeps CSR + CSC copies): construction, evaluation and
    the convergence check work with Tensor.to_dense disabled, and agree with the dense-input objective."""
    d = np.load(os.path.join(GOLDEN, "lp_eq_cone.npz"))
    dense_obj = MIPLIB2017ObjectiveFunction(_input_args(d, sparse=False))
    lam = torch.from_numpy(d["lams"][1]).to(DEV)
    want = dense_obj.calculate(lam, gamma=float(d["gamma"]), save_primal=True)
    want_bound = dense_obj.calculate_convergence_bound(lam, x=want.primal_var, tol=1e-4)

    def boom(self, *a, **k):
        raise AssertionError("to_dense() called on a sparse LP")

    monkeypatch.setattr(torch.Tensor, "to_dense", boom)
    obj = MIPLIB2017ObjectiveFunction(_input_args(d, sparse=True))
    got = obj.calculate(lam, gamma=float(d["gamma"]))
    assert torch.allclose(got.dual_gradient, want.dual_gradient, rtol=1e-6, atol=1e-6)
    got_bound = obj.calculate_convergence_bound(lam, x=want.primal_var, tol=1e-4)
    for a, b in zip(got_bound[:4], want_bound[:4]):
        assert torch.allclose(torch.as_tensor(a, dtype=torch.float32).cpu(), torch.as_tensor(b, dtype=torch.float32).cpu(),
                              rtol=1e-4, atol=1e-6, equal_nan=True)


def test_lp_projection_entries_apply_in_order():
    """Entries of the projection map are applied one after the other (reference miplib.py:80-90): a later clamp acts on
    the result of an earlier one, also when their intervals are disjoint."""
    d = np.load(os.path.join(GOLDEN, "lp_eq_cone.npz"))
    n = d["c"].size
    args = _input_args(d, sparse=False)
    pm = {}
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n, indices=list(range(n))))
    pm.update({"late": type(next(iter(pm.values())))(proj_type="box", proj_params={"lower": 2.0, "upper": 3.0}, indices=[0, 1])})
    args.projection_map = pm
    obj = MIPLIB2017ObjectiveFunction(args)
    r = obj.calculate(torch.zeros(d["b"].size, device=DEV), gamma=1e-2, save_primal=True)
    x = r.primal_var.cpu().numpy()
    assert x[0] == 2.0 and x[1] == 2.0 and (x[2:] <= 1.0).all() and (x[2:] >= 0.0).all()


def test_lp_rejects_non_elementwise_projection_and_cpu_tensors():
    d = np.load(os.path.join(GOLDEN, "lp_eq_cone.npz"))
    args = _input_args(d, sparse=False)
    args.projection_map = create_projection_map("simplex", {"z": 1.0}, d["c"].size)
    with pytest.raises(ValueError):
        MIPLIB2017ObjectiveFunction(args)
    cpu = MIPLIBInputArgs(A=torch.from_numpy(d["A"]), c=torch.from_numpy(d["c"]), projection_map={}, b_vec=torch.from_numpy(d["b"]))
    with pytest.raises(ValueError):
        MIPLIB2017ObjectiveFunction(cpu)
