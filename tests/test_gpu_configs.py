"""-m gpu: the CUDA path on fixtures shaped like BASELINE.json's configurations, against outputs of the unmodified reference
(tests/golden/make_golden_configs.py).  configs[0]: MovieLens-shaped (a == 1, ratings, columns of 20..2600 entries: the
generic and the warp-per-column kernels).  configs[1..3]: the reference's synthetic generator, simplex, batching on and
off, the Maximizer from zero, Jacobi row scaling, and a warm start through run_solver(initial_dual_path=...)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.preprocessing.precondition import jacobi_precondition
from dualip_b200.projections import create_projection_map
from dualip_b200.run_solver import run_solver
from dualip_b200.types import ComputeArgs, ObjectiveArgs, SolverArgs
from test_config_golden import CFG1_MAPS, _check_trace
from test_gpu_parity import DEV, _csc

pytestmark = pytest.mark.gpu


def _check_calc(r, d, tag):
    assert np.array_equal(r.primal_var.cpu().numpy(), d[f"x_{tag}"]), "primal x differs from the reference"
    scal, got = d[f"scal_{tag}"], r.scalars64.cpu().numpy()  # reference order: dual_obj, reg, primal_obj, lam.grad, max, sum
    assert abs(got[0] - scal[0]) <= 1e-5 * abs(scal[0])
    assert abs(got[2] - scal[1]) <= 1e-5 * abs(scal[1]) + 1e-9
    assert abs(got[1] - scal[2]) <= 1e-5 * abs(scal[2])
    assert abs(got[4] - scal[4]) <= 1e-5 * max(1.0, abs(scal[4]))
    assert abs(got[5] - scal[5]) <= 1e-5 * max(1.0, abs(scal[5]))
    g, ref = r.dual_gradient.cpu().numpy(), d[f"grad_{tag}"]
    assert np.abs(g - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def _solve(A, C, pm, b, gamma, iters, start=None):
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=gamma, batching=False)
    solver = AcceleratedGradientDescent(max_iter=iters, gamma=gamma, initial_step_size=1e-3, max_step_size=1e-1,
                                        iteration_callback=lambda i, r: None)
    res = solver.maximize(obj, torch.zeros(b.numel(), device=DEV) if start is None else start)
    return res.dual_val.cpu().numpy(), res.dual_objective_log, res.step_size_log


@pytest.mark.parametrize("tag", ["box", "simplex"])
def test_movielens_shaped(tag):
    d = np.load(f"{GOLDEN}/cfg1_movielens_shaped.npz")
    ptype, params = CFG1_MAPS[tag]
    n, gamma = d["ccol"].size - 1, float(d["gamma"])
    A, C = _csc(d)
    b = torch.from_numpy(d["b"]).to(DEV)
    pm = create_projection_map(ptype, params, n)
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma=gamma)
    info = obj.plan_info()
    assert info["n_long_cols"] == 2 and info["launches_per_calc"] == 2
    _check_calc(obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True), d, tag)
    y, obj_log, step_log = _solve(A, C, pm, b, gamma, 30)
    _check_trace(y, obj_log, step_log, d, tag)


@pytest.mark.parametrize("batching", [True, False])
def test_reference_generator_calculate(batching):
    d = np.load(f"{GOLDEN}/cfg2_synthetic.npz")
    n, m, gamma = d["ccol"].size - 1, int(d["n_rows"]), float(d["gamma"])
    A, C = _csc(d)
    obj = MatchingSolverDualObjectiveFunction(
        MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), torch.from_numpy(d["b"]).to(DEV)),
        gamma=gamma, batching=batching)
    btag = "b1" if batching else "b0"
    _check_calc(obj.calculate(torch.zeros(m, device=DEV), save_primal=True), d, f"zero_{btag}")
    _check_calc(obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True), d, f"rand_{btag}")


def test_reference_generator_ascent_warm_start_and_jacobi(tmp_path):
    d = np.load(f"{GOLDEN}/cfg2_synthetic.npz")
    n, gamma = d["ccol"].size - 1, float(d["gamma"])
    A, C = _csc(d)
    b = torch.from_numpy(d["b"]).to(DEV)
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    y, obj_log, step_log = _solve(A, C, pm, b, gamma, 40)
    # from iteration 15 on the Lipschitz step overshoots on this problem (the reference's own log zig-zags) and the 1e-7
    # summation-order difference of the gradient grows ~1.5x per iteration: tight bar on the stable part, 2e-3 after
    _check_trace(y, obj_log, step_log, d, "plain", tight=26)
    # configs[3]: warm start from the reference's own saved dual, through run_solver like the reference (run_solver.py:121-126)
    path = str(tmp_path / "dual.pt")
    torch.save(torch.from_numpy(d["plain_dual"].copy()), path)
    res = run_solver(MatchingInputArgs(A.cpu(), C.cpu(), pm, b.cpu()),
                     SolverArgs(max_iter=20, gamma=gamma, initial_step_size=1e-3, max_step_size=1e-1, initial_dual_path=path),
                     ComputeArgs(host_device=DEV), ObjectiveArgs(objective_type="matching", objective_kwargs={"batching": False}))
    _check_trace(res.dual_val.cpu().numpy(), res.dual_objective_log, res.step_size_log, d, "warm", tight=15)
    # configs[2]: Jacobi row scaling on the device, then the same ascent
    A2, _ = _csc(d)
    b2 = b.clone()
    norms = jacobi_precondition(A2, b2)
    assert np.allclose(norms.cpu().numpy(), d["jacobi_norms"], rtol=2e-6)
    assert np.allclose(A2.values().cpu().numpy(), d["jacobi_a"], rtol=2e-6) and np.allclose(b2.cpu().numpy(), d["jacobi_b"], rtol=2e-6)
    # the ascent itself from the reference's scaled values, so that the trace comparison starts from identical inputs
    A3, _ = _csc(dict(ccol=d["ccol"], row=d["row"], a=d["jacobi_a"], c=d["c"], n_rows=d["n_rows"]))
    y, obj_log, step_log = _solve(A3, C, pm, torch.from_numpy(d["jacobi_b"]).to(DEV), gamma, 40)
    _check_trace(y, obj_log, step_log, d, "jacobi", tight=26)


def test_bisection_method_in_the_fused_kernel():
    """`method="bisection_search"` (reference projections/simplex.py:6-123) runs inside the fused kernel and in
    dualip_project_block, restated per column: no pre-clamp, sums down the rows of the zero-padded block in fp32, the same
    19 halvings of [-1, 0] for every column.  Bit-identical to the reference on padded blocks (simplex / simplex_eq, z = 1 and
    2.5) and through the objective, where the padded length of a column's bucket is part of the result (batching on / off)."""
    from dualip_b200.projections import project

    d = np.load(f"{GOLDEN}/projection_bisection.npz")
    for L in (1, 2, 7, 16, 33):
        x = torch.from_numpy(d[f"L{L}_x"]).to(DEV)
        for name in ("simplex", "simplex_eq"):
            for z in (1.0, 2.5):
                out = project(name, z=z, method="bisection_search")(x).cpu().numpy()
                assert np.array_equal(out, d[f"L{L}_{name}_z{z}"]), (L, name, z)
    n, m, gamma = d["ccol"].size - 1, int(d["n_rows"]), float(d["gamma"])
    assert not np.array_equal(d["x_b1"], d["x_b0"])
    for index_dtype in (torch.int64, torch.int32):
        A, C = _csc(d, index_dtype=index_dtype)
        pm = create_projection_map("simplex", {"z": 1.0, "method": "bisection_search"}, n)
        for batching, tag in ((True, "b1"), (False, "b0")):
            obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(d["b"]).to(DEV)), gamma=gamma,
                                                      batching=batching)
            assert not obj.has_block_entries, "bisection columns must be projected by the fused kernel"
            r = obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True)
            assert np.array_equal(r.primal_var.cpu().numpy(), d[f"x_{tag}"])
            scal, got = d[f"scal_{tag}"], r.scalars64.cpu().numpy()
            assert abs(got[0] - scal[0]) <= 1e-5 * abs(scal[0])
            g, ref = r.dual_gradient.cpu().numpy(), d[f"grad_{tag}"]
            assert np.abs(g - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
            r2 = obj.calculate(torch.from_numpy(d["lam"]).to(DEV))  # the launch without primal output takes the same branch
            assert torch.allclose(r2.dual_gradient, r.dual_gradient, rtol=1e-6, atol=1e-6)
    # mixed with a Duchi entry and under the Maximizer (one launch per iteration)
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0, "method": "bisection_search"}, n, indices=list(range(0, n, 2))))
    pm.update(create_projection_map("simplex", {"z": 1.0}, n, indices=list(range(1, n, 2))))
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(d["b"]).to(DEV)), gamma=gamma)
    out = AcceleratedGradientDescent(max_iter=30, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                     iteration_callback=no_iteration_callback).maximize(obj, torch.zeros(m, device=DEV))
    assert all(np.isfinite(out.dual_objective_log)) and out.dual_objective_log[-1] > out.dual_objective_log[0]


def test_one_launch_iteration_equals_two_launches(monkeypatch):
    """dualip_matching_ascent_step (evaluation + accelerated step in one kernel launch, the default of the fused loop) against
    dualip_matching_calc followed by dualip_agd_step: same logs and iterate, with step gamma decay and save_primal."""
    import os

    from conftest import random_problem
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
    from dualip_b200.projections import create_projection_map

    p = random_problem(5, 20011, 300, 9.0, scale_c=10.0, lam_scale=0.5, long_cols=((3, 40), (77, 150)))
    n, m = p["n_cols"], p["n_rows"]
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]), size=(m, n)).to("cuda:0")
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]), size=(m, n)).to("cuda:0")
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0}, n, indices=list(range(0, n, 2))))
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n, indices=list(range(1, n, 2))))
    out = {}
    seen = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DUALIP_ONE_LAUNCH", mode)
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(p["b"]).to("cuda:0")), gamma=2e-2)
        seen[mode] = []
        solver = AcceleratedGradientDescent(max_iter=45, gamma=2e-2, initial_step_size=1e-3, max_step_size=0.1, save_primal=True,
                                            gamma_decay_type="step", gamma_decay_params={"decay_steps": 10, "decay_factor": 0.5},
                                            iteration_callback=lambda i, r, mode=mode: seen[mode].append(float(r.dual_objective)))
        out[mode] = solver.maximize(obj, torch.zeros(m, device="cuda:0"))
    a, b = out["1"], out["0"]
    assert np.allclose(a.dual_objective_log, b.dual_objective_log, rtol=1e-6) and np.allclose(a.step_size_log, b.step_size_log, rtol=1e-5)
    assert torch.allclose(a.dual_val, b.dual_val, rtol=1e-4, atol=1e-6)
    assert torch.allclose(a.objective_result.primal_var, b.objective_result.primal_var, rtol=1e-4, atol=1e-6)
    # the callback saw every iteration's result, as written by the kernel before it stepped
    assert np.allclose(seen["1"], a.dual_objective_log, rtol=1e-6) and len(seen["1"]) == 45
