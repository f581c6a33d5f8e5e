"""-m gpu: CUDA-graph replay of the ascent loop and the per-row fixed-point scaling.

The graph path (dualip_agd_set_schedule / dualip_ascent_graph_*) must be indistinguishable from launching the iterations
one by one: same kernel, same arithmetic, the per-iteration scalars merely come from a device-resident schedule.  The
reference loop is optimizers/agd.py:150-206."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import random_problem
from dualip_b200 import _native
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, FusedAscentLoop, no_iteration_callback
from dualip_b200.projections import create_projection_map
from oracle import dualip_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _objective(p, pm, gamma):
    m, n = int(p["n_rows"]), p["ccol"].size - 1
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]), size=(m, n)).to(DEV)
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]), size=(m, n)).to(DEV)
    return MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(p["b"]).to(DEV)), gamma=gamma)


def _mixed_map(n):
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0}, n, indices=list(range(0, n, 2))))
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n, indices=list(range(1, n, 2))))
    return pm


@pytest.mark.parametrize("decay", [False, True])
@pytest.mark.parametrize("save_primal", [False, True])
def test_graph_replay_is_bit_identical_to_single_launches(monkeypatch, decay, save_primal):
    """maximize() with graph replay (chunks of 8) against the same run launched one iteration at a time: dual iterate,
    objective log, step-size log and final gradient must agree bit for bit, with and without gamma decay (the schedule
    carries gamma, beta and the step-cap flag), and with a last iteration that writes the primal (never part of a graph)."""
    monkeypatch.setenv("DUALIP_REBALANCE", "0")  # the plan is settled from the first launch: graphs start at iteration 1
    p = random_problem(11, 6000, 96, 8.0)
    n = p["ccol"].size - 1
    kw = dict(max_iter=53, gamma=2e-2, initial_step_size=1e-3, max_step_size=0.1, iteration_callback=no_iteration_callback,
              save_primal=save_primal)
    if decay:
        kw.update(gamma_decay_type="step", gamma_decay_params={"decay_steps": 7, "decay_factor": 0.7})
    lam0 = torch.zeros(96, device=DEV)

    monkeypatch.setenv("DUALIP_GRAPH", "0")
    ref = AcceleratedGradientDescent(**kw).maximize(_objective(p, _mixed_map(n), 2e-2), lam0)

    monkeypatch.setenv("DUALIP_GRAPH", "1")
    monkeypatch.setenv("DUALIP_GRAPH_CHUNK", "8")
    solver = AcceleratedGradientDescent(**kw)
    obj = _objective(p, _mixed_map(n), 2e-2)
    loop = FusedAscentLoop(solver, obj, lam0)
    try:
        loop.run(1, solver.max_iter)
        out = loop.finish()
        assert loop.graph_launches == 6, "six whole chunks of 8 fit into 53 iterations"
    finally:
        loop.close()
    assert out.dual_objective_log == ref.dual_objective_log
    assert out.step_size_log == ref.step_size_log
    assert torch.equal(out.dual_val, ref.dual_val)
    assert torch.equal(out.objective_result.dual_gradient, ref.objective_result.dual_gradient)
    assert solver.gamma == pytest.approx(2e-2 * (0.7 ** (53 // 7) if decay else 1.0), rel=1e-12)
    if save_primal:
        assert torch.equal(out.objective_result.primal_var, ref.objective_result.primal_var)


def test_scheduled_step_through_the_raw_abi():
    """dualip_matching_ascent_step_scheduled called directly (no Python loop object): three iterations equal three
    dualip_matching_ascent_step calls with explicit arguments."""
    lib = _native.lib()
    p = random_problem(5, 3000, 64, 7.0)
    n, m = p["ccol"].size - 1, 64
    gamma, betas = 3e-2, [0.0, 0.28, 0.43]

    def run(scheduled):
        obj = _objective(p, create_projection_map("simplex", {"z": 1.0}, n), gamma)
        h = ctypes.c_void_p()
        _native.check(lib.dualip_agd_create(ctypes.byref(h), m, 0, None, None, 1e-3, 0.1, 15))
        _native.check(lib.dualip_agd_reserve_log(h, 3))
        grad = torch.empty(m, device=DEV)
        scal = torch.zeros(8, dtype=torch.float64, device=DEV)
        st = torch.cuda.current_stream().cuda_stream
        if scheduled:
            _native.check(lib.dualip_agd_set_schedule(h, 3, (ctypes.c_double * 3)(gamma, gamma, gamma), (ctypes.c_float * 3)(*betas),
                                                      None, 1.0))
            for _ in range(3):
                _native.check(lib.dualip_matching_ascent_step_scheduled(obj._plan, h, None, obj.b_vec.data_ptr(), grad.data_ptr(),
                                                                        scal.data_ptr(), st))
            assert lib.dualip_agd_steps_launched(h) == 3
            with pytest.raises(ValueError):  # the schedule is exhausted
                _native.check(lib.dualip_matching_ascent_step_scheduled(obj._plan, h, None, obj.b_vec.data_ptr(), grad.data_ptr(),
                                                                        scal.data_ptr(), st))
        else:
            for i in range(3):
                _native.check(lib.dualip_matching_ascent_step(obj._plan, h, obj.b_vec.data_ptr(), gamma, grad.data_ptr(),
                                                              scal.data_ptr(), None, betas[i], 0, 1.0, i, st))
        y = torch.empty(m, device=DEV)
        _native.check(lib.dualip_agd_get(h, None, y.data_ptr(), st))
        log = (ctypes.c_double * 3)()
        steps = (ctypes.c_double * 3)()
        _native.check(lib.dualip_agd_read_log(h, 3, log, steps, st))
        torch.cuda.synchronize()
        lib.dualip_agd_destroy(h)
        return y.cpu(), list(log), grad.cpu()

    y0, log0, g0 = run(False)
    y1, log1, g1 = run(True)
    assert torch.equal(y0, y1) and log0 == log1 and torch.equal(g0, g1)


def _heavy_tailed_rows(seed=3, n=40000, m=200):
    """Rows whose coefficients differ by orders of magnitude (no Jacobi scaling): one fixed-point quantum for all rows cannot
    resolve the small ones."""
    p = random_problem(seed, n, m, 9.0)
    rng = np.random.default_rng(seed + 1)
    row_scale = np.exp2(rng.integers(-18, 6, m)).astype(np.float32)
    p["a"] = (p["a"] * row_scale[p["row"]]).astype(np.float32)
    p["b"] = (p["b"] * row_scale).astype(np.float32)
    return p


def test_row_equilibration_keeps_x_and_restores_the_fixed_point_accumulator(monkeypatch):
    """Power-of-two row scaling at plan time: the plan takes the deterministic fixed-point accumulator where a single quantum
    lacked resolution, the primal is bit-identical to the unscaled plan and to the oracle (powers of two commute with fp32
    rounding), and the gradient matches the fp64-accumulated oracle row by row."""
    p = _heavy_tailed_rows()
    n, m, gamma = p["ccol"].size - 1, int(p["n_rows"]), 2e-2
    lam = torch.from_numpy(p["lam"]).to(DEV)

    monkeypatch.setenv("DUALIP_ROW_SCALE", "0")
    plain = _objective(p, _mixed_map(n), gamma)
    info0 = plain.plan_info()
    r0 = plain.calculate(lam, save_primal=True)
    monkeypatch.delenv("DUALIP_ROW_SCALE")
    scaled = _objective(p, _mixed_map(n), gamma)
    info1 = scaled.plan_info()
    r1 = scaled.calculate(lam, save_primal=True)
    assert info0["fixed_point"] == 0 and info0["row_scaled"] == 0, "the fixture must defeat the single-quantum accumulator"
    assert info1["fixed_point"] == 1 and info1["row_scaled"] == 1
    assert torch.equal(r0.primal_var, r1.primal_var)

    even, odd = np.arange(0, n, 2), np.arange(1, n, 2)
    opm = {"s": O.ProjEntry("simplex", {"z": 1.0}, even), "b": O.ProjEntry("box", {"lower": 0.0, "upper": 1.0}, odd)}
    ref = O.matching_calculate(p["ccol"], p["row"], p["a"], p["c"], m, opm, p["lam"], gamma, p["b"])
    assert np.array_equal(r1.primal_var.cpu().numpy(), ref.primal_var)
    # exact row sums of the (bit-identical) primal in fp64: the fixed-point result must sit within a few fp32 ulps of them,
    # small rows included
    x64 = ref.primal_var.astype(np.float64)
    exact = np.zeros(m)
    np.add.at(exact, p["row"], (p["a"] * ref.primal_var).astype(np.float64))
    absum = np.zeros(m)
    np.add.at(absum, p["row"], np.abs(p["a"].astype(np.float64) * x64))
    g1 = r1.dual_gradient.cpu().numpy().astype(np.float64) + p["b"].astype(np.float64)
    b64 = np.abs(p["b"].astype(np.float64))
    assert np.all(np.abs(g1 - exact) <= 4e-7 * absum + 1.3e-7 * (np.abs(exact) + 2 * b64) + 1e-30)  # + the fp32 `sum - b` itself
    # and it is reproducible bit for bit (integer adds commute); the fp32-atomic plan is only close
    r2 = scaled.calculate(lam)
    assert torch.equal(r1.dual_gradient, r2.dual_gradient)
    assert torch.allclose(r0.dual_gradient, r1.dual_gradient, rtol=2e-5, atol=1e-6)
    assert abs(float(r1.scalars64[0]) - ref.dual_objective) <= 1e-5 * abs(ref.dual_objective)


@pytest.mark.parametrize("decay", [False, True])
def test_host_buffer_loop_one_native_call_per_iteration(monkeypatch, decay):
    """maximize() with a host-resident dual: dualip_matching_step_host (lambda up, fused kernel, gradient + scalars down, host
    step: one call) against the loop that goes through calculate() and dualip_agd_host_step separately.  Same kernels, same
    host arithmetic: identical iterates and step sizes; callbacks see every iteration's result as fresh tensors."""
    monkeypatch.setenv("DUALIP_REBALANCE", "0")
    p = random_problem(13, 5000, 80, 8.0)
    n = p["ccol"].size - 1
    kw = dict(max_iter=30, gamma=2e-2, initial_step_size=1e-3, max_step_size=0.1)
    if decay:
        kw.update(gamma_decay_type="step", gamma_decay_params={"decay_steps": 7, "decay_factor": 0.7})
    lam0 = torch.zeros(80)
    seen = {"0": [], "1": []}
    outs = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("DUALIP_HOST_FUSED", fused)
        cb = (lambda i, r, key=fused: seen[key].append((i, float(r.dual_objective), r.dual_gradient)))
        outs[fused] = AcceleratedGradientDescent(iteration_callback=cb, **kw).maximize(_objective(p, _mixed_map(n), 2e-2), lam0)
    a, b = outs["0"], outs["1"]
    assert torch.equal(a.dual_val, b.dual_val) and a.step_size_log == b.step_size_log
    assert np.allclose(a.dual_objective_log, b.dual_objective_log, rtol=1e-7)  # float32-rounded vs double log entries
    assert len(seen["1"]) == 30 and [i for i, _, _ in seen["1"]] == list(range(1, 31))
    for (_, o0, g0), (_, o1, g1) in zip(seen["0"], seen["1"]):
        assert abs(o0 - o1) <= 1e-6 * abs(o0) and torch.equal(g0, g1)
    assert len({g.data_ptr() for _, _, g in seen["1"]}) == 30, "callback results must not alias the loop's buffers"
    assert b.dual_val.device.type == "cpu" and b.objective_result.dual_gradient.device.type == "cpu"
    quiet = AcceleratedGradientDescent(iteration_callback=no_iteration_callback, **kw).maximize(_objective(p, _mixed_map(n), 2e-2), lam0)
    assert torch.equal(quiet.dual_val, b.dual_val) and quiet.objective_result is not None


@pytest.mark.parametrize("decay", [False, True])
def test_all_cta_tail_equals_the_last_cta_tail(monkeypatch, decay):
    """grid_tail.cuh: every CTA takes a slice of the m-length tail and of the accelerated step, with grid-wide barriers in
    between (default for m >= 16384) against the tail run by the last CTA alone.  Element-wise arithmetic is identical; only the
    order of the double-precision partial sums differs, so the logs agree to rounding and the iterates to a few ulps.  Also
    through CUDA-graph replay (barrier state persists across the captured launches)."""
    monkeypatch.setenv("DUALIP_REBALANCE", "0")
    p = random_problem(19, 9000, 700, 8.0)
    n = p["ccol"].size - 1
    kw = dict(max_iter=45, gamma=2e-2, initial_step_size=1e-3, max_step_size=0.1, iteration_callback=no_iteration_callback)
    if decay:
        kw.update(gamma_decay_type="step", gamma_decay_params={"decay_steps": 7, "decay_factor": 0.7})
    lam0 = torch.zeros(700, device=DEV)
    outs = {}
    for tag, env, graph in (("last", "0", "0"), ("grid", "1", "0"), ("grid_graph", "1", "1")):
        monkeypatch.setenv("DUALIP_GRID_TAIL", env)
        monkeypatch.setenv("DUALIP_GRAPH", graph)
        monkeypatch.setenv("DUALIP_GRAPH_CHUNK", "8")
        obj = _objective(p, _mixed_map(n), 2e-2)
        assert obj.plan_info()["n_ctas"] > 1 and obj.plan_info()["grid_tail"] == int(env)
        outs[tag] = AcceleratedGradientDescent(**kw).maximize(obj, lam0)
        assert obj.plan_info()["grid_barrier_status"] == 0  # (2 after a grid-wide barrier timed out: maximize() would raise)
        obj.check_grid_barrier()
    a, b, c = outs["last"], outs["grid"], outs["grid_graph"]
    assert np.allclose(a.dual_objective_log, b.dual_objective_log, rtol=1e-9)
    assert np.allclose(a.step_size_log, b.step_size_log, rtol=1e-6)
    assert torch.allclose(a.dual_val, b.dual_val, rtol=1e-5, atol=1e-7)
    assert b.dual_objective_log == c.dual_objective_log and torch.equal(b.dual_val, c.dual_val), "graph replay of the all-CTA tail"
    assert torch.allclose(a.objective_result.dual_gradient, b.objective_result.dual_gradient, rtol=1e-5, atol=1e-6)
    # evaluation only (dualip_matching_calc, device and host buffers): the all-CTA tail without the optimizer's part
    lam = a.dual_val
    res = {}
    for env in ("0", "1"):
        monkeypatch.setenv("DUALIP_GRID_TAIL", env)
        obj = _objective(p, _mixed_map(n), 2e-2)
        res[env] = (obj.calculate(lam), obj.calculate(lam.cpu()))
    for k in (0, 1):
        assert torch.equal(res["0"][k].dual_gradient, res["1"][k].dual_gradient)
        assert abs(float(res["0"][k].scalars64[0]) - float(res["1"][k].scalars64[0])) <= 1e-12 * abs(float(res["0"][k].scalars64[0]))
