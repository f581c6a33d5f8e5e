"""-m gpu: the CUDA path (through the C-ABI) against the reference-generated fixtures and the oracles.

Bars (BASELINE.json north_star): primal x and dual objective within 1e-5 relative; projection branch
(feasible / top-2 shortcut / Duchi) and support size rho bit-exact.  On the committed fixtures x is in fact bit-identical
to the reference's, which the tests assert where it holds by construction (same fp32 operation order)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases, load_case, random_problem
from dualip_b200 import _native
from dualip_b200.objectives.matching import (
    MatchingInputArgs,
    MatchingSolverDualObjectiveFunction,
    MatchingSolverDualObjectiveFunctionDistributed,
)
from dualip_b200.optimizers.agd import AcceleratedGradientDescent
from dualip_b200.preprocessing.precondition import jacobi_invert_precondition, jacobi_precondition
from dualip_b200.projections import create_projection_map, project
from dualip_b200.run_solver import run_solver
from dualip_b200.types import ComputeArgs, ObjectiveArgs, SolverArgs
from oracle import c_oracle
from oracle import dualip_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _csc(p, dev=DEV, index_dtype=torch.int64):
    m, n = int(p["n_rows"]), p["ccol"].size - 1
    ccol = torch.from_numpy(np.asarray(p["ccol"])).to(index_dtype)
    row = torch.from_numpy(np.asarray(p["row"])).to(index_dtype)
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(np.asarray(p["a"])), size=(m, n)).to(dev)
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(np.asarray(p["c"])), size=(m, n)).to(dev)
    return A, C


def _diag_per_column(diag, ccol):
    first = ccol[:-1][np.diff(ccol) > 0]
    d = diag[first]
    return d & 3, d >> 2, np.diff(ccol) > 0


@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("index_dtype", [torch.int64, torch.int32])
def test_fixtures_generated_by_the_reference(name, index_dtype):
    d, ptype, params = load_case(name)
    n, m = d["ccol"].size - 1, int(d["n_rows"])
    A, C = _csc(d, index_dtype=index_dtype)
    obj = MatchingSolverDualObjectiveFunction(
        MatchingInputArgs(A, C, create_projection_map(ptype, params, n), torch.from_numpy(d["b"]).to(DEV)), gamma=float(d["gamma"]))
    r = obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True, diagnostics=True)
    x = r.primal_var.cpu().numpy()
    assert np.array_equal(x, d["x_b1"]), "primal x differs from the reference"
    scal = d["scal_b1"]  # dual_obj, reg, primal_obj, lam.grad, max_pos_slack, sum_pos_slack
    got = r.scalars64.cpu().numpy()
    assert abs(got[0] - scal[0]) <= 1e-5 * abs(scal[0])
    assert abs(got[2] - scal[1]) <= 1e-5 * abs(scal[1]) + 1e-9
    assert abs(got[1] - scal[2]) <= 1e-5 * abs(scal[2])
    assert abs(got[3] - scal[3]) <= 1e-5 * abs(scal[3]) + 1e-4
    assert abs(got[4] - scal[4]) <= 1e-5 * max(1.0, abs(scal[4]))
    assert abs(got[5] - scal[5]) <= 1e-5 * max(1.0, abs(scal[5]))
    g = r.dual_gradient.cpu().numpy()
    assert np.allclose(g, d["grad_b1"], rtol=1e-5, atol=1e-4 * max(1.0, np.abs(d["grad_b1"]).max() * 1e-2))
    assert r.dual_objective.dtype == torch.float32 and r.dual_gradient.dtype == torch.float32
    if ptype.startswith("simplex"):
        orc = O.matching_calculate(d["ccol"], d["row"], d["a"], d["c"], m, {"k": O.ProjEntry(ptype, params, np.arange(n))},
                                   d["lam"], float(d["gamma"]), d["b"])
        br, rho, nonempty = _diag_per_column(r.projection_diag.cpu().numpy(), d["ccol"])
        assert np.array_equal(br, orc.branch[nonempty]), "projection branch selection differs"
        sel = orc.branch[nonempty] > 0
        assert np.array_equal(rho[sel], np.minimum(orc.rho[nonempty][sel], 63)), "support size differs"


@pytest.mark.parametrize("batching", [True, False])
@pytest.mark.parametrize("index_dtype", [torch.int64, torch.int32])
def test_mixed_map_fixture_derived_from_the_reference(batching, index_dtype):
    """configs[2]: simplex(z=1) on even entities, box[0,1] on odd ones, Jacobi-scaled C3-shaped data from the reference's
    generator, at a late dual.  Expected values = the reference's own `calculate` per entry on that entry's column
    sub-matrix, combined (tests/golden/make_golden_r2.py) -- the reference cannot evaluate a two-entry map itself
    (utils/sparse_utils.py:177,220).  x must be bit-identical."""
    d = np.load(f"{GOLDEN}/mixed_c3shape.npz")
    n, m, tag = d["ccol"].size - 1, int(d["n_rows"]), "b1" if batching else "b0"
    A, C = _csc(d, index_dtype=index_dtype)
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0}, n, indices=list(range(0, n, 2))))
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, n, indices=list(range(1, n, 2))))
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(d["b"]).to(DEV)),
                                              gamma=float(d["gamma"]), batching=batching)
    r = obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True, diagnostics=True)
    assert np.array_equal(r.primal_var.cpu().numpy(), d[f"x_{tag}"]), "primal x differs from the reference"
    scal, got = d[f"scal_{tag}"], r.scalars64.cpu().numpy()
    assert abs(got[0] - scal[0]) <= 1e-5 * abs(scal[0])
    assert abs(got[2] - scal[1]) <= 1e-5 * abs(scal[1])
    assert abs(got[1] - scal[2]) <= 1e-5 * abs(scal[2])
    assert np.allclose(r.dual_gradient.cpu().numpy(), d[f"grad_{tag}"], rtol=1e-5, atol=1e-4)
    br, _, nonempty = _diag_per_column(r.projection_diag.cpu().numpy(), d["ccol"])
    even_nonempty = (np.arange(n) % 2 == 0)[nonempty]
    assert np.mean(br[even_nonempty] == 2) > 0.05, "the fixture must exercise the sorted scan (late iterate)"


@pytest.mark.parametrize("batching", [True, False])
def test_simplex_eq_has_the_reference_padded_block_semantics(batching):
    """`simplex_eq` through the objective depends on the padded length of the reference's length buckets
    (simplex.py:160-161, sparse_utils.py:197-211): with and without batching the reference returns
    different x for the same column.  Both must be reproduced bit for bit (fixture: tests/golden/make_golden_r2.py)."""
    d, ptype, params = load_case("simplex_eq")
    n = d["ccol"].size - 1
    tag = "b1" if batching else "b0"
    assert not np.array_equal(d["x_b1"], d["x_b0"])
    A, C = _csc(d)
    obj = MatchingSolverDualObjectiveFunction(
        MatchingInputArgs(A, C, create_projection_map(ptype, params, n), torch.from_numpy(d["b"]).to(DEV)),
        gamma=float(d["gamma"]), batching=batching)
    r = obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True)
    assert np.array_equal(r.primal_var.cpu().numpy(), d[f"x_{tag}"])
    assert abs(r.scalars64.cpu().numpy()[0] - d[f"scal_{tag}"][0]) <= 1e-5 * abs(d[f"scal_{tag}"][0])
    assert np.allclose(r.dual_gradient.cpu().numpy(), d[f"grad_{tag}"], rtol=1e-5, atol=1e-4)
    # and on a map with an identity class in front (class ids shift by one), against the per-column C oracle
    pm = create_projection_map(ptype, params, n, indices=list(range(n - 5)))
    obj2 = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(d["b"]).to(DEV)),
                                               gamma=float(d["gamma"]), batching=batching)
    r2 = obj2.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True)
    pad = O.pad_table(d["ccol"], int(d["n_rows"]), [O.ProjEntry(ptype, params, np.arange(n - 5))], batching)
    rc = c_oracle.calculate(d["ccol"], d["row"], d["a"], d["c"], int(d["n_rows"]),
                            [c_oracle.make_class("identity", {}), c_oracle.make_class(ptype, params)], d["lam"], float(d["gamma"]),
                            d["b"], col_class=(np.arange(n) < n - 5).astype(np.uint8), pad_len=np.vstack([np.zeros(32, np.int32), pad[0]]))
    assert np.array_equal(r2.primal_var.cpu().numpy(), rc["x"])


def test_reference_known_answer_through_fused_maximizer():
    """Reference tests/objectives/test_dualip_matching_simplex.py:102-141 on the GPU path."""
    from test_oracle_golden import _scala_5x5

    ccol, row, a, c, b = _scala_5x5()
    A, C = _csc(dict(ccol=ccol, row=row, a=a, c=c, n_rows=5))
    obj = MatchingSolverDualObjectiveFunction(
        MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1}, 5), torch.from_numpy(b).to(DEV), None), gamma=1e-3)
    solver = AcceleratedGradientDescent(max_iter=30, gamma=1e-3, iteration_callback=lambda i, r: None)
    res = solver.maximize(obj, 0.1 * torch.ones(5, device=DEV))
    for i, true_val in [(2, -3.6010155991401818), (16, -3.60842718733725), (23, -3.5080258013053136), (29, -3.4868496294227143)]:
        assert abs(res.dual_objective_log[i - 1] - true_val) < 1e-5
    assert res.dual_val.shape == (5,) and len(res.step_size_log) == 30


def test_maximizer_trace_against_reference_run():
    d = np.load(f"{GOLDEN}/agd_trace.npz")
    n, m = d["ccol"].size - 1, int(d["n_rows"])
    A, C = _csc(d)
    for name, kw in (("plain", {}), ("decay", dict(gamma_decay_type="step", gamma_decay_params={"decay_steps": 8, "decay_factor": 0.5}))):
        obj = MatchingSolverDualObjectiveFunction(
            MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), torch.from_numpy(d["b"]).to(DEV)), gamma=1e-2)
        solver = AcceleratedGradientDescent(max_iter=40, gamma=1e-2, initial_step_size=1e-3, max_step_size=0.1,
                                            iteration_callback=lambda i, r: None, **kw)
        res = solver.maximize(obj, torch.zeros(m, device=DEV))
        assert np.allclose(res.dual_objective_log, d[f"{name}_obj_log"], rtol=1e-5)
        # the Lipschitz step is a ratio of small differences of successive gradients, so the fp32 atomics' summation order
        # shows up at the 1e-2 level late in the run (the objective log above is the parity statement)
        assert np.allclose(res.step_size_log, d[f"{name}_step_log"], rtol=3e-2)
        assert np.allclose(res.dual_val.cpu().numpy(), d[f"{name}_dual"], rtol=1e-3, atol=1e-3)
        if name == "decay":
            assert abs(solver.gamma - 1e-2 * 0.5**5) < 1e-12  # host-side gamma schedule (agd.py:102-109)


MIXED_CASES = [
    dict(seed=1, n_cols=4000, n_rows=64, mean_deg=9.0, scale_c=1.0, lam_scale=0.05, gamma=1e-3),
    dict(seed=2, n_cols=3000, n_rows=48, mean_deg=10.0, scale_c=20.0, lam_scale=2.0, gamma=1e-1),
    dict(seed=3, n_cols=2500, n_rows=300, mean_deg=20.0, scale_c=10.0, lam_scale=1.0, gamma=5e-2, max_deg=60,
         long_cols=[(3, 250), (99, 300), (1000, 33)]),
    dict(seed=4, n_cols=5000, n_rows=40, mean_deg=1.2, scale_c=30.0, lam_scale=1.0, gamma=1e-1),  # mostly 1-entry columns
    dict(seed=5, n_cols=700, n_rows=2000, mean_deg=150.0, scale_c=10.0, lam_scale=1.0, gamma=5e-2, max_deg=1500,
         long_cols=[(7, 1400), (8, 1100)]),  # MovieLens-shaped: long columns, warp-per-column kernel above 1024
]


@pytest.mark.parametrize("case", MIXED_CASES, ids=lambda c: f"seed{c['seed']}")
def test_mixed_projection_map_against_oracles(case):
    case = dict(case)
    gamma = case.pop("gamma")
    p = random_problem(**case)
    n, m = p["n_cols"], p["n_rows"]
    idx = np.arange(n)
    groups = [("simplex", {"z": 1.0}, idx[idx % 5 == 0]), ("box", {"lower": 0.0, "upper": 1.0}, idx[idx % 5 == 1]),
              ("simplex", {"z": 2.5}, idx[idx % 5 == 2]), ("cone", {"lower": 0.0}, idx[idx % 5 == 3])]  # idx % 5 == 4: none
    pm, opm = {}, {}
    classes = [c_oracle.make_class("identity", {})]
    col_class = np.zeros(n, dtype=np.uint8)
    deg = np.diff(p["ccol"])
    for k, (ptype, params, cols) in enumerate(groups):
        pm.update(create_projection_map(ptype, params, n, indices=cols.tolist(), key_prefix=f"g{k}_"))
        opm[f"g{k}"] = O.ProjEntry(ptype, params, cols)
        unpadded = ptype == "simplex" and bool((deg[cols] == 1).any() and not (deg[cols] == 2).any())
        classes.append(c_oracle.make_class(ptype, params, d1_unpadded=unpadded))
        col_class[cols] = k + 1
    A, C = _csc(p)
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(p["b"]).to(DEV)), gamma=gamma)
    info = obj.plan_info()
    assert info["n_slab_cols"] + info["n_mid_cols"] + info["n_long_cols"] == int((deg > 0).sum())
    r = obj.calculate(torch.from_numpy(p["lam"]).to(DEV), save_primal=True, diagnostics=True)
    x = r.primal_var.cpu().numpy()
    ref_c = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, classes, p["lam"], gamma, p["b"], col_class)
    rel = np.abs(x - ref_c["x"]) / np.maximum(np.abs(ref_c["x"]), 1e-6)
    assert rel.max() <= 1e-5, f"primal x: max relative difference {rel.max()}"
    assert int(((x != 0) != (ref_c["x"] != 0)).sum()) == 0, "support of x differs"
    got = r.scalars64.cpu().numpy()
    assert abs(got[0] - ref_c["scal"][0]) <= 1e-5 * abs(ref_c["scal"][0])
    assert np.allclose(r.dual_gradient.cpu().numpy(), ref_c["grad"], rtol=1e-5, atol=1e-5 * np.abs(ref_c["grad"]).max())
    # branch / support size, bit-exact against the C port (per column)
    br, rho, nonempty = _diag_per_column(r.projection_diag.cpu().numpy(), p["ccol"])
    cd = ref_c["diag"][nonempty]
    is_sx = cd != 255
    assert np.array_equal(br[is_sx], cd[is_sx] & 3)
    dsel = is_sx & ((cd & 3) > 0)
    assert np.array_equal(rho[dsel], cd[dsel] >> 2)
    if max(deg) <= 400:  # the numpy oracle (padded blocks) as a second opinion where it is fast enough
        ref_np = O.matching_calculate(p["ccol"], p["row"], p["a"], p["c"], m, opm, p["lam"], gamma, p["b"])
        assert np.abs(x - ref_np.primal_var).max() <= 1e-5 * max(1.0, np.abs(ref_np.primal_var).max())


def test_empty_and_degenerate_inputs():
    m = 7
    # no columns at all / all columns empty / a single entry
    for n, ccol in ((0, [0]), (5, [0] * 6)):
        A = torch.sparse_csc_tensor(torch.tensor(ccol), torch.zeros(0, dtype=torch.int64), torch.zeros(0), size=(m, n)).to(DEV)
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, A, create_projection_map("simplex", {"z": 1.0}, n), torch.ones(m, device=DEV)), gamma=1e-2)
        r = obj.calculate(torch.ones(m, device=DEV), save_primal=True)
        assert torch.equal(r.dual_gradient, -torch.ones(m, device=DEV)) and r.primal_var.numel() == 0
        assert abs(float(r.dual_objective) + m) < 1e-6
    A = torch.sparse_csc_tensor(torch.tensor([0, 1]), torch.tensor([3]), torch.tensor([2.0]), size=(m, 1)).to(DEV)
    C = torch.sparse_csc_tensor(torch.tensor([0, 1]), torch.tensor([3]), torch.tensor([-5.0]), size=(m, 1)).to(DEV)
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, 1), torch.zeros(m, device=DEV)), gamma=1.0)
    r = obj.calculate(torch.zeros(m, device=DEV), save_primal=True)  # v = 5 -> projected to z = 1
    assert float(r.primal_var[0]) == 1.0 and float(r.dual_gradient[3]) == 2.0
    with pytest.raises(ValueError):  # row index out of range
        bad = torch.sparse_csc_tensor(torch.tensor([0, 1]), torch.tensor([9]), torch.tensor([2.0]), size=(m, 1)).to(DEV)
        MatchingSolverDualObjectiveFunction(MatchingInputArgs(bad, bad, create_projection_map("box", {}, 1), torch.zeros(m, device=DEV)), gamma=1.0)
    with pytest.raises(ValueError):  # unknown projection name, like the reference (projections/base.py:55-56)
        MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("l2ball", {}, 1), torch.zeros(m, device=DEV)), gamma=1.0)


def test_projection_operators_on_padded_blocks():
    """ProjectionOperator.__call__ (dualip_project_block) against reference outputs on zero-padded [L x K] blocks."""
    pv = np.load(f"{GOLDEN}/projection_vectors.npz")
    table = [("simplex", {"z": 1.0}, "simplex_z1.0"), ("simplex", {"z": 2.5}, "simplex_z2.5"), ("simplex_eq", {"z": 1.0}, "simplex_eq_z1.0"),
             ("box", {"lower": 0.0, "upper": 1.0}, "box_lower0.0_upper1.0"), ("box", {"lower": -0.5, "upper": 0.25}, "box_lower-0.5_upper0.25"),
             ("cone", {"lower": 0.0}, "cone_lower0.0"), ("cone", {"upper": 0.1}, "cone_upper0.1"), ("cone", {}, "cone")]
    for L in (1, 2, 7, 16, 33, 150):
        x = torch.from_numpy(pv[f"L{L}_x"]).to(DEV)
        for name, params, tag in table:
            out = project(name, **params)(x)
            assert np.array_equal(out.cpu().numpy(), pv[f"L{L}_{tag}"]), (L, tag)
            assert torch.equal(x, torch.from_numpy(pv[f"L{L}_x"]).to(DEV))  # input untouched
    v = torch.tensor([-0.5, 0.2, -1.0, 0.3], device=DEV)  # reference tests/projections/test_simplex.py:270-284
    assert torch.equal(project("simplex", z=1.0)(v).squeeze(1), torch.tensor([0.0, 0.2, 0.0, 0.3], device=DEV))


def test_sharded_path_single_process():
    """partial -> (no-op all-reduce) -> epilogue reproduces the single-device result; shards sum up."""
    p = random_problem(21, 6000, 80, 8.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma = p["n_cols"], p["n_rows"], 2e-2
    A, C = _csc(p)
    b = torch.from_numpy(p["b"]).to(DEV)
    lam = torch.from_numpy(p["lam"]).to(DEV)
    full = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), b), gamma).calculate(lam)
    local = MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), None)
    dist_obj = MatchingSolverDualObjectiveFunctionDistributed(local, b, gamma, host_device=DEV)
    r = dist_obj.calculate(lam)
    assert torch.allclose(r.dual_gradient, full.dual_gradient, rtol=1e-6, atol=1e-5)
    assert abs(float(r.scalars64[0]) - float(full.scalars64[0])) <= 1e-6 * abs(float(full.scalars64[0]))
    with pytest.raises(NotImplementedError):
        dist_obj.calculate(lam, save_primal=True)
    # two shards, partials summed by hand (what the all-reduce does)
    from dualip_b200.utils.dist_utils import global_to_local_projection_map, split_tensors_to_devices

    a_s, c_s, index_map = split_tensors_to_devices(A, C, [DEV, DEV])
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    total = torch.zeros(m + 2, device=DEV)
    for k in range(2):
        o = MatchingSolverDualObjectiveFunction(MatchingInputArgs(a_s[k], c_s[k], global_to_local_projection_map(pm, index_map[k]), None), gamma)
        part = torch.empty(m + 2, device=DEV)
        o.launch_partial(lam.data_ptr(), gamma, part.data_ptr())
        total += part
        loc = o.calculate(lam)  # local-shard mode of the reference (matching.py:179-184)
        assert loc.dual_val_times_grad is None and torch.allclose(loc.dual_gradient, part[:m])
    assert torch.allclose(total[:m] - b, full.dual_gradient, rtol=1e-5, atol=1e-5)


def test_fused_sharded_step_equals_epilogue_plus_step():
    """Sharded loop: the single-launch objective tail + optimizer update (dualip_agd_step_sharded, taken without an
    iteration callback) gives the same ascent as epilogue kernel + plain step, and as the single-device loop."""
    from dualip_b200.optimizers.agd import no_iteration_callback

    p = random_problem(22, 5000, 96, 8.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma = p["n_cols"], p["n_rows"], 2e-2
    A, C = _csc(p)
    b = torch.from_numpy(p["b"]).to(DEV)
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    logs = {}
    for tag, cb in (("fused", no_iteration_callback), ("split", lambda i, r: None)):
        obj = MatchingSolverDualObjectiveFunctionDistributed(MatchingInputArgs(A, C, pm, None), b, gamma, host_device=DEV)
        solver = AcceleratedGradientDescent(max_iter=40, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                            gamma_decay_type="step", gamma_decay_params={"decay_steps": 9, "decay_factor": 0.5},
                                            iteration_callback=cb)
        logs[tag] = solver.maximize(obj, torch.zeros(m, device=DEV))
    single = AcceleratedGradientDescent(max_iter=40, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                        gamma_decay_type="step", gamma_decay_params={"decay_steps": 9, "decay_factor": 0.5},
                                        iteration_callback=no_iteration_callback).maximize(
        MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, b), gamma), torch.zeros(m, device=DEV))
    for other in (logs["split"], single):
        # the packed partial vector carries c.x and ||x||^2 as float32 (one all-reduce of m+2 floats): 6e-8 relative
        assert np.allclose(logs["fused"].dual_objective_log, other.dual_objective_log, rtol=1e-6, atol=0)
        assert np.allclose(logs["fused"].step_size_log, other.step_size_log, rtol=1e-6)
        assert torch.allclose(logs["fused"].dual_val, other.dual_val, rtol=1e-6, atol=1e-7)
    r = logs["fused"].objective_result
    assert torch.isfinite(r.dual_gradient).all() and float(r.max_pos_slack) >= 0.0


def test_host_buffer_path_and_run_solver(tmp_path):
    p = random_problem(31, 3000, 50, 8.0, scale_c=10.0)
    n, m, gamma = p["n_cols"], p["n_rows"], 2e-2
    A, C = _csc(p, dev="cpu")
    args = MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), torch.from_numpy(p["b"]))
    res = run_solver(args, SolverArgs(max_iter=20, gamma=gamma, initial_step_size=1e-3, save_primal=True), ComputeArgs(DEV),
                     ObjectiveArgs("matching"))
    assert res.dual_val.device.type == "cuda" and res.objective_result.primal_var.numel() == p["row"].size
    opm = {"k": O.ProjEntry("simplex", {"z": 1.0}, np.arange(n))}

    def calc(lam, g):
        r = O.matching_calculate(p["ccol"], p["row"], p["a"], p["c"], m, opm, lam, g, p["b"])
        return r.dual_gradient, np.float32(r.dual_objective)

    y, obj_log, _, _ = O.agd_maximize(calc, np.zeros(m, np.float32), 20, gamma, 1e-3, 0.1)
    assert np.allclose(res.dual_objective_log, obj_log, rtol=1e-5)
    # warm start from a saved dual (SolverArgs.initial_dual_path, run_solver.py:127-132)
    path = str(tmp_path / "dual.pt")
    torch.save(res.dual_val.cpu(), path)
    res2 = run_solver(args, SolverArgs(max_iter=1, gamma=gamma, initial_step_size=1e-3, initial_dual_path=path), ComputeArgs(DEV),
                      ObjectiveArgs("matching"))
    g1, o1 = calc(res.dual_val.cpu().numpy(), gamma)
    assert abs(res2.dual_objective_log[0] - float(o1)) <= 1e-5 * abs(float(o1))
    # host buffers: dual on the CPU, results on the CPU, same numbers
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A.to(DEV), C.to(DEV), args.projection_map, args.b_vec.to(DEV)), gamma)
    lam = torch.from_numpy(p["lam"])
    rh = obj.calculate(lam, gamma=gamma)
    rd = obj.calculate(lam.to(DEV), gamma=gamma)
    assert rh.dual_gradient.device.type == "cpu" and torch.allclose(rh.dual_gradient, rd.dual_gradient.cpu(), rtol=1e-6, atol=1e-6)
    assert abs(float(rh.dual_objective) - float(rd.dual_objective)) <= 1e-6 * abs(float(rd.dual_objective))


def test_jacobi_precondition_matches_oracle():
    p = random_problem(41, 2000, 30, 6.0)
    A, _ = _csc(p)
    b = torch.from_numpy(p["b"]).to(DEV)
    a_ref, b_ref, norms_ref = O.jacobi_precondition(p["a"], p["row"], p["b"], p["n_rows"])
    norms = jacobi_precondition(A, b)
    assert np.allclose(norms.cpu().numpy(), norms_ref, rtol=1e-6)
    assert np.allclose(A.values().cpu().numpy(), a_ref, rtol=1e-6) and np.allclose(b.cpu().numpy(), b_ref, rtol=1e-6)
    lam = torch.rand(p["n_rows"], device=DEV)
    assert torch.allclose(jacobi_invert_precondition(lam, norms), lam / norms)


def test_c_abi_direct_call_without_python_wrappers():
    """The boundary itself: raw pointers into dualip_plan_create / dualip_matching_calc, as a non-Python host would."""
    p = random_problem(51, 1500, 33, 7.0, scale_c=10.0)
    n, m, gamma = p["n_cols"], p["n_rows"], 3e-2
    lib = _native.lib()
    t = {k: torch.from_numpy(np.ascontiguousarray(p[k])).to(DEV) for k in ("ccol", "row", "a", "c", "b", "lam")}
    cls = (_native.ProjClass * 1)(_native.ProjClass(_native.PROJ_SIMPLEX, 0.0, 0.0, 1.0, float(np.float32(1.0 + 1e-6)), 0))
    desc = _native.CscDesc(n_cols=n, nnz=p["row"].size, n_rows=m, index_bits=64, ccol_dev=t["ccol"].data_ptr(), row_dev=t["row"].data_ptr(),
                           a_dev=t["a"].data_ptr(), c_dev=t["c"].data_ptr(), col_class_dev=None,
                           classes=ctypes.cast(cls, ctypes.POINTER(_native.ProjClass)), n_classes=1, device=0)
    plan = ctypes.c_void_p()
    assert lib.dualip_plan_create(ctypes.byref(plan), ctypes.byref(desc)) == 0
    del t["row"], t["a"], t["c"]  # the plan owns copies: the caller's arrays may go away
    torch.cuda.empty_cache()
    grad = torch.empty(m, device=DEV)
    scal = torch.empty(8, dtype=torch.float64, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):  # accumulators are self-cleaning: repeated calls give identical results
        assert lib.dualip_matching_calc(plan, t["lam"].data_ptr(), t["b"].data_ptr(), gamma, grad.data_ptr(), scal.data_ptr(), None, None, 0, stream) == 0
    torch.cuda.synchronize()
    ref = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, [c_oracle.make_class("simplex", {"z": 1.0})], p["lam"], gamma, p["b"])
    assert abs(float(scal[0]) - ref["scal"][0]) <= 1e-5 * abs(ref["scal"][0])
    assert np.allclose(grad.cpu().numpy(), ref["grad"], rtol=1e-5, atol=1e-5 * np.abs(ref["grad"]).max())
    host_grad = np.empty(m, dtype=np.float32)
    host_scal = _native.Scalars()
    lam_host = np.ascontiguousarray(p["lam"])
    assert lib.dualip_matching_calc_host(plan, lam_host.ctypes.data, t["b"].data_ptr(), gamma, host_grad.ctypes.data, ctypes.byref(host_scal), stream) == 0
    assert np.array_equal(host_grad, grad.cpu().numpy()) or np.allclose(host_grad, grad.cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert abs(host_scal.dual_objective - ref["scal"][0]) <= 1e-5 * abs(ref["scal"][0])
    lib.dualip_plan_destroy(plan)
