"""Host-side mirror of the reference API: registry, projection-map keys, sharding helpers, Maximizer on CPU objectives,
step-size rule.  Modelled on the reference's tests/test_agd.py, tests/test_utils.py, tests/test_dist_utils.py."""
import math

import numpy as np
import pytest
import torch

from dualip_b200 import _native
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction, _build_class_table
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, project_on_nn_cone
from dualip_b200.optimizers.agd_utils import (
    calculate_step_size,
    estimate_lipschitz_constant,
    step_size_from_lipschitz_constants,
    update_dual_gradient_history,
)
from dualip_b200.projections import ProjectionEntry, create_projection_map, project
from dualip_b200.run_solver import run_solver
from dualip_b200.types import ComputeArgs, ObjectiveArgs, ObjectiveResult, SolverArgs, SolverResult
from dualip_b200.utils.dist_utils import global_to_local_projection_map, shard_sizes, split_tensors_to_devices
from dualip_b200.utils.sparse_utils import split_csc_by_cols
from conftest import GOLDEN, random_problem
from oracle import dualip_oracle as O


# ---- types / defaults (reference types.py:7-50) ----
def test_dataclass_defaults_match_reference():
    s = SolverArgs()
    assert (s.max_iter, s.initial_step_size, s.gamma, s.max_step_size) == (10000, 1e-5, 1e-3, 0.1)
    assert s.initial_dual_path is None and s.gamma_decay_type is None and s.save_primal is False
    assert ComputeArgs(host_device="cuda:0").compute_device_num == 1
    o = ObjectiveArgs(objective_type="matching")
    assert o.use_jacobi_precondition is False and o.objective_kwargs is None
    r = ObjectiveResult(dual_gradient=torch.zeros(1), dual_objective=torch.tensor(0.0))
    assert r.reg_penalty is None and r.primal_var is None and r.max_pos_slack is None
    assert [f for f in SolverResult.__dataclass_fields__] == ["dual_val", "dual_objective", "objective_result",
                                                               "dual_objective_log", "step_size_log"]


# ---- projections registry (reference projections/base.py) ----
def test_registry_and_keys():
    assert project("box").native_class().kind == _native.PROJ_CLAMP
    c = project("cone", lower=0.0).native_class()
    assert c.lo == 0.0 and math.isinf(c.hi)
    c = project("cone").native_class()
    assert math.isinf(c.lo) and math.isinf(c.hi)
    sx = project("simplex", z=2.0).native_class()
    assert sx.kind == _native.PROJ_SIMPLEX and sx.z == 2.0 and sx.z_thr == float(np.float32(2.0 + 1e-6))
    assert project("simplex_eq").native_class().kind == _native.PROJ_SIMPLEX_EQ
    with pytest.raises(ValueError):
        project("nope")
    with pytest.raises(ValueError):
        project("cone", lower=0.0, upper=1.0)
    with pytest.raises(ValueError):
        project("simplex", method="newton")
    with pytest.raises(TypeError):
        project("box", l=0, u=1)  # the reference's operators take lower/upper
    pm = create_projection_map("simplex", {"z": 1}, 5)
    assert list(pm) == ["simplex_z_1"] and list(pm["simplex_z_1"].indices) == [0, 1, 2, 3, 4]
    pm = create_projection_map("box", {"upper": 1.0, "lower": 0.0}, 10, indices=[0, 2], key_prefix="p_")
    assert list(pm) == ["p_box_lower_0.0_upper_1.0"] and pm["p_box_lower_0.0_upper_1.0"].indices == [0, 2]


def test_projection_call_requires_cuda():
    with pytest.raises(RuntimeError, match="CUDA"):
        project("simplex")(torch.zeros(3, 2))


def test_objective_requires_cuda_and_csc():
    A = torch.eye(3).to_sparse_csc()
    args = MatchingInputArgs(A, A, create_projection_map("simplex", {"z": 1.0}, 3), torch.ones(3))
    with pytest.raises(RuntimeError, match="CUDA"):
        MatchingSolverDualObjectiveFunction(args, gamma=1e-3)
    with pytest.raises(ValueError):
        MatchingSolverDualObjectiveFunction(MatchingInputArgs(torch.eye(3), A, {}, torch.ones(3)), gamma=1e-3)


def test_class_table_from_projection_map():
    ccol = torch.tensor([0, 1, 3, 3, 6, 7])  # lengths 1,2,0,3,1
    pm = create_projection_map("simplex", {"z": 1.0}, 5)
    classes, n, col_class = _build_class_table(ccol, 5, pm, batching=True)
    assert n == 1 and col_class is None and classes[0].flags == 0  # a 2-entry column pads the {1,2} bucket
    pm = {}
    pm.update(create_projection_map("simplex", {"z": 1.0}, 5, indices=[0, 4]))
    pm.update(create_projection_map("box", {"lower": 0.0, "upper": 1.0}, 5, indices=[1, 3]))
    classes, n, col_class = _build_class_table(ccol, 5, pm, batching=True)
    assert n == 3 and col_class.tolist() == [1, 2, 0, 2, 1]
    assert classes[0].kind == _native.PROJ_CLAMP and math.isinf(classes[0].hi)  # identity for unlisted columns
    assert classes[1].flags & _native.PROJ_FLAG_D1_UNPADDED  # its 1-entry columns have no 2-entry bucket mate
    with pytest.raises(ValueError, match="overlaps"):
        bad = dict(pm)
        bad.update(create_projection_map("cone", {"lower": 0.0}, 5, indices=[3], key_prefix="x"))
        _build_class_table(ccol, 5, bad, batching=True)
    with pytest.raises(IndexError):
        _build_class_table(ccol, 5, create_projection_map("box", {}, 5, indices=[7]), batching=True)


# ---- sharding helpers (reference utils/dist_utils.py, tests/test_dist_utils.py) ----
def test_shard_sizes_and_split():
    assert shard_sizes(5, 2) == [3, 2] and shard_sizes(7, 3) == [3, 2, 2] and shard_sizes(6, 2) == [3, 3]
    dense = torch.arange(30, dtype=torch.float32).reshape(5, 6) * (torch.rand(5, 6) > 0.3)
    a = dense.to_sparse_csc()
    parts = split_csc_by_cols(a, [2, 3, 1])
    assert torch.equal(torch.cat([p.to_dense() for p in parts], dim=1), dense)
    with pytest.raises(ValueError):
        split_csc_by_cols(a, [2, 2])
    a_s, c_s, index_map = split_tensors_to_devices(a, a, ["cpu", "cpu"])
    assert a_s[0].shape == (5, 3) and c_s[1].shape == (5, 3) and index_map == [[0, 1, 2], [3, 4, 5]]
    a_s, c_s, index_map = split_tensors_to_devices(a, a, [])
    assert len(a_s) == 1 and index_map == list(range(6))


def test_global_to_local_projection_map():
    pm = create_projection_map("simplex_ineq", {"z": 1}, 6)  # the reference's test uses this unregistered name
    loc = [global_to_local_projection_map(pm, cols) for cols in ([0, 1, 2], [3, 4, 5])]
    assert list(loc[0]["simplex_ineq_z_1"].indices) == [0, 1, 2] and list(loc[1]["simplex_ineq_z_1"].indices) == [0, 1, 2]
    pm = {**create_projection_map("simplex", {"z": 1}, 10, indices=[0, 1]),
          **create_projection_map("simplex_eq", {"z": 2}, 10, indices=[2, 3, 4, 5, 6, 7, 8, 9])}
    l0 = global_to_local_projection_map(pm, list(range(0, 5)))
    l1 = global_to_local_projection_map(pm, range(5, 10))
    assert l0["simplex_z_1"].indices == [0, 1] and l0["simplex_eq_z_2"].indices == [2, 3, 4]
    assert "simplex_z_1" not in l1 and l1["simplex_eq_z_2"].indices == [0, 1, 2, 3, 4]
    big = create_projection_map("box", {}, 10**9)  # range-based: no per-column dictionary
    loc = global_to_local_projection_map(big, range(250_000_000, 500_000_000))
    assert loc["box_"].indices == range(0, 250_000_000)


# ---- Maximizer on host objectives (reference tests/test_agd.py) ----
class _Quadratic2D:
    equality_mask = None

    def calculate(self, dual_val, save_primal=False, **kwargs):
        x, y = dual_val
        obj = -((x - 3.0) ** 2) - (y + 5.0) ** 2
        return ObjectiveResult(dual_gradient=torch.tensor([-2.0 * (x - 3.0), -2.0 * (y + 5.0)]), dual_objective=obj)


def test_agd_known_answers_on_cpu_objective():
    solver = AcceleratedGradientDescent(max_iter=30, gamma=None, initial_step_size=1e-5, iteration_callback=lambda i, r: None)
    res = solver.maximize(_Quadratic2D(), torch.tensor([0.0, 0.0]))
    for i, true_val in [(2, -33.9996400036), (16, -28.60551547593112), (23, -25.473701313626133), (29, -25.00382134903756)]:
        assert abs(res.dual_objective_log[i - 1] - true_val) < 1e-5
    one = AcceleratedGradientDescent(max_iter=1, gamma=None, initial_step_size=0.1, iteration_callback=lambda i, r: None)
    assert abs(float(one.maximize(_Quadratic2D(), torch.tensor([0.0, 0.0])).dual_val[0]) - 0.6) < 1e-6



class _RandomConcave:
    """Concave quadratic with a fixed random curvature; rows 0..4 are equalities (free multipliers)."""

    def __init__(self, m=300):
        g = torch.Generator().manual_seed(5)
        self.q = torch.rand(m, generator=g) * 40 + 0.5
        self.t = torch.randn(m, generator=g)
        self.equality_mask = torch.zeros(m, dtype=torch.bool)
        self.equality_mask[:5] = True

    def calculate(self, dual_val, gamma=None, **kwargs):
        scale = 1.0 if gamma is None else gamma / 1e-2
        d = dual_val - self.t
        return ObjectiveResult(dual_gradient=-(self.q * d) * scale, dual_objective=-0.5 * (self.q * d * d).sum() * scale)


@pytest.mark.parametrize("decay", [False, True])
def test_native_host_step_matches_reference_style_loop(decay, monkeypatch):
    """dualip_agd_host_step (one native call per iteration) against the reference's tensor-op loop and the numpy oracle:
    history ring, first-14-iterations rule, step cap on gamma decay, equality rows, momentum."""
    f = _RandomConcave()
    kw = dict(gamma_decay_type="step", gamma_decay_params={"decay_steps": 7, "decay_factor": 0.5}) if decay else {}
    runs = {}
    for mode in ("native", "torch"):
        monkeypatch.setenv("DUALIP_HOST_STEP", mode)
        solver = AcceleratedGradientDescent(max_iter=45, gamma=1e-2, initial_step_size=1e-3, max_step_size=0.1,
                                            iteration_callback=lambda i, r: None, **kw)
        runs[mode] = (solver.maximize(f, torch.zeros(300)), solver.gamma, solver.max_step_size)
    (a, ga, ma), (b, gb, mb) = runs["native"], runs["torch"]
    assert ga == gb and abs(ma - mb) <= 1e-6 * mb
    assert np.allclose(a.step_size_log, b.step_size_log, rtol=1e-5)
    assert np.allclose(a.dual_objective_log, b.dual_objective_log, rtol=1e-4, atol=1e-6)
    assert torch.allclose(a.dual_val, b.dual_val, rtol=1e-4, atol=1e-5)
    assert bool((a.dual_val[5:] >= 0).all()) and bool((a.dual_val[:5] < 0).any())  # equality rows stay free

    def calc(lam, g):
        r = f.calculate(torch.from_numpy(lam), gamma=g)
        return r.dual_gradient.numpy(), np.float32(r.dual_objective)

    y, obj_log, step_log, _ = O.agd_maximize(calc, np.zeros(300, np.float32), 45, 1e-2, 1e-3, 0.1,
                                             equality_mask=f.equality_mask.numpy(), **kw)
    assert np.allclose(a.step_size_log, step_log, rtol=1e-5) and np.allclose(a.dual_val.numpy(), y, rtol=1e-4, atol=1e-5)


def test_beta_sequence_and_cone_projection():
    s = AcceleratedGradientDescent(max_iter=50, gamma=1e-3)
    assert np.array_equal(s.beta_seq.numpy(), O.compute_beta_seq(50))
    y = torch.tensor([-1.0, 2.0, -3.0])
    assert project_on_nn_cone(y).tolist() == [0.0, 2.0, 0.0]
    assert project_on_nn_cone(y, torch.tensor([True, False, False])).tolist() == [-1.0, 2.0, 0.0]
    with pytest.raises(ValueError):
        AcceleratedGradientDescent(max_iter=2, gamma=1.0, gamma_decay_type="exp", gamma_decay_params={})._update_gamma(1, 0.1)


def test_step_size_rule():
    gh, dh = [], []
    for k in range(3):
        update_dual_gradient_history(torch.tensor([float(k)]), torch.tensor([float(k)]), gh, dh, 2)
    assert len(gh) == 2 and gh[0].item() == 1.0
    L = estimate_lipschitz_constant(torch.tensor([0.0]), torch.tensor([2.0]), torch.tensor([0.0]), torch.tensor([1.0]))
    assert L.item() == 2.0
    assert step_size_from_lipschitz_constants([L] * 3, 15, 1e-5, 0.1) == 1e-5  # history incomplete
    assert step_size_from_lipschitz_constants([torch.tensor(float("nan"))] + [L] * 13, 15, 1e-5, 0.1) == 1e-5
    assert step_size_from_lipschitz_constants([L] * 14, 15, 1e-5, 0.1) == 0.1  # 1/2 clamped to max_step_size
    assert step_size_from_lipschitz_constants([torch.tensor(100.0)] * 14, 15, 1e-5, 0.1) == 0.01
    assert step_size_from_lipschitz_constants([torch.tensor(0.0)] * 14, 15, 1e-5, 0.1) == 0.1
    gh, dh = [], []
    assert calculate_step_size(torch.tensor([1.0]), torch.tensor([0.0]), gh, dh) == 1e-5


def test_numpy_step_size_utility_follows_the_tensor_rule_and_the_reference_module():
    """utils/step_size_utility.py (numpy iterates): same steps as optimizers/agd_utils.py on the same sequence, and -- where
    the reference tree is present -- as the reference's own module (src/dualip/utils/step_size_utility.py)."""
    import importlib.util
    import os

    import numpy as np

    from dualip_b200.utils import step_size_utility as S

    rng = np.random.default_rng(3)
    seq = [(rng.standard_normal(40), rng.standard_normal(40)) for _ in range(33)]
    # a repeated pair gives a 0/0 = NaN estimate: Python's max() skips it unless it is the FIRST of the ring (agd_utils.py:58-60)
    seq[17] = (seq[16][0].copy(), seq[16][1].copy())
    ref_path = "/root/reference/src/dualip/utils/step_size_utility.py"
    R = None
    if os.path.exists(ref_path):
        spec = importlib.util.spec_from_file_location("_ref_step_size_utility", ref_path)
        R = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(R)
    gh, dh, gt, dt, gr, dr = [], [], [], [], [], []
    steps = []
    with np.errstate(invalid="ignore", divide="ignore"):
        for g, d in seq:
            s_np = S.calculate_step_size(g, d, gh, dh, 15, 1e-5, 0.1)
            s_t = calculate_step_size(torch.from_numpy(g), torch.from_numpy(d), gt, dt, 15, 1e-5, 0.1)
            assert s_np == pytest.approx(s_t, rel=1e-12)
            if R is not None:
                assert s_np == R.calculate_step_size(g, d, gr, dr, 15, 1e-5, 0.1)
            steps.append(s_np)
    assert len(gh) == len(dh) == 15 and steps[:14] == [1e-5] * 14 and steps[14] == 0.1 and steps[17] == 0.1
    assert steps[30] == 1e-5 and steps[31] == 0.1  # the NaN estimate heads the ring exactly once
    assert S.estimate_lipschitz_constant(np.zeros(1), np.full(1, 2.0), np.zeros(1), np.ones(1)) == 2.0
    assert S.step_size_from_lipschitz_constants([0.0] * 14, 15, 1e-5, 0.1) == 0.1


def test_run_solver_argument_errors():
    A = torch.eye(3).to_sparse_csc()
    args = MatchingInputArgs(A, A, create_projection_map("simplex", {"z": 1.0}, 3), torch.ones(3))
    with pytest.raises(ValueError):
        run_solver(args, SolverArgs(max_iter=1), ComputeArgs("cpu"), ObjectiveArgs("unknown"))
    with pytest.raises(RuntimeError, match="CUDA"):
        run_solver(args, SolverArgs(max_iter=1), ComputeArgs("cpu"), ObjectiveArgs("matching"))


def test_install_as_alias():
    import importlib
    import sys

    import dualip_b200

    dualip_b200.install_as("dualip_alias_for_test")
    mod = importlib.import_module("dualip_alias_for_test.objectives.matching")
    assert mod.MatchingInputArgs is MatchingInputArgs
    for k in [k for k in sys.modules if k.startswith("dualip_alias_for_test")]:
        del sys.modules[k]


def test_reference_cache_layout_shard_direct_reader(tmp_path):
    """dualip_b200.utils.data_cache against the cache layout of the reference's generator
    (benchmark/generate_synthetic_data.py:172-342): metadata identical to what the reference itself wrote for this
    problem (kept in the fixture), shards read directly by column range, rebased, c negated."""
    import json
    import os
    import sys

    from dualip_b200.utils import data_cache

    d = np.load(os.path.join(GOLDEN, "cfg2_synthetic.npz"))
    ref_meta = json.loads(str(d["meta_json"]))
    n, m = ref_meta["num_sources"], ref_meta["num_destinations"]
    key = dict(num_sources=n, num_destinations=m, target_sparsity=ref_meta["target_sparsity"], dtype=torch.float32, seed=ref_meta["seed"])
    assert data_cache.cache_prefix(**key) + "_meta.json" == str(d["meta_name"])
    # the reference stores c positive (float64 from its numpy generator) and negates on load (:448)
    c_pos = (-d["c"]).astype(np.dtype(ref_meta["array_dtypes"]["c_vals"]))
    a = d["a"].astype(np.dtype(ref_meta["array_dtypes"]["A_vals"]))
    b = d["b"].astype(np.dtype(ref_meta["array_dtypes"]["b_vec"]))
    prefix = data_cache.save_cache(str(tmp_path), key, d["ccol"], d["row"], a, c_pos, b)
    assert json.load(open(tmp_path / f"{prefix}_meta.json")) == ref_meta
    whole = data_cache.load_shard(str(tmp_path), prefix)
    assert torch.equal(whole.ccol, torch.from_numpy(d["ccol"])) and torch.equal(whole.row, torch.from_numpy(d["row"]))
    assert torch.equal(whole.a, torch.from_numpy(d["a"])) and torch.equal(whole.c, torch.from_numpy(d["c"]))
    assert torch.equal(whole.b, torch.from_numpy(d["b"])) and whole.a.dtype == torch.float32
    world, parts = 3, []
    for r in range(world):
        s = data_cache.load_shard(str(tmp_path), prefix, rank=r, world=world)
        assert int(s.ccol[0]) == 0 and s.ccol.numel() == shard_sizes(n, world)[r] + 1
        A, C = s.csc()
        assert A.shape == (m, s.col_end - s.col_start) and torch.equal(A.ccol_indices(), C.ccol_indices())
        parts.append(s)
    assert torch.equal(torch.cat([s.row for s in parts]), whole.row) and torch.equal(torch.cat([s.c for s in parts]), whole.c)
    assert [s.col_start for s in parts] == [0, 6667, 13334] and parts[-1].col_end == n
    with pytest.raises(ValueError):
        data_cache.load_shard(str(tmp_path), prefix, col_range=(5, n + 1))
    if os.path.isdir("/root/reference/benchmark"):  # build container only: the reference's own loader accepts our files
        sys.path.insert(0, "/root/reference")
        sys.path.insert(0, "/root/reference/src")
        try:
            from benchmark import generate_synthetic_data as ref_gen
        except Exception:
            ref_gen = None
        finally:
            sys.path.remove("/root/reference")
            sys.path.remove("/root/reference/src")
        if ref_gen is not None and hasattr(ref_gen, "_load_cached_numpy"):
            got = ref_gen._load_cached_numpy(ref_gen._get_cache_key(n, m, key["target_sparsity"], torch.float32, key["seed"]), str(tmp_path))
            assert got is not None and np.array_equal(got[0], d["ccol"]) and np.array_equal(got[3], c_pos)


@pytest.mark.parametrize("batching", [True, False])
def test_block_entry_buckets_follow_the_reference_rule(batching):
    """_BlockEntry (padded-block route for user-registered projections) buckets columns like matching.py:87-114:
    thresholds {1,2},{3,4},{5..8},..., empty columns dropped, one bucket without batching."""
    from conftest import random_csc
    from dualip_b200.objectives.matching import _BlockEntry

    rng = np.random.default_rng(3)
    n_rows = 40
    ccol, row = random_csc(rng, 500, n_rows, 6.0, long_cols=[(3, 33), (8, 17)])
    cols = np.arange(0, 500, 3)
    ent = _BlockEntry("k", ProjectionEntry("user_op", {}, cols.tolist()), torch.from_numpy(cols), torch.from_numpy(ccol), n_rows, batching)
    ref = O.compute_buckets(ccol, n_rows, cols, batching=batching)
    ref = [b for b in ref if np.diff(ccol)[b].sum() > 0]
    assert len(ent.buckets) == len(ref)
    lens = np.diff(ccol)
    seen = []
    for (off, total, idx_in_col, cols_rep, L, K), rb in zip(ent.buckets, ref):
        assert K == rb.size and L == lens[rb].max() and total == lens[rb].sum()
        e = ent.entries[off: off + total].numpy()
        assert np.array_equal(e, ccol[rb][cols_rep.numpy()] + idx_in_col.numpy())
        seen.append(e)
    expect = np.concatenate([np.arange(ccol[j], ccol[j + 1]) for j in cols])
    assert np.array_equal(np.sort(np.concatenate(seen)), np.sort(expect))


def test_block_entry_partial_sums_against_oracle_on_cpu():
    """The tensor-op half of the padded-block route (add_block_entries: v, blocks per bucket, operator, row sums, c.x, ||x||^2)
    checked on CPU tensors against the numpy oracle, with a user-style operator that equals box[0, 0.4]."""
    from dualip_b200.objectives.matching import _BlockEntry
    from dualip_b200.projections.base import ProjectionOperator, register

    @register("cpu_test_capped_box")
    class _Capped(ProjectionOperator):
        def __init__(self, upper=0.4):
            self.upper = upper

        def __call__(self, x):
            return torch.clamp(x, min=0.0, max=self.upper)

    p = random_problem(9, 800, 30, 6.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma = p["n_cols"], p["n_rows"], 3e-2
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    a, c = torch.from_numpy(p["a"]), torch.from_numpy(p["c"])
    cols = torch.arange(1, n, 2)
    entry = ProjectionEntry("cpu_test_capped_box", {"upper": 0.4}, cols.tolist())
    obj = object.__new__(MatchingSolverDualObjectiveFunction)  # only the fields add_block_entries reads
    obj.device, obj.m = torch.device("cpu"), m
    obj._block_entries = [_BlockEntry("k", entry, cols, ccol, m, True)]
    ec = obj._block_entries[0].entries
    obj._blk_idx, obj._blk_a, obj._blk_c, obj._blk_row = ec, a[ec], c[ec], row[ec]
    partial = torch.zeros(m + 2)
    x_out = torch.zeros(row.numel())
    obj.add_block_entries(torch.from_numpy(p["lam"]), gamma, partial, x_out)
    # oracle: the same columns with box[0, 0.4]; the other columns contribute nothing (clamp to [0, 0] in the kernel)
    pm = {"u": O.ProjEntry("box", {"lower": 0.0, "upper": 0.4}, cols.numpy()),
          "z": O.ProjEntry("box", {"lower": 0.0, "upper": 0.0}, np.arange(0, n, 2))}
    r = O.matching_calculate(p["ccol"], p["row"], p["a"], p["c"], m, pm, p["lam"], gamma, None)
    assert np.array_equal(x_out.numpy(), r.primal_var)
    assert np.allclose(partial[:m].numpy(), r.dual_gradient, rtol=1e-5, atol=1e-5)
    assert abs(float(partial[m]) - r.primal_objective) <= 1e-5 * abs(r.primal_objective)
    assert abs(0.5 * gamma * float(partial[m + 1]) - r.reg_penalty) <= 1e-5 * abs(r.reg_penalty)


def test_vectorised_input_validation():
    """preprocessing/input_validation.py: same checks, exception type and messages as the reference (:4-103) without its
    per-column Python loop."""
    from conftest import random_csc
    from dualip_b200.preprocessing.input_validation import (InputValidationError, check_correct_csc_construction,
                                                           check_nan_or_inf, check_no_zero_row_or_col, run_all_checks)

    rng = np.random.default_rng(4)
    ccol, row = random_csc(rng, 300_000, 50, 6.0)
    vals = (rng.random(row.size) + 0.1).astype(np.float32)

    def csc(cc=ccol, rr=row, vv=vals, m=50):
        return torch.sparse_csc_tensor(torch.from_numpy(cc), torch.from_numpy(rr), torch.from_numpy(vv), size=(m, cc.size - 1))

    run_all_checks(csc())  # 3e5 columns, 1.8e6 entries: vectorised, well under a second
    j = int(np.nonzero(np.diff(ccol) >= 3)[0][1234])
    bad = row.copy()
    bad[ccol[j]], bad[ccol[j] + 1] = bad[ccol[j] + 1], bad[ccol[j]]  # column j no longer increasing
    with pytest.raises(InputValidationError, match=f"row indices in column {j} are not strictly increasing"):
        check_correct_csc_construction(csc(rr=bad))
    dup = row.copy()
    dup[ccol[j] + 1] = dup[ccol[j]]  # duplicate row index
    with pytest.raises(InputValidationError, match=f"column {j} "):
        check_correct_csc_construction(csc(rr=dup))
    cc2 = ccol.copy()
    cc2[10] = cc2[11] + 1
    with pytest.raises(InputValidationError, match="non-decreasing"):
        check_correct_csc_construction(csc(cc=cc2))
    v0 = vals.copy()
    v0[77] = 0.0
    with pytest.raises(InputValidationError, match="No zeroes"):
        check_correct_csc_construction(csc(vv=v0))
    vn = vals.copy()
    vn[5] = np.inf
    with pytest.raises(InputValidationError, match="nan or infinite"):
        check_nan_or_inf(csc(vv=vn))
    with pytest.raises(InputValidationError, match="all-zero row"):
        check_no_zero_row_or_col(csc(m=51))  # row 50 never appears
    dense = torch.tensor([[1.0, 0.0], [2.0, 0.0]])
    with pytest.raises(InputValidationError, match="all-zero column"):
        check_no_zero_row_or_col(dense)
    assert issubclass(InputValidationError, ValueError)


def test_gamma_schedule_equals_the_reference_loop():
    """The device-resident schedule of the graph-replayed loop (gamma per iteration, step-cap flags) against the reference's
    own bookkeeping: AcceleratedGradientDescent._update_gamma called once per iteration (agd.py:102-109,186-187)."""
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, gamma_schedule

    for gamma0, steps, factor, n in ((1e-3, 35, 0.7, 200), (0.1, 1, 0.5, 12), (2.0, 7, 0.9, 7), (5e-2, 1000, 0.1, 50)):
        solver = AcceleratedGradientDescent(max_iter=n, gamma=gamma0, gamma_decay_type="step",
                                            gamma_decay_params={"decay_steps": steps, "decay_factor": factor})
        seen, caps = [], []
        for i in range(1, n + 1):
            seen.append(solver.gamma)
            before = solver.max_step_size
            solver._update_gamma(i, 0.5 + 1e-3 * i)  # the step size only feeds the cap (varied, so that every change shows)
            caps.append(1 if solver.max_step_size != before else 0)
        gammas, flags = gamma_schedule(gamma0, n, steps, factor)
        assert gammas[:n] == seen and gammas[n] == solver.gamma  # bit-identical doubles
        assert flags == caps
    gammas, flags = gamma_schedule(3e-3, 5)
    assert gammas == [3e-3] * 6 and flags == [0] * 5
