"""-m gpu: size-independent properties at benchmark-like sizes (no oracle run needed), plus a sampled oracle check."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def big():
    from benchmark.synthetic import capacity_vector, generate_shard
    from dualip_b200.preprocessing.precondition import jacobi_precondition

    n, m, sp = 2_000_000, 10_000, 1e-3
    sh = generate_shard(n, m, sp, 42, DEV)
    b = capacity_vector(sh.greedy_load, m, sp, 42, DEV)
    A = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.a, size=(m, n))
    C = torch.sparse_csc_tensor(sh.ccol, sh.row, sh.c, size=(m, n))
    jacobi_precondition(A, b)
    return dict(n=n, m=m, sh=sh, A=A, C=C, b=b)


def _objective(big, mixed=True, b=True):
    import bench
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.projections import create_projection_map

    pm = bench.mixed_projection_map(big["n"], 0, DEV) if mixed else create_projection_map("simplex", {"z": 1.0}, big["n"])
    return MatchingSolverDualObjectiveFunction(MatchingInputArgs(big["A"], big["C"], pm, big["b"] if b else None), gamma=1e-3)


def test_projection_invariants_and_consistency(big):
    obj = _objective(big)
    sh, n, m = big["sh"], big["n"], big["m"]
    lam = torch.rand(m, device=DEV) * 50
    r = obj.calculate(lam, save_primal=True)
    x = r.primal_var
    assert bool((x >= 0).all()) and bool(torch.isfinite(x).all())
    col = torch.repeat_interleave(torch.arange(n, device=DEV), sh.ccol[1:] - sh.ccol[:-1])
    colsum = torch.zeros(n, device=DEV, dtype=torch.float64).index_add_(0, col, x.double())
    even = torch.arange(n, device=DEV) % 2 == 0
    # simplex columns: sum x <= z up to fp32 rounding of theta at |v| ~ 500 (ulp 3e-5 per support entry), as in the reference
    assert float(colsum[even].max()) <= 1.0 + 1e-3
    assert float(x[(col % 2) == 1].max()) <= 1.0  # box columns: x <= 1
    # gradient / scalars recomputed from x with plain torch in fp64
    g = torch.zeros(m, device=DEV, dtype=torch.float64).index_add_(0, sh.row, (big["A"].values().double() * x.double())) - big["b"].double()
    assert torch.allclose(r.dual_gradient.double(), g, rtol=1e-5, atol=1e-5 * float(g.abs().max()))
    cx = float((big["C"].values().double() * x.double()).sum())
    xx = float((x.double() ** 2).sum())
    s = r.scalars64
    assert abs(float(s[1]) - cx) <= 1e-7 * abs(cx) and abs(float(s[6]) - xx) <= 1e-7 * xx
    assert abs(float(s[0]) - (cx + 0.5e-3 * xx + float((lam.double() * g).sum()))) <= 1e-6 * abs(float(s[0]))
    # idempotence / determinism of the projection pattern: same lambda -> same x, bit for bit
    x2 = obj.calculate(lam, save_primal=True).primal_var
    assert torch.equal(x, x2)


def test_sampled_columns_against_c_oracle(big):
    from oracle import c_oracle

    obj = _objective(big)
    sh, m = big["sh"], big["m"]
    lam = torch.rand(m, device=DEV) * 120  # the regime of late iterations: many multi-entry supports
    r = obj.calculate(lam, save_primal=True, diagnostics=True)
    n_s = 200_000
    e_s = int(sh.ccol[n_s])
    ccol, row = sh.ccol[: n_s + 1].cpu().numpy(), sh.row[:e_s].cpu().numpy()
    a, c = big["A"].values()[:e_s].cpu().numpy(), big["C"].values()[:e_s].cpu().numpy()
    classes = [c_oracle.make_class("simplex", {"z": 1.0}), c_oracle.make_class("box", {"lower": 0.0, "upper": 1.0})]
    ref = c_oracle.calculate(ccol, row, a, c, m, classes, lam.cpu().numpy(), 1e-3, None, (np.arange(n_s) % 2).astype(np.uint8))
    x = r.primal_var[:e_s].cpu().numpy()
    assert int(((x != 0) != (ref["x"] != 0)).sum()) == 0
    assert (np.abs(x - ref["x"]) / np.maximum(np.abs(ref["x"]), 1e-6)).max() <= 1e-5
    diag = r.projection_diag[:e_s].cpu().numpy()
    nonempty = np.diff(ccol) > 0
    d = diag[ccol[:-1][nonempty]]
    cd = ref["diag"][nonempty]
    sx = cd != 255
    assert np.array_equal(d[sx] & 3, cd[sx] & 3) and np.array_equal((d[sx] >> 2)[(cd[sx] & 3) > 0], (cd[sx] >> 2)[(cd[sx] & 3) > 0])
    assert np.bincount(cd[sx] & 3, minlength=3).min() > 10  # all three branches are exercised


def test_linearity_of_partial_sums_over_shards(big):
    """Sum of per-shard partials == unsharded partial (the algebra the all-reduce relies on), at 2M entities."""
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.projections import create_projection_map
    from dualip_b200.utils.dist_utils import global_to_local_projection_map, split_tensors_to_devices

    m, n = big["m"], big["n"]
    lam = torch.rand(m, device=DEV) * 80
    whole = _objective(big, mixed=False, b=False)
    part = torch.empty(m + 2, device=DEV)
    whole.launch_partial(lam.data_ptr(), 1e-3, part.data_ptr())
    a_s, c_s, index_map = split_tensors_to_devices(big["A"], big["C"], [DEV] * 4)
    pm = create_projection_map("simplex", {"z": 1.0}, n)
    total = torch.zeros(m + 2, device=DEV, dtype=torch.float64)
    for k in range(4):
        lo = index_map[k][0]
        o = MatchingSolverDualObjectiveFunction(MatchingInputArgs(a_s[k], c_s[k], global_to_local_projection_map(pm, range(lo, lo + len(index_map[k]))), None), 1e-3)
        pk = torch.empty(m + 2, device=DEV)
        o.launch_partial(lam.data_ptr(), 1e-3, pk.data_ptr())
        total += pk.double()
    assert torch.allclose(total, part.double(), rtol=1e-5, atol=1e-5 * float(part.abs().max()))


def test_fixed_point_gradient_is_reproducible_and_agrees_with_fp32_atomics(big, monkeypatch):
    """Bounded projection classes accumulate the dual gradient in 32-bit fixed point with native shared-memory integer
    adds: the result does not depend on the order of the adds (bitwise reproducible) and agrees with the fp32-atomics
    build of the same plan to fp32 accuracy.  Open cones have no overflow bound and keep fp32 atomics."""
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.projections import create_projection_map

    m = big["m"]
    lam = torch.rand(m, device=DEV) * 60
    fx = _objective(big)
    info = fx.plan_info()
    assert info["fixed_point"] == 1 and info["fixed_point_relerr_e12"] <= 240_000  # <= 2.4e-7 (4 fp32 ulps)
    g1 = fx.calculate(lam).dual_gradient.clone()
    for _ in range(3):
        assert torch.equal(fx.calculate(lam).dual_gradient, g1)
    monkeypatch.setenv("DUALIP_ACCUM", "f32")
    f32 = _objective(big)
    assert f32.plan_info()["fixed_point"] == 0
    r32, rfx = f32.calculate(lam), fx.calculate(lam)
    scale = float(rfx.dual_gradient.abs().max())
    assert torch.allclose(r32.dual_gradient, rfx.dual_gradient, rtol=2e-6, atol=2e-6 * scale)
    assert abs(float(r32.scalars64[0]) - float(rfx.scalars64[0])) <= 1e-6 * abs(float(rfx.scalars64[0]))
    monkeypatch.delenv("DUALIP_ACCUM")
    cone = MatchingSolverDualObjectiveFunction(
        MatchingInputArgs(big["A"], big["C"], create_projection_map("cone", {"lower": 0.0}, big["n"]), big["b"]), gamma=1e-3)
    assert cone.plan_info()["fixed_point"] == 0


@pytest.mark.parametrize("stage", [20, 12, 7])
def test_tma_staged_slabs_are_bit_identical_to_plain_loads(big, monkeypatch, stage):
    """Staging (per-warp shared-memory buffer filled by cp.async.bulk while the previous slab is processed) only changes how
    a slab reaches the registers: x and the fixed-point gradient must not change by a bit, whatever mix of staged and
    unstaged slabs a warp sees (the degrees move the boundary through the column-length range)."""
    lam = torch.rand(big["m"], device=DEV) * 40
    monkeypatch.setenv("DUALIP_STAGE", "0")
    plain = _objective(big)
    assert plain.plan_info()["staged_degree"] == 0
    r0 = plain.calculate(lam, save_primal=True)
    monkeypatch.setenv("DUALIP_STAGE", str(stage))
    staged = _objective(big)
    assert staged.plan_info()["staged_degree"] == stage
    for _ in range(2):
        r1 = staged.calculate(lam, save_primal=True)
        assert torch.equal(r1.primal_var, r0.primal_var)
        assert torch.equal(r1.dual_gradient, r0.dual_gradient)
        assert abs(float(r1.scalars64[0]) - float(r0.scalars64[0])) <= 1e-9 * abs(float(r0.scalars64[0]))


def test_in_place_edits_after_construction_are_detected():
    """The plan snapshots A's and c's values (the reference reads the tensors at every call, matching.py:136-142): an in-place
    change after the objective was built -- an element-wise op, or jacobi_precondition called too late -- raises instead of
    being silently ignored, through calculate() and through the Maximizer's raw launches."""
    from conftest import random_problem
    from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
    from dualip_b200.preprocessing.precondition import jacobi_precondition
    from dualip_b200.projections import create_projection_map

    p = random_problem(9, 500, 32, 6.0)
    n, m = p["ccol"].size - 1, 32
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])

    def build():
        A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]).clone(), size=(m, n)).to(DEV)
        C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]).clone(), size=(m, n)).to(DEV)
        b = torch.from_numpy(p["b"]).to(DEV)
        return A, C, b, MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), b), gamma=1e-2)

    lam = torch.zeros(m, device=DEV)
    A, C, b, obj = build()
    obj.calculate(lam)
    C.values().mul_(2.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        obj.calculate(lam)
    A, C, b, obj = build()
    jacobi_precondition(A, b)
    with pytest.raises(RuntimeError, match="modified in place"):
        AcceleratedGradientDescent(max_iter=3, gamma=1e-2, iteration_callback=no_iteration_callback).maximize(obj, lam)
    A, C, b, _ = build()
    jacobi_precondition(A, b)  # preprocessing first, then the objective: fine
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), b), gamma=1e-2)
    assert torch.isfinite(obj.calculate(lam).dual_objective)
