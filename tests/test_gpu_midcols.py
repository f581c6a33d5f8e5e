"""-m gpu: columns of 21..1024 entries.  Plans with the register path keep them column-contiguous and process them a warp
per column inside the fused launch (csrc/mid_col.cuh); DUALIP_MID=0 keeps them in the lane-per-column slab layout (the generic
path).  Two independent implementations of the same reference semantics (simplex.py:143-236 per column): they must agree
bit for bit on x, on the projection branch and on the support size, and both with the oracle."""
import numpy as np
import pytest
import torch

from conftest import random_csc
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.projections import create_projection_map
from oracle import dualip_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(seed, n=3000, m=1500, scale=1.0):
    rng = np.random.default_rng(seed)
    deg = np.clip(rng.lognormal(4.0, 1.0, n).astype(np.int64), 1, 1024)
    deg[:40] = rng.integers(21, 40, 40)        # one entry per lane
    deg[40:60] = [1024, 1023, 993, 992, 961, 960, 33, 32, 31, 64, 65, 128, 129, 256, 257, 512, 513, 21, 22, 1000]
    deg[60:70] = rng.integers(1025, 1400, 10)  # beyond the mid path: the long-column kernel
    deg[70:90] = rng.integers(1, 21, 20)       # register path
    deg[90] = 0
    ccol = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=ccol[1:])
    row = np.concatenate([np.sort(rng.choice(m, size=d, replace=False)) for d in deg]).astype(np.int64)
    E = row.size
    c = (-rng.choice(np.arange(0.5, 5.01, 0.5), size=E) * scale).astype(np.float32)
    a = np.where(rng.random(E) < 0.5, 1.0, rng.lognormal(0, 0.5, E)).astype(np.float32)
    b = np.full(m, 0.3, dtype=np.float32)
    lam = (rng.random(m) * 4.0 * scale).astype(np.float32)
    return dict(ccol=ccol, row=row, a=a, c=c, b=b, lam=lam, n_rows=m)


def _objective(p, pm, gamma, batching=True):
    m, n = int(p["n_rows"]), p["ccol"].size - 1
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]), size=(m, n)).to(DEV)
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]), size=(m, n)).to(DEV)
    return MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(p["b"]).to(DEV)), gamma=gamma, batching=batching)


def _maps(n):
    third = [list(range(k, n, 3)) for k in range(3)]
    mixed = {}
    mixed.update(create_projection_map("simplex", {"z": 1.0}, n, indices=third[0]))
    mixed.update(create_projection_map("simplex_eq", {"z": 2.5}, n, indices=third[1]))
    mixed.update(create_projection_map("box", {"lower": 0.0, "upper": 0.7}, n, indices=third[2]))
    return {"simplex": create_projection_map("simplex", {"z": 1.0}, n), "simplex_z": create_projection_map("simplex", {"z": 3.0}, n),
            "mixed": mixed}


@pytest.mark.parametrize("which", ["simplex", "simplex_z", "mixed"])
@pytest.mark.parametrize("gamma,scale", [(0.1, 1.0), (2.0, 1.0), (30.0, 0.2)])
def test_warp_per_column_path_equals_the_generic_path_and_the_oracle(monkeypatch, which, gamma, scale):
    """gamma spreads the iterate over the branches: small gamma -> top-2 shortcut and short supports, large gamma -> long
    supports and feasible columns (and, for simplex_eq, sums below z)."""
    p = _problem(7 + int(gamma * 10), scale=scale)
    n, m = p["ccol"].size - 1, int(p["n_rows"])
    lam = torch.from_numpy(p["lam"]).to(DEV)
    monkeypatch.setenv("DUALIP_MID", "0")
    slab = _objective(p, _maps(n)[which], gamma)
    monkeypatch.delenv("DUALIP_MID")
    mid = _objective(p, _maps(n)[which], gamma)
    i0, i1 = slab.plan_info(), mid.plan_info()
    assert i0["n_mid_cols"] == 0 and i1["n_mid_cols"] > 2000 and i1["n_long_cols"] == i0["n_long_cols"] == 10
    assert i1["launches_per_calc"] == 2
    r0 = slab.calculate(lam, save_primal=True, diagnostics=True)
    r1 = mid.calculate(lam, save_primal=True, diagnostics=True)
    assert torch.equal(r0.primal_var, r1.primal_var), "x differs between the two implementations"
    assert torch.equal(r0.projection_diag, r1.projection_diag), "projection branch / support size differ"
    assert torch.allclose(r0.dual_gradient, r1.dual_gradient, rtol=2e-5, atol=2e-5)
    assert abs(float(r0.scalars64[0]) - float(r1.scalars64[0])) <= 1e-6 * abs(float(r0.scalars64[0]))
    r2 = mid.calculate(lam)  # the launch without output takes the same decisions
    assert torch.allclose(r2.dual_gradient, r1.dual_gradient, rtol=1e-6, atol=1e-6)
    # the oracle (numpy restatement of the reference on padded blocks)
    if which == "mixed":
        opm = {"s": O.ProjEntry("simplex", {"z": 1.0}, np.arange(0, n, 3)), "e": O.ProjEntry("simplex_eq", {"z": 2.5}, np.arange(1, n, 3)),
               "b": O.ProjEntry("box", {"lower": 0.0, "upper": 0.7}, np.arange(2, n, 3))}
    else:
        opm = {"s": O.ProjEntry("simplex", {"z": 1.0 if which == "simplex" else 3.0}, np.arange(n))}
    ref = O.matching_calculate(p["ccol"], p["row"], p["a"], p["c"], m, opm, p["lam"], gamma, p["b"])
    x = r1.primal_var.cpu().numpy()
    assert np.array_equal(x, ref.primal_var), f"x differs from the oracle at {np.flatnonzero(x != ref.primal_var)[:5]}"
    d = r1.projection_diag.cpu().numpy()
    first = p["ccol"][:-1][np.diff(p["ccol"]) > 0]
    is_sx = ref.branch[np.diff(p["ccol"]) > 0] >= 0
    br = (d[first] & 3)[is_sx]
    assert np.array_equal(br, ref.branch[np.diff(p["ccol"]) > 0][is_sx])
    counts = np.bincount(br, minlength=3)
    assert counts[2] > 0 and (counts[1] > 0 or gamma > 1.0) and (counts[0] > 0 or gamma < 1.0), f"branches not exercised: {counts}"


def test_movielens_shaped_ascent_agrees_on_both_paths(monkeypatch):
    """40 accelerated iterations on ratings-like data (a = 1 or log-normal, c = -rating) through the warp-per-column path and
    through the slab path.  Per evaluation x is bit-identical (previous test); the gradient sums are formed with different
    partitions (and fixed-point scales), so the iterates agree to rounding, not to the bit."""
    p = _problem(3)
    n = p["ccol"].size - 1
    kw = dict(max_iter=40, gamma=0.1, initial_step_size=1e-3, max_step_size=0.1, iteration_callback=no_iteration_callback)
    monkeypatch.setenv("DUALIP_MID", "0")
    o0 = _objective(p, _maps(n)["simplex"], 0.1)
    a = AcceleratedGradientDescent(**kw).maximize(o0, torch.zeros(int(p["n_rows"]), device=DEV))
    monkeypatch.delenv("DUALIP_MID")
    o1 = _objective(p, _maps(n)["simplex"], 0.1)
    bb = AcceleratedGradientDescent(**kw).maximize(o1, torch.zeros(int(p["n_rows"]), device=DEV))
    assert np.allclose(a.dual_objective_log, bb.dual_objective_log, rtol=1e-5)
    assert np.allclose(a.step_size_log, bb.step_size_log, rtol=1e-3)
    assert torch.allclose(a.dual_val, bb.dual_val, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("which", ["simplex", "mixed"])
@pytest.mark.parametrize("gamma", [0.1, 5.0])
def test_long_columns_cta_per_column_kernel_against_the_oracle(monkeypatch, which, gamma):
    """Columns beyond 1024 entries: a CTA per column with u in a shared-memory stash (csrc/long_col.cuh) up to 12288 entries,
    the warp-per-column kernel above.  x and the projection branch must equal the numpy restatement of the reference; with
    DUALIP_LONG_CTA=0 every long column takes the older kernel, whose x must agree to rounding."""
    rng = np.random.default_rng(5)
    m = 14000
    deg = np.concatenate([rng.integers(1025, 9000, 24), [12288, 12289, 13000, 1025, 1026, 4096, 4097], rng.integers(2, 60, 30)]).astype(np.int64)
    n = deg.size
    ccol = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=ccol[1:])
    row = np.concatenate([np.sort(rng.choice(m, size=d, replace=False)) for d in deg]).astype(np.int64)
    E = row.size
    c = (-rng.choice(np.arange(0.5, 5.01, 0.5), size=E)).astype(np.float32)
    a = np.where(rng.random(E) < 0.5, 1.0, rng.lognormal(0, 0.5, E)).astype(np.float32)
    p = dict(ccol=ccol, row=row, a=a, c=c, b=np.full(m, 0.3, dtype=np.float32), lam=(rng.random(m) * 4.0).astype(np.float32), n_rows=m)
    lam = torch.from_numpy(p["lam"]).to(DEV)
    obj = _objective(p, _maps(n)[which], gamma)
    info = obj.plan_info()
    assert info["n_long_cols"] == 31 and info["launches_per_calc"] == 3  # CTA-per-column kernel, warp-per-column kernel, slab kernel
    r = obj.calculate(lam, save_primal=True, diagnostics=True)
    if which == "mixed":
        opm = {"s": O.ProjEntry("simplex", {"z": 1.0}, np.arange(0, n, 3)), "e": O.ProjEntry("simplex_eq", {"z": 2.5}, np.arange(1, n, 3)),
               "b": O.ProjEntry("box", {"lower": 0.0, "upper": 0.7}, np.arange(2, n, 3))}
    else:
        opm = {"s": O.ProjEntry("simplex", {"z": 1.0}, np.arange(n))}
    ref = O.matching_calculate(p["ccol"], p["row"], p["a"], p["c"], m, opm, p["lam"], gamma, p["b"])
    x = r.primal_var.cpu().numpy()
    bad = np.flatnonzero(x != ref.primal_var)
    assert bad.size == 0, f"x differs from the oracle at {bad[:5]} (columns {np.searchsorted(ccol, bad[:5], side='right') - 1})"
    d = r.projection_diag.cpu().numpy()
    is_sx = ref.branch >= 0
    assert np.array_equal((d[ccol[:-1]] & 3)[is_sx], ref.branch[is_sx])
    assert abs(float(r.scalars64[0]) - ref.dual_objective) <= 1e-5 * abs(ref.dual_objective)
    assert np.allclose(r.dual_gradient.cpu().numpy(), ref.dual_gradient, rtol=1e-5, atol=1e-4)
    monkeypatch.setenv("DUALIP_LONG_CTA", "0")
    old = _objective(p, _maps(n)[which], gamma)
    assert old.plan_info()["launches_per_calc"] == 2
    r_old = old.calculate(lam, save_primal=True)
    assert torch.allclose(r_old.primal_var, r.primal_var, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("which", ["simplex", "mixed"])
@pytest.mark.parametrize("gamma", [0.1, 5.0])
def test_team_per_column_kernel_equals_the_in_kernel_path(monkeypatch, which, gamma):
    """Plans with MANY mid columns give them their own launch (32 threads per column, u in a shared-memory stash, global
    accumulators: matching_long_cta_kernel<ACC, 32>) instead of the slab kernel's warp-per-column path; DUALIP_MID_KERNEL
    forces either.  Same decisions, same x, same branch / support size; sums agree to rounding."""
    p = _problem(17)
    n = p["ccol"].size - 1
    lam = torch.from_numpy(p["lam"]).to(DEV)
    monkeypatch.setenv("DUALIP_MID_KERNEL", "0")
    inner = _objective(p, _maps(n)[which], gamma)
    monkeypatch.setenv("DUALIP_MID_KERNEL", "1")
    team = _objective(p, _maps(n)[which], gamma)
    assert inner.plan_info()["launches_per_calc"] == 2 and team.plan_info()["launches_per_calc"] == 3
    assert inner.plan_info()["n_mid_cols"] == team.plan_info()["n_mid_cols"] > 2000
    r0 = inner.calculate(lam, save_primal=True, diagnostics=True)
    r1 = team.calculate(lam, save_primal=True, diagnostics=True)
    assert torch.equal(r0.primal_var, r1.primal_var) and torch.equal(r0.projection_diag, r1.projection_diag)
    assert torch.allclose(r0.dual_gradient, r1.dual_gradient, rtol=2e-5, atol=2e-5)
    assert abs(float(r0.scalars64[0]) - float(r1.scalars64[0])) <= 1e-6 * abs(float(r0.scalars64[0]))
    r2 = team.calculate(lam)
    assert torch.allclose(r2.dual_gradient, r1.dual_gradient, rtol=1e-6, atol=1e-6)
    out = AcceleratedGradientDescent(max_iter=20, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1,
                                     iteration_callback=no_iteration_callback).maximize(team, torch.zeros(int(p["n_rows"]), device=DEV))
    assert all(np.isfinite(out.dual_objective_log))
