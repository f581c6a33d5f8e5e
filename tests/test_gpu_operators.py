"""-m gpu: the public CSC operators (reference src/dualip/utils/sparse_utils.py) and the fairness-row objective of the
reference's extension demo (docs/demo/matching_complex.rst:82-168) on CUDA tensors.

Operator cases mirror the reference's own tests/test_sparse_utils.py:95-223 (same matrices and callables, tensors on the
GPU); ops_reference.npz / fair_*.npz hold outputs of the reference's operators and of the demo recipe executed with them
(tests/golden/make_golden_fair.py)."""
from operator import add, mul

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from dualip_b200.objectives.matching import MatchingInputArgs, calc_grad
from dualip_b200.objectives.matching_fairness import MatchingFairnessDualObjectiveFunction, build_fairness_constraints
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.projections import create_projection_map, project
from dualip_b200.utils.sparse_utils import (
    apply_F_to_columns,
    elementwise_csc,
    hstack_csc,
    left_multiply_sparse,
    right_multiply_sparse,
    row_norms_csc,
    row_sums_csc,
    split_csc_by_cols,
    vstack_csc,
)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _sp(dense):
    return dense.to_sparse_csc().to(DEV)


# ---- the reference's operator tests, on the device ----
def test_stacking_like_the_reference_tests():
    A, B = torch.tensor([[1.0, 0.0, 2.0], [0.0, 3.0, 0.0]]), torch.tensor([[4.0, 5.0, 0.0], [0.0, 0.0, 6.0]])
    v = vstack_csc([_sp(A), _sp(B)])
    assert v.layout == torch.sparse_csc and torch.equal(v.to_dense().cpu(), torch.vstack([A, B]))
    A2, B2 = torch.tensor([[1.0, 2.0], [3.0, 0.0]]), torch.tensor([[0.0, 4.0, 5.0], [6.0, 0.0, 7.0]])
    h = hstack_csc([_sp(A2), _sp(B2)])
    assert h.layout == torch.sparse_csc and torch.equal(h.to_dense().cpu(), torch.hstack([A2, B2]))
    parts = split_csc_by_cols(h, [2, 3])
    assert torch.equal(parts[0].to_dense().cpu(), A2) and torch.equal(parts[1].to_dense().cpu(), B2)


def test_left_and_right_multiply_like_the_reference_tests():
    M = torch.tensor([[1.0, 0.0, 3.0], [0.0, 2.0, 0.0], [4.0, 0.0, 5.0]])
    v = torch.tensor([2.0, 3.0, 0.5])
    r = left_multiply_sparse(v.to(DEV), _sp(M))
    assert r.layout == torch.sparse_csc and torch.allclose(r.to_dense().cpu(), torch.diag(v) @ M)
    out = _sp(M).clone()
    left_multiply_sparse(v.to(DEV), _sp(M), output_tensor=out)
    assert torch.allclose(out.to_dense().cpu(), torch.diag(v) @ M)
    r = right_multiply_sparse(_sp(M), v.to(DEV))
    assert torch.allclose(r.to_dense().cpu(), M @ torch.diag(v))


@pytest.mark.parametrize("fn,expect", [
    (lambda x: x, lambda M: M),
    (lambda x: 2 * x, lambda M: 2 * M),
    (lambda x: x * 0.5, lambda M: M * 0.5),
    (lambda x: x.clamp(min=0), lambda M: M.clamp(min=0)),
    (lambda x: -x, lambda M: -M),
])
def test_apply_F_to_columns_like_the_reference_tests(fn, expect):
    M = torch.tensor([[1.0, 0.0, -3.0, 0.0, 7.0], [0.0, -2.0, 0.0, 4.0, 0.0], [5.0, 0.0, 6.0, 0.0, 8.0]])
    single = apply_F_to_columns(_sp(M), fn, [torch.arange(5)])
    multi = apply_F_to_columns(_sp(M), fn, [torch.tensor([0, 2, 4]), torch.tensor([], dtype=torch.long), torch.tensor([1, 3])])
    assert torch.allclose(single.to_dense().cpu(), expect(M)) and torch.allclose(multi.to_dense().cpu(), expect(M))
    out = _sp(M).clone()
    apply_F_to_columns(_sp(M), fn, [torch.arange(5)], output_tensor=out)
    assert torch.allclose(out.to_dense().cpu(), expect(M))


def test_apply_F_varying_column_lengths_sees_zero_padded_blocks():
    M = torch.tensor([[1.0, 0.0, 3.0], [2.0, 0.0, 0.0], [3.0, 4.0, 0.0], [4.0, 0.0, 0.0]])
    seen = []

    def f(block):
        seen.append(block.clone())
        return block**2

    r = apply_F_to_columns(_sp(M), f, [torch.arange(3)])
    assert torch.allclose(r.to_dense().cpu(), M**2)
    assert seen[0].shape == (4, 3) and torch.equal(seen[0].cpu(), torch.tensor([[1.0, 4.0, 3.0], [2.0, 0, 0], [3.0, 0, 0], [4.0, 0, 0]]))


def test_operators_against_outputs_of_the_reference_operators():
    d = np.load(f"{GOLDEN}/ops_reference.npz")
    m, n = int(d["n_rows"]), d["ccol"].size - 1
    for idx in (torch.int64, torch.int32):
        M = torch.sparse_csc_tensor(torch.from_numpy(d["ccol"]).to(idx), torch.from_numpy(d["row"]).to(idx), torch.from_numpy(d["vals"]),
                                    size=(m, n)).to(DEV)
        lm = left_multiply_sparse(torch.from_numpy(d["v"]).to(DEV), M)
        assert np.array_equal(lm.values().cpu().numpy(), d["left_multiply"])  # one rounding per entry: bit-identical
        assert np.allclose(row_sums_csc(M).cpu().numpy(), d["row_sums"], rtol=1e-5, atol=1e-6)
        assert np.allclose(row_norms_csc(M).cpu().numpy() ** 2, np.bincount(d["row"], weights=d["vals"].astype(np.float64) ** 2, minlength=m),
                           rtol=1e-5)
        ap = apply_F_to_columns(M, project("simplex", z=0.05), [torch.arange(n)])
        assert np.array_equal(ap.values().cpu().numpy(), d["apply_simplex_all"])  # the package's simplex operator == the reference's
        even = torch.from_numpy(d["even_cols"])
        aff = apply_F_to_columns(M, lambda blk: blk * 2 + (blk != 0) * 1.0, [even]).values().cpu().numpy()
        pos = np.concatenate([np.arange(d["ccol"][j], d["ccol"][j + 1]) for j in d["even_cols"]])
        assert np.array_equal(aff[pos], d["apply_affine_even"][pos])
        rest = np.setdiff1d(np.arange(d["vals"].size), pos)
        assert np.array_equal(aff[rest], d["vals"][rest]), "columns outside every bucket keep M's values"


def test_elementwise_and_calc_grad():
    A = torch.tensor([[1.0, 0.0], [2.0, 3.0]])
    B = torch.tensor([[5.0, 0.0], [7.0, 11.0]])
    assert torch.equal(elementwise_csc(_sp(A), _sp(B), add).to_dense().cpu(), A + B)
    assert torch.equal(elementwise_csc(_sp(A), _sp(B), mul).to_dense().cpu(), A * B)
    with pytest.raises(ValueError):
        elementwise_csc(_sp(A), _sp(torch.tensor([[1.0, 1.0], [0.0, 1.0]])), add)
    with pytest.raises(ValueError):
        elementwise_csc(A.to(DEV), _sp(B), add)
    with pytest.raises(RuntimeError):
        row_sums_csc(A.to_sparse_csc())  # CPU tensor: no fallback
    g, o = calc_grad(torch.tensor([1.0, 2.0]), torch.tensor(3.0), torch.tensor([0.5, 0.25]), torch.tensor([0.5, 0.5]), torch.tensor(0.125))
    assert torch.equal(g, torch.tensor([0.5, 1.5])) and float(o) == 3.0 + 0.125 + 0.25 + 0.375


# ---- fairness rows ----
def _fair_case(name):
    d = np.load(f"{GOLDEN}/fair_{name}.npz")
    params = {str(k): float(v) for k, v in zip(d["proj_keys"], d["proj_vals"])}
    return d, str(d["proj_type"]), params


def _fair_objective(d, ptype, params, batching, index_dtype=torch.int64, explicit=False):
    m, n = int(d["n_rows"]), d["ccol"].size - 1
    ccol, row = torch.from_numpy(d["ccol"]).to(index_dtype), torch.from_numpy(d["row"]).to(index_dtype)
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(d["a"]), size=(m, n)).to(DEV)
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(d["c"]), size=(m, n)).to(DEV)
    args = MatchingInputArgs(A, C, create_projection_map(ptype, params, n), torch.from_numpy(d["b"]).to(DEV))
    if explicit:
        F = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(d["f"]), size=(m, n)).to(DEV)
        return MatchingFairnessDualObjectiveFunction(args, float(d["gamma"]), batching=batching, A_fairness=F)
    return MatchingFairnessDualObjectiveFunction(args, float(d["gamma"]), batching=batching, group_ratio=float(d["group_ratio"]))


@pytest.mark.parametrize("name", ["simplex", "box", "simplex_eq"])
@pytest.mark.parametrize("batching", [True, False])
def test_fairness_objective_against_the_demo_recipe_run_with_the_reference_operators(name, batching):
    d, ptype, params = _fair_case(name)
    tag = "b1" if batching else "b0"
    for idx, explicit in ((torch.int64, False), (torch.int32, True)):
        obj = _fair_objective(d, ptype, params, batching, idx, explicit)
        assert np.array_equal(obj._f.cpu().numpy(), d["f"]), "A_fairness values (rst:46-64)"
        r = obj.calculate(torch.from_numpy(d["lam"]).to(DEV), save_primal=True)
        assert np.array_equal(r.primal_var.cpu().numpy(), d[f"x_{tag}"]), "primal x differs from the demo recipe on the reference"
        scal, got = d[f"scal_{tag}"], r.scalars64.cpu().numpy()  # dual_obj, primal_obj, reg, lam.grad, max_pos, sum_pos
        assert abs(got[0] - scal[0]) <= 1e-5 * abs(scal[0])
        assert abs(got[1] - scal[1]) <= 1e-5 * abs(scal[1])
        # reg = gamma/2 * torch.norm(x)**2: the reference's fp32 norm is itself only good to a few 1e-5 (22k terms); the kernel
        # sums in fp64, so it is compared with the exact value of the (bit-identical) x and the reference with a looser bar
        reg64 = 0.5 * float(d["gamma"]) * float(np.sum(d[f"x_{tag}"].astype(np.float64) ** 2))
        assert abs(got[2] - reg64) <= 2e-7 * reg64 and abs(scal[2] - reg64) <= 5e-5 * reg64
        assert abs(got[3] - scal[3]) <= 1e-5 * abs(scal[3]) + 1e-4
        assert abs(got[4] - scal[4]) <= 1e-5 * max(1.0, abs(scal[4])) and abs(got[5] - scal[5]) <= 1e-5 * max(1.0, abs(scal[5]))
        g, gref = r.dual_gradient.cpu().numpy(), d[f"grad_{tag}"]
        assert g.shape == gref.shape == (int(d["n_rows"]) + 2,)
        assert np.allclose(g, gref, rtol=1e-5, atol=1e-5)
        assert np.isclose(g[-2] + d["b"][-2], -(g[-1] + d["b"][-1]), rtol=1e-6, atol=1e-7), "the fairness rows are negatives of each other"


def test_demo_recipe_composed_from_this_package_operators_matches_the_fused_fairness_kernel():
    """The rst:86-167 `calculate`, written with dualip_b200's own operators on CUDA tensors, against dualip_fair_calc."""
    d, ptype, params = _fair_case("simplex")
    m, n, gamma = int(d["n_rows"]), d["ccol"].size - 1, float(d["gamma"])
    obj = _fair_objective(d, ptype, params, True)
    lam = torch.from_numpy(d["lam"]).to(DEV)
    A, Fm, C = obj.A, obj.A_fairness, obj.c
    inter = torch.sparse_csc_tensor(A.ccol_indices(), A.row_indices(), torch.zeros_like(A.values()), size=A.size())
    c_rescaled = -1.0 / gamma * C
    scaled = -1.0 / gamma * lam
    left_multiply_sparse(scaled[:-2], A, output_tensor=inter)
    elementwise_csc(inter, scaled[-2] * Fm, add, output_tensor=inter)
    elementwise_csc(inter, -1 * scaled[-1] * Fm, add, output_tensor=inter)
    elementwise_csc(inter, c_rescaled, add, output_tensor=inter)
    apply_F_to_columns(inter, project(ptype, **params), [torch.arange(n)], output_tensor=inter)
    grad = torch.zeros_like(lam)
    grad[:-2] = row_sums_csc(elementwise_csc(A, inter, mul))
    grad[-2] = elementwise_csc(Fm, inter, mul).values().sum()
    grad[-1] = elementwise_csc(-1 * Fm, inter, mul).values().sum()
    vals = inter.values()
    reg = (gamma / 2) * torch.norm(vals) ** 2
    grad, dual_obj = calc_grad(grad, torch.dot(C.values(), vals), lam, obj.b_vec, reg)
    r = obj.calculate(lam, save_primal=True)
    assert np.array_equal(vals.cpu().numpy(), d["x_b0"]) and torch.equal(vals, r.primal_var)  # one bucket = batching False
    assert torch.allclose(grad, r.dual_gradient, rtol=1e-5, atol=1e-5)
    assert abs(float(dual_obj) - float(r.scalars64[0])) <= 1e-5 * abs(float(dual_obj))


def test_fairness_objective_under_the_maximizer():
    """The same Maximizer drives it (device loop: dualip_fair_calc + dualip_agd_step): the dual objective ascends and the
    fairness gap |row m| shrinks against the unconstrained start."""
    d, ptype, params = _fair_case("simplex")
    obj = _fair_objective(d, ptype, params, True)
    m = int(d["n_rows"])
    lam0 = torch.zeros(m + 2, device=DEV)
    r0 = obj.calculate(lam0)
    solver = AcceleratedGradientDescent(max_iter=300, gamma=float(d["gamma"]), initial_step_size=1e-3, max_step_size=0.1,
                                        iteration_callback=no_iteration_callback)
    out = solver.maximize(obj, lam0)
    log = out.dual_objective_log
    assert len(log) == 300 and log[-1] > log[0] and all(np.isfinite(log))
    viol0 = float(torch.relu(r0.dual_gradient[-2:]).max())
    viol1 = float(torch.relu(out.objective_result.dual_gradient[-2:]).max())
    assert viol1 <= viol0 + 1e-6
    assert out.dual_val.shape == (m + 2,) and bool((out.dual_val >= 0).all())
