"""not gpu: the numpy restatement of the generic-LP objective (oracle/dualip_oracle.py: lp_calculate) against outputs of
the reference's MIPLIB2017ObjectiveFunction (tests/golden/lp_*.npz, written by tests/golden/make_golden_lp.py)."""
import os

import numpy as np
import pytest

from oracle import dualip_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["lp_miplib", "lp_eq_cone"])
def test_lp_oracle_matches_reference_outputs(name):
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    A, c, b = d["A"], d["c"], d["b"]
    gamma, jacobi = float(d["gamma"]), bool(d["jacobi"])
    row_norms = None
    if jacobi:
        row_norms = np.linalg.norm(A.astype(np.float32), axis=1).astype(np.float32)
        row_norms[row_norms == 0] = 1.0
    for k, lam in enumerate(d["lams"]):
        grad, obj, reg, x, primal = O.lp_calculate(A, c, b, d["lower"], d["upper"], lam, gamma, row_norms)
        assert np.allclose(x, d[f"x{k}"], rtol=1e-5, atol=1e-5)
        assert np.array_equal(x == d["lower"], d[f"x{k}"] == d["lower"]) and np.array_equal(x == d["upper"], d[f"x{k}"] == d["upper"])
        scale = max(1.0, float(np.abs(d[f"grad{k}"]).max()))
        assert np.allclose(grad, d[f"grad{k}"], rtol=1e-5, atol=1e-5 * scale)
        ref_obj, ref_reg, ref_primal = d[f"scal{k}"]
        assert abs(obj - ref_obj) <= 1e-5 * max(1.0, abs(ref_obj))
        assert abs(reg - ref_reg) <= 1e-5 * max(1.0, abs(ref_reg)) and abs(primal - ref_primal) <= 1e-5 * max(1.0, abs(ref_primal))


@pytest.mark.parametrize("name", ["lp_miplib", "lp_eq_cone"])
def test_lp_oracle_ascent_trace_matches_reference_run(name):
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    A, c, b = d["A"], d["c"], d["b"]
    jacobi = bool(d["jacobi"])
    row_norms = None
    if jacobi:
        row_norms = np.linalg.norm(A.astype(np.float32), axis=1).astype(np.float32)
        row_norms[row_norms == 0] = 1.0
    eq = d["eq_mask"] if d["eq_mask"].size else None
    steps, factor = int(d["decay"][0]), float(d["decay"][1])

    def calc(lam, gamma):
        grad, obj, *_ = O.lp_calculate(A, c, b, d["lower"], d["upper"], lam, gamma, row_norms)
        return grad, obj

    y, obj_log, step_log, _ = O.agd_maximize(calc, np.zeros(b.size, dtype=np.float32), int(d["iters"]), float(d["gamma"]), 1e-3, 0.1,
                                             gamma_decay_type="step" if steps else None,
                                             gamma_decay_params={"decay_steps": steps, "decay_factor": factor}, equality_mask=eq)
    ref, got = d["obj_log"], np.array(obj_log)
    # the ascent on these LPs amplifies rounding differences (clamp pattern switches, Lipschitz step from differences of
    # gradients): the first iterations agree to fp32 accuracy, the full trace to a few per cent of its range
    assert np.allclose(got[:12], ref[:12], rtol=1e-4, atol=1e-4 * np.abs(ref[:12]).max())
    assert np.abs(got - ref).max() <= 5e-2 * np.abs(ref).max()


def test_reference_known_answer_equality_constrained_lp():
    """Reference tests/test_equality_constraints.py:19-66: minimise x1 + 2 x2 s.t. x1 + x2 = 4, 0 <= x1 <= 1 (x2 unprojected),
    gamma 1e-5, 1000 AGD iterations with the default step sizes, free multiplier on the equality row -> objective 7.0 within
    torch.isclose(atol=1e-5) (i.e. 1e-5 + 1e-5*7).  The reference itself, run in the build container, ends at
    7.000050067901611 with lambda = -2.  Also project_on_nn_cone with an equality mask (:8-16)."""
    from oracle import dualip_oracle as O

    y = np.array([-1.0, -1.0, 2.0, -3.0, 4.0], dtype=np.float32)
    mask = np.array([False, True, False, True, False])
    assert np.array_equal(O.project_on_nn_cone(y, mask), np.array([0.0, -1.0, 2.0, -3.0, 4.0], dtype=np.float32))

    A = np.array([[1.0, 1.0]], dtype=np.float32)
    c = np.array([1.0, 2.0], dtype=np.float32)
    b = np.array([4.0], dtype=np.float32)
    lower = np.array([0.0, -np.inf], dtype=np.float32)  # box(upper=1) on x1 has lower = 0 (projections/box.py:6-13)
    upper = np.array([1.0, np.inf], dtype=np.float32)

    def calc(lam, gamma):
        grad, dual_obj, _, _, _ = O.lp_calculate(A, c, b, lower, upper, lam, gamma)
        return grad, np.float32(dual_obj)

    y, obj_log, _, _ = O.agd_maximize(calc, np.zeros(1, np.float32), 1000, 1e-5, equality_mask=np.array([True]))
    assert abs(obj_log[-1] - 7.0) <= 1e-5 + 1e-5 * 7.0
    assert abs(obj_log[-1] - 7.000050067901611) < 1e-6 and abs(float(y[0]) + 2.0) < 1e-4 and obj_log[0] == -200000.0
