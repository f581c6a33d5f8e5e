"""-m gpu: dual vectors too long for the register path's shared-memory layout.

The plan picks its layout from m (dualip_plan_create): up to ~27 k duals lambda and the fixed-point accumulator live in shared
memory (mode 0, every other GPU test); up to ~48 k only the accumulator does (mode 1); beyond that neither (mode 2, global
atomics), and past 65 536 rows the row ids are 32-bit.  These plans run the generic streaming path.  Same bars as
tests/test_gpu_parity.py: primal x within 1e-5 relative with identical support, projection branch and support size bit-exact
against the C restatement of the reference (matching.py:116-188, simplex.py:126-236), objective and gradient within 1e-5; and
the three forms of the Maximizer's iteration (all-CTA tail, last-CTA tail, two launches) must agree on such plans too."""
import numpy as np
import pytest
import torch

from conftest import random_problem
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.projections import create_projection_map
from oracle import c_oracle
from oracle import dualip_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (m, shared-memory mode, row id bits)
LAYOUTS = [(40_000, 1, 16), (60_000, 2, 16), (70_000, 2, 32)]


def _problem(m):
    p = random_problem(seed=m // 1000, n_cols=2500, n_rows=m, mean_deg=9.0, scale_c=10.0, lam_scale=1.0, max_deg=40,
                       long_cols=[(5, 300), (77, 33)])
    n = p["n_cols"]
    idx = np.arange(n)
    groups = [("simplex", {"z": 1.0}, idx[idx % 4 == 0]), ("box", {"lower": 0.0, "upper": 1.0}, idx[idx % 4 == 1]),
              ("simplex", {"z": 2.5}, idx[idx % 4 == 2])]  # idx % 4 == 3: no entry (left unprojected)
    deg = np.diff(p["ccol"])
    pm, classes = {}, [c_oracle.make_class("identity", {})]
    col_class = np.zeros(n, dtype=np.uint8)
    for k, (ptype, params, cols) in enumerate(groups):
        pm.update(create_projection_map(ptype, params, n, indices=cols.tolist(), key_prefix=f"g{k}_"))
        unpadded = ptype == "simplex" and bool((deg[cols] == 1).any() and not (deg[cols] == 2).any())
        classes.append(c_oracle.make_class(ptype, params, d1_unpadded=unpadded))
        col_class[cols] = k + 1
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]), size=(m, n)).to(DEV)
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]), size=(m, n)).to(DEV)
    return p, pm, classes, col_class, A, C


def _objective(p, pm, A, C, gamma):
    return MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(p["b"]).to(DEV)), gamma=gamma)


@pytest.mark.parametrize("m,mode,row_bits", LAYOUTS)
def test_evaluation_against_the_c_oracle(m, mode, row_bits):
    gamma = 5e-2
    p, pm, classes, col_class, A, C = _problem(m)
    obj = _objective(p, pm, A, C, gamma)
    info = obj.plan_info()
    assert (info["smem_mode"], info["row_bits"]) == (mode, row_bits), info
    deg = np.diff(p["ccol"])
    assert info["n_slab_cols"] + info["n_mid_cols"] + info["n_long_cols"] == int((deg > 0).sum())
    ref = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, classes, p["lam"], gamma, p["b"], col_class)
    for lam in (torch.from_numpy(p["lam"]).to(DEV), torch.from_numpy(p["lam"])):  # device pointers, then host buffers
        r = obj.calculate(lam, save_primal=lam.is_cuda, diagnostics=lam.is_cuda)
        got = r.scalars64.cpu().numpy() if hasattr(r, "scalars64") and r.scalars64 is not None else None
        obj_val = float(got[0]) if got is not None else float(r.dual_objective)
        assert abs(obj_val - ref["scal"][0]) <= 1e-5 * abs(ref["scal"][0])
        assert np.allclose(r.dual_gradient.cpu().numpy(), ref["grad"], rtol=1e-5, atol=1e-5 * np.abs(ref["grad"]).max())
        if not lam.is_cuda:
            continue
        x = r.primal_var.cpu().numpy()
        rel = np.abs(x - ref["x"]) / np.maximum(np.abs(ref["x"]), 1e-6)
        assert rel.max() <= 1e-5, f"primal x: max relative difference {rel.max()}"
        assert int(((x != 0) != (ref["x"] != 0)).sum()) == 0, "support of x differs"
        diag = r.projection_diag.cpu().numpy()[p["ccol"][:-1][deg > 0]]
        cd = ref["diag"][deg > 0]
        is_sx = cd != 255
        assert np.array_equal((diag & 3)[is_sx], cd[is_sx] & 3), "projection branch selection differs"
        dsel = is_sx & ((cd & 3) > 0)
        assert np.array_equal((diag >> 2)[dsel], cd[dsel] >> 2), "support size differs"


@pytest.mark.parametrize("m,mode,row_bits", LAYOUTS)
def test_the_three_forms_of_the_iteration_agree(m, mode, row_bits, monkeypatch):
    """20 iterations from zero at a step size the problem is stable at (the first 14 take the initial step size,
    agd_utils.py:56-57; the last six the Lipschitz estimate): all-CTA tail (default from m = 16384), last-CTA tail, evaluation
    and update as two launches, and the oracle's loop over the C restatement.  The gradient of these plans is summed with fp32
    atomics, hence bars of a few ulps instead of equality."""
    gamma = 5e-2
    monkeypatch.setenv("DUALIP_REBALANCE", "0")
    p, pm, classes, col_class, A, C = _problem(m)
    kw = dict(max_iter=20, gamma=gamma, initial_step_size=1e-4, max_step_size=0.1, iteration_callback=no_iteration_callback)
    outs = {}
    for tag, env in (("grid", {}), ("last", {"DUALIP_GRID_TAIL": "0"}), ("two", {"DUALIP_ONE_LAUNCH": "0"})):
        for k in ("DUALIP_GRID_TAIL", "DUALIP_ONE_LAUNCH"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        obj = _objective(p, pm, A, C, gamma)
        assert obj.plan_info()["grid_tail"] == (0 if tag == "last" else 1)
        outs[tag] = AcceleratedGradientDescent(**kw).maximize(obj, torch.zeros(m, device=DEV))
        assert obj.plan_info()["grid_barrier_status"] == 0
    ref = outs["two"]
    assert len(ref.dual_objective_log) == 20 and all(np.isfinite(ref.dual_objective_log))
    assert all(b > a for a, b in zip(ref.dual_objective_log, ref.dual_objective_log[1:])), "the dual objective must ascend"
    scale = float(ref.dual_val.abs().max())
    for tag in ("grid", "last"):
        o = outs[tag]
        assert np.allclose(o.dual_objective_log, ref.dual_objective_log, rtol=1e-5)
        assert np.allclose(o.step_size_log[:14], ref.step_size_log[:14], rtol=1e-12)
        assert np.allclose(o.step_size_log, ref.step_size_log, rtol=1e-3)
        assert torch.allclose(o.dual_val, ref.dual_val, rtol=1e-4, atol=1e-5 * scale)
        assert torch.allclose(o.objective_result.dual_gradient, ref.objective_result.dual_gradient, rtol=1e-4,
                              atol=1e-5 * float(ref.objective_result.dual_gradient.abs().max()))

    def calc(lam, g):
        r = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, classes, lam, g, p["b"], col_class, want_x=False,
                               want_diag=False)
        return r["grad"], r["scal"][0]

    _, obj_log, step_log, _ = O.agd_maximize(calc, np.zeros(m, dtype=np.float32), 20, gamma, 1e-4, 0.1)
    assert np.allclose(outs["grid"].dual_objective_log, obj_log, rtol=1e-4)
    assert np.allclose(outs["grid"].step_size_log, step_log, rtol=1e-2)
