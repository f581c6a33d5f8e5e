"""-m gpu: equality rows (`MatchingInputArgs.equality_mask`) through the matching objective's fused iteration.

The Maximizer keeps the dual of an equality row free and clamps the others at zero (reference optimizers/agd.py:13-21,
:181-183; reference tests/test_equality_constraints.py exercises it on the generic-LP objective only).  Here the projection on
the dual cone happens inside the slab kernel's tail (last CTA or all CTAs), in the separate update kernel of the two-launch
form, inside a CUDA-graph replay, and in the host-side step of the host-buffer loop: all five against the oracle's loop over the
C restatement of `calculate`."""
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
GAMMA, ITERS, STEP = 5e-2, 40, 1e-4

FORMS = {
    "last_cta_tail": {"DUALIP_GRID_TAIL": "0", "DUALIP_GRAPH": "0"},
    "all_cta_tail": {"DUALIP_GRID_TAIL": "1", "DUALIP_GRAPH": "0"},
    "graph_replay": {"DUALIP_GRID_TAIL": "0", "DUALIP_GRAPH": "1", "DUALIP_GRAPH_CHUNK": "8"},
    "two_launches": {"DUALIP_ONE_LAUNCH": "0", "DUALIP_GRAPH": "0"},
    "host_buffers": {},
}


@pytest.fixture(scope="module")
def problem():
    p = random_problem(23, 4000, 96, 8.0, scale_c=10.0)
    m = p["n_rows"]
    eq = np.zeros(m, dtype=bool)
    eq[::3] = True
    classes = [c_oracle.make_class("simplex", {"z": 1.0})]
    zero = np.zeros(m, dtype=np.float32)
    ax0 = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, classes, zero, GAMMA, zero, None, want_x=False,
                             want_diag=False)["grad"]
    # even rows can never be filled (their dual wants to go negative), odd rows are tight
    p["b"] = (ax0 * np.where(np.arange(m) % 2 == 0, 1.5, 0.3)).astype(np.float32)

    def calc(lam, g):
        r = c_oracle.calculate(p["ccol"], p["row"], p["a"], p["c"], m, classes, lam, g, p["b"], None, want_x=False, want_diag=False)
        return r["grad"], r["scal"][0]

    lam, obj_log, step_log, _ = O.agd_maximize(calc, zero, ITERS, GAMMA, STEP, 0.1, equality_mask=eq)
    assert (lam[eq] < 0).sum() >= 5 and (lam[~eq] == 0).sum() >= 5 and lam[~eq].min() == 0.0  # the mask matters on this problem
    return p, eq, lam, obj_log, step_log


@pytest.mark.parametrize("form", list(FORMS))
def test_equality_rows_in_every_form_of_the_iteration(problem, form, monkeypatch):
    p, eq, lam_ref, obj_log, step_log = problem
    m, n = p["n_rows"], p["n_cols"]
    monkeypatch.setenv("DUALIP_REBALANCE", "0")  # graph replay from the first chunk on
    for k, v in FORMS[form].items():
        monkeypatch.setenv(k, v)
    ccol, row = torch.from_numpy(p["ccol"]), torch.from_numpy(p["row"])
    A = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["a"]), size=(m, n)).to(DEV)
    C = torch.sparse_csc_tensor(ccol, row, torch.from_numpy(p["c"]), size=(m, n)).to(DEV)
    args = MatchingInputArgs(A, C, create_projection_map("simplex", {"z": 1.0}, n), torch.from_numpy(p["b"]).to(DEV),
                             equality_mask=torch.from_numpy(eq).to(DEV))
    obj = MatchingSolverDualObjectiveFunction(args, gamma=GAMMA)
    start = torch.zeros(m) if form == "host_buffers" else torch.zeros(m, device=DEV)
    out = AcceleratedGradientDescent(max_iter=ITERS, gamma=GAMMA, initial_step_size=STEP, max_step_size=0.1,
                                     iteration_callback=no_iteration_callback).maximize(obj, start)
    assert out.dual_val.device == start.device
    lam = out.dual_val.cpu().numpy()
    assert (lam[eq] < -1e-3).sum() == (lam_ref[eq] < -1e-3).sum() >= 5, "equality rows must be left free"
    assert lam[~eq].min() == 0.0 and (lam[~eq] == 0).sum() >= 5, "inequality rows must be clamped at zero"
    assert np.allclose(out.dual_objective_log, obj_log, rtol=1e-5)
    assert np.allclose(out.step_size_log[:14], step_log[:14], rtol=1e-12)
    assert np.allclose(out.step_size_log, step_log, rtol=1e-2)
    assert np.allclose(lam, lam_ref, rtol=1e-3, atol=1e-4)
