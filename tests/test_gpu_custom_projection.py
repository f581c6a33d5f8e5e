"""-m gpu: user-registered projection operators (the reference's `register` / `project` seam, projections/base.py:39-57).
An operator written for the reference -- `__init__` plus `__call__` on a zero-padded [L x K] block, tensor ops only -- has no
native class; its columns go through padded blocks on the device (the reference's apply_F_to_columns scheme,
utils/sparse_utils.py:133-220) while natively projected columns stay in the fused kernel."""
import numpy as np
import pytest
import torch

from conftest import random_problem
from dualip_b200.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
from dualip_b200.optimizers.agd import AcceleratedGradientDescent, no_iteration_callback
from dualip_b200.projections import create_projection_map
from dualip_b200.projections.base import ProjectionOperator, register
from test_gpu_parity import DEV, _csc

pytestmark = pytest.mark.gpu


@register("user_capped_box")
class UserCappedBox(ProjectionOperator):
    def __init__(self, upper: float = 0.5):
        self.upper = upper

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return torch.clamp(x, min=0.0, max=self.upper)


@register("user_simplex")
class UserSimplex(ProjectionOperator):
    """{x >= 0, sum x <= z} by sorting, written with tensor ops on the padded block (zero padding stays zero)."""

    def __init__(self, z: float = 1.0):
        self.z = z

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        u = torch.clamp(x, min=0.0)
        srt, _ = torch.sort(u, dim=0, descending=True)
        css = torch.cumsum(srt.double(), dim=0)
        k = torch.arange(1, x.shape[0] + 1, device=x.device, dtype=torch.float64).unsqueeze(1)
        cond = srt.double() - (css - self.z) / k > 0
        rho = cond.to(torch.int64).cumsum(0).argmax(0)  # last index where cond holds
        theta = ((css.gather(0, rho.unsqueeze(0)) - self.z) / (rho + 1).unsqueeze(0)).squeeze(0)
        theta = torch.where(u.sum(0) <= self.z, torch.zeros_like(theta), theta)
        return torch.clamp(u - theta.float().unsqueeze(0), min=0.0)


def _objective(p, pm, gamma, batching=True):
    A, C = _csc(p)
    return MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(p["b"]).to(DEV)), gamma=gamma, batching=batching)


def _close(r_user, r_native, x_exact):
    xu, xn = r_user.primal_var, r_native.primal_var
    if x_exact:
        assert torch.equal(xu, xn)
    else:
        # a different (float64) formula for the same projection: agreement to a few float32 ulps of the largest |v| in a column,
        # which is ~|c|/gamma ~ 1e2 here
        assert float((xu - xn).abs().max()) <= 5e-5
    gs = float(r_native.dual_gradient.abs().max())
    assert torch.allclose(r_user.dual_gradient, r_native.dual_gradient, rtol=1e-5, atol=1e-5 * gs)
    su, sn = r_user.scalars64.cpu().numpy(), r_native.scalars64.cpu().numpy()
    assert np.allclose(su[:4], sn[:4], rtol=2e-6, atol=1e-6 * abs(sn[0]))


@pytest.mark.parametrize("batching", [True, False])
def test_user_operator_equals_native_operator(batching):
    p = random_problem(61, 5000, 60, 9.0, scale_c=10.0, lam_scale=0.5, long_cols=[(5, 50), (77, 33)])
    n, gamma = p["n_cols"], 3e-2
    lam = torch.from_numpy(p["lam"]).to(DEV)
    user = _objective(p, create_projection_map("user_capped_box", {"upper": 0.5}, n), gamma, batching)
    native = _objective(p, create_projection_map("box", {"lower": 0.0, "upper": 0.5}, n), gamma, batching)
    assert user.has_block_entries and not native.has_block_entries
    _close(user.calculate(lam, save_primal=True), native.calculate(lam, save_primal=True), x_exact=True)
    usx = _objective(p, create_projection_map("user_simplex", {"z": 1.0}, n), gamma, batching)
    nsx = _objective(p, create_projection_map("simplex", {"z": 1.0}, n), gamma, batching)
    _close(usx.calculate(lam, save_primal=True), nsx.calculate(lam, save_primal=True), x_exact=False)


def test_mixed_native_and_user_entries_and_the_fused_loop():
    p = random_problem(62, 6000, 80, 8.0, scale_c=10.0, lam_scale=0.5)
    n, m, gamma = p["n_cols"], p["n_rows"], 2e-2
    even, odd = np.arange(0, n, 2), np.arange(1, n, 2)
    pm_user, pm_native = {}, {}
    pm_user.update(create_projection_map("simplex", {"z": 1.0}, n, indices=even))
    pm_user.update(create_projection_map("user_capped_box", {"upper": 0.5}, n, indices=odd))
    pm_native.update(create_projection_map("simplex", {"z": 1.0}, n, indices=even))
    pm_native.update(create_projection_map("box", {"lower": 0.0, "upper": 0.5}, n, indices=odd))
    user, native = _objective(p, pm_user, gamma), _objective(p, pm_native, gamma)
    lam = torch.from_numpy(p["lam"]).to(DEV)
    _close(user.calculate(lam, save_primal=True), native.calculate(lam, save_primal=True), x_exact=True)
    # host-buffer call and the device-resident loop (with and without a per-iteration callback, step gamma decay, save_primal)
    rh = user.calculate(lam.cpu())
    assert rh.dual_gradient.device.type == "cpu"
    assert abs(float(rh.dual_objective) - float(native.calculate(lam).dual_objective)) <= 2e-6 * abs(float(rh.dual_objective))
    runs = {}
    for tag, obj, cb in (("user", user, no_iteration_callback), ("user_cb", user, lambda i, r: None), ("native", native, no_iteration_callback)):
        solver = AcceleratedGradientDescent(max_iter=30, gamma=gamma, initial_step_size=1e-3, max_step_size=0.1, save_primal=True,
                                            gamma_decay_type="step", gamma_decay_params={"decay_steps": 9, "decay_factor": 0.5},
                                            iteration_callback=cb)
        runs[tag] = solver.maximize(obj, torch.zeros(m, device=DEV))
    for tag in ("user", "user_cb"):
        assert np.allclose(runs[tag].dual_objective_log, runs["native"].dual_objective_log, rtol=1e-5)
        assert np.allclose(runs[tag].step_size_log, runs["native"].step_size_log, rtol=2e-2)
        assert torch.allclose(runs[tag].dual_val, runs["native"].dual_val, rtol=1e-3, atol=1e-4)
        # the final primal sits behind 30 iterations whose gradients were summed in different orders (index_add_ vs the
        # kernel's accumulator): the iterates agree to ~1e-4, and so does x except where a column sits on a projection boundary
        dx = (runs[tag].objective_result.primal_var - runs["native"].objective_result.primal_var).abs()
        assert float((dx > 1e-2).float().mean()) < 1e-3 and float(dx.mean()) < 1e-4
