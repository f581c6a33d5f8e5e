"""Step-size rule of the Maximizer for host-driven objectives: same function names and semantics as the reference
(src/dualip/optimizers/agd_utils.py:4-89).  The CUDA matching objectives do not go through here: their loop uses the
device-resident ring in csrc/agd.cu (dualip_agd_step)."""
import math

import torch


def norm_of_difference(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return torch.linalg.vector_norm(x - y)


def update_dual_gradient_history(gradient, dual_val, grad_history: list, dual_history: list, max_history_length: int) -> None:
    """Bounded FIFO of (gradient, dual) snapshots; both lists stay aligned (reference agd_utils.py:13-27)."""
    while len(grad_history) >= max_history_length:
        del grad_history[0]
        del dual_history[0]
    grad_history.append(gradient.detach().clone())
    dual_history.append(dual_val.detach().clone())


def estimate_lipschitz_constant(grad_one, grad_two, dual_one, dual_two) -> torch.Tensor:
    """||g2 - g1|| / ||d2 - d1|| (reference agd_utils.py:30-41)."""
    return norm_of_difference(grad_one, grad_two) / norm_of_difference(dual_one, dual_two)


def step_size_from_lipschitz_constants(lipschitz_constants: list, max_history_length: int, initial_step_size: float,
                                       max_step_size: float) -> float:
    """1/max(L) clamped to max_step_size once max_history_length-1 estimates exist; initial_step_size before that or
    when the maximum is NaN/Inf (reference agd_utils.py:44-62).  The maximum follows Python's max(): the first element
    stays unless a later one compares greater."""
    if len(lipschitz_constants) < max_history_length - 1 or not lipschitz_constants:
        return initial_step_size
    values = [float(v) for v in lipschitz_constants]
    top = values[0]
    for v in values[1:]:
        if v > top:
            top = v
    if math.isnan(top) or math.isinf(top):
        return initial_step_size
    candidate = 1.0 / top if top != 0 else max_step_size
    return min(candidate, max_step_size)


def calculate_step_size(dual_grad, dual_val, grad_history: list, dual_history: list, max_history_length: int = 15,
                        initial_step_size: float = 1e-5, max_step_size: float = 0.1, lipschitz_cache: list = None) -> float:
    """Pushes the newest (gradient, dual) pair and returns the step (reference agd_utils.py:65-89).

    The reference recomputes every ||g_{k+1}-g_k|| / ||y_{k+1}-y_k|| of the ring on each call (:86-88) although only the
    newest pair changed.  With `lipschitz_cache` (a list owned by the caller, one entry per adjacent pair of the ring)
    the old estimates are reused: same numbers, one estimate per call instead of fourteen."""
    if lipschitz_cache is None:
        update_dual_gradient_history(dual_grad, dual_val, grad_history, dual_history, max_history_length)
        estimates = [
            estimate_lipschitz_constant(grad_history[k], grad_history[k + 1], dual_history[k], dual_history[k + 1])
            for k in range(len(grad_history) - 1)
        ]
        return step_size_from_lipschitz_constants(estimates, max_history_length, initial_step_size, max_step_size)
    before = len(grad_history)
    update_dual_gradient_history(dual_grad, dual_val, grad_history, dual_history, max_history_length)
    dropped = before + 1 - len(grad_history)  # entries evicted from the front of the ring
    del lipschitz_cache[:dropped]
    if len(grad_history) >= 2:
        lipschitz_cache.append(float(estimate_lipschitz_constant(grad_history[-2], grad_history[-1], dual_history[-2], dual_history[-1])))
    del lipschitz_cache[: max(0, len(lipschitz_cache) - (len(grad_history) - 1))]
    return step_size_from_lipschitz_constants(lipschitz_cache, max_history_length, initial_step_size, max_step_size)
