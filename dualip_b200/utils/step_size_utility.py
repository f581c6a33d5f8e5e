"""numpy twin of optimizers/agd_utils.py: the same five names on `np.ndarray` iterates (reference
src/dualip/utils/step_size_utility.py:4-90; nothing in the reference calls it, it is kept for callers that drive the step-size
rule from numpy code).  The rule itself -- first maximum in list order, NaN/Inf -> initial step, 1/L clamped -- is the one
function shared with the tensor version."""
import numpy as np

from dualip_b200.optimizers.agd_utils import step_size_from_lipschitz_constants

__all__ = ["norm_of_difference", "update_dual_gradient_history", "estimate_lipschitz_constant",
           "step_size_from_lipschitz_constants", "calculate_step_size"]


def norm_of_difference(x: np.ndarray, y: np.ndarray) -> float:
    return np.linalg.norm(np.asarray(x) - np.asarray(y))


def update_dual_gradient_history(gradient, dual_val, grad_history: list, dual_history: list, max_history_length: int) -> None:
    """Bounded FIFO of (gradient, dual) copies; the two lists stay aligned (reference :12-29)."""
    while len(grad_history) >= max_history_length:
        del grad_history[0]
        del dual_history[0]
    grad_history.append(np.array(gradient))
    dual_history.append(np.array(dual_val))


def estimate_lipschitz_constant(grad_one, grad_two, dual_one, dual_two) -> float:
    """||g2 - g1|| / ||d2 - d1|| (reference :31-41)."""
    return norm_of_difference(grad_one, grad_two) / norm_of_difference(dual_one, dual_two)


def calculate_step_size(dual_grad, dual_val, grad_history: list, dual_history: list, max_history_length: int = 15,
                        initial_step_size: float = 1e-5, max_step_size: float = 0.1) -> float:
    """Pushes the newest pair and returns the step from every adjacent pair of the ring (reference :66-90)."""
    update_dual_gradient_history(dual_grad, dual_val, grad_history, dual_history, max_history_length)
    estimates = [estimate_lipschitz_constant(grad_history[k], grad_history[k + 1], dual_history[k], dual_history[k + 1])
                 for k in range(len(grad_history) - 1)]
    return step_size_from_lipschitz_constants(estimates, max_history_length, initial_step_size, max_step_size)
