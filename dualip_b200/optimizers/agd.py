"""Maximizer: Nesterov-accelerated projected gradient ascent on the dual.

Same class name, constructor and `maximize(f, initial_value, rank=0) -> SolverResult` as the reference
(src/dualip/optimizers/agd.py:66-229).  Three loops:

* fused (CUDA objectives, device-resident dual): iterate, history ring, step-size rule and logs live on the device
  (csrc/agd_step.cuh).  Matching objectives take evaluation AND step in ONE launch per iteration (the slab kernel's last CTA
  steps; sharded: it also exchanges the partial sums through peer memory), and once the plan has settled whole chunks of
  iterations replay as one CUDA graph (device-resident gamma / beta schedule).  Other objectives (generic LP, fairness rows,
  user-registered projections, per-iteration callbacks) launch their kernels plus one update kernel.  No host synchronisation.
* host-buffer (float32 CPU dual, native matching objective): one native call per iteration, dualip_matching_step_host
  (lambda up, fused kernel, gradient + scalars down, host-side step).
* generic (any object with `calculate` / `equality_mask`): host-driven, the reference's semantics op for op.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Callable, Optional

import torch
import torch.distributed as dist

from dualip_b200 import _native
from dualip_b200.objectives.base import BaseObjective
from dualip_b200.optimizers.agd_utils import calculate_step_size
from dualip_b200.types import ObjectiveResult, SolverResult

_IDX = {name: i for i, name in enumerate(_native.SCALAR_FIELDS)}


def project_on_nn_cone(y: torch.Tensor, equality_mask: torch.Tensor | None = None) -> torch.Tensor:
    """Dual cone: lambda_r >= 0 on inequality rows, free on equality rows (reference agd.py:13-21)."""
    clipped = y.clamp_min(0.0)
    return clipped if equality_mask is None else torch.where(equality_mask, y, clipped)


def format_objective_result_summary(iteration: int, objective_result: ObjectiveResult) -> str:
    """One-line iteration summary with the reference's field names (agd.py:24-63)."""

    def show(name, val):
        if val is None:
            return None
        try:
            if isinstance(val, torch.Tensor):
                return f"{name}={val.item()}" if val.numel() == 1 else f"{name}.shape={tuple(val.shape)}"
            return f"{name}={val}"
        except Exception:
            return f"{name}=<unprintable>"

    try:
        grad_norm = f"dual_grad_norm={float(objective_result.dual_gradient.norm().item())}"
    except Exception:
        grad_norm = "dual_grad_norm=<unprintable>"
    parts = [f"iter={iteration}", show("dual_objective", objective_result.dual_objective), grad_norm]
    for name in ("reg_penalty", "primal_objective", "primal_var", "dual_val_times_grad", "max_pos_slack", "sum_pos_slack"):
        parts.append(show(name, getattr(objective_result, name, None)))
    return " | ".join(p for p in parts if p is not None)


def gamma_schedule(gamma0: float, n_iters: int, decay_steps: int = 0, decay_factor: float = 1.0):
    """What the reference's loop does to gamma, unrolled (agd.py:102-109, :186-187): iteration i (1-based) runs with
    gammas[i-1]; after it, if i % decay_steps == 0, gamma *= decay_factor (and the step cap follows: flags[i-1] = 1).
    Returns (gammas of length n_iters + 1 -- the last entry is gamma after the final iteration -- , flags of length n_iters).
    The products are formed one after the other in Python doubles, like the reference's `self.gamma = self.gamma * factor`."""
    g = gamma0
    gammas, flags = [], []
    for i in range(1, n_iters + 1):
        gammas.append(float(g))
        d = 1 if (decay_steps and i % decay_steps == 0) else 0
        flags.append(d)
        if d:
            g = g * decay_factor
    gammas.append(float(g))
    return gammas, flags


def no_iteration_callback(iteration: int, objective_result: ObjectiveResult) -> None:
    """Pass as `iteration_callback` to run without per-iteration reporting: the fused loop then never materialises an
    ObjectiveResult per iteration (and, sharded, takes the objective's tail and the update in one launch)."""


class AcceleratedGradientDescent:
    def __init__(
        self,
        max_iter: int,
        gamma: float,
        initial_step_size: float = 1e-5,
        max_step_size: float = 0.1,
        gamma_decay_type: str = None,
        gamma_decay_params: dict = {},
        save_primal: bool = False,
        iteration_callback: Optional[Callable[[int, ObjectiveResult], None]] = None,
    ):
        self.initial_step_size = initial_step_size
        self.max_step_size = max_step_size
        self.max_iter = max_iter
        self.beta_seq = self._compute_beta_seq(self.max_iter)
        self.streams = None
        self.gamma = gamma
        self.gamma_decay_type = gamma_decay_type
        self.gamma_decay_params = gamma_decay_params
        self.save_primal = save_primal
        self.iteration_callback = iteration_callback if iteration_callback is not None else self._default_iteration_callback

    def _compute_beta_seq(self, max_iter: int) -> torch.Tensor:
        """beta_i = (1 - t_{i+1}) / t_{i+2}, t_{k} = (1 + sqrt(1 + 4 t_{k-1}^2))/2; t is stored in float32 and the
        square root is taken in double, exactly like the reference (agd.py:93-100)."""
        t = torch.zeros(max_iter + 2)
        for k in range(1, max_iter + 2):
            t[k] = (1 + math.sqrt(1 + 4 * (t[k - 1] ** 2))) / 2
        return (1 - t[1 : max_iter + 1]) / t[2 : max_iter + 2]

    def _update_gamma(self, itr: int, step_size: float):
        """Step decay: every `decay_steps` iterations gamma *= decay_factor and the step cap becomes
        step * decay_factor (reference agd.py:102-109)."""
        if self.gamma_decay_type != "step":
            raise ValueError(f"Unsupported gamma decay type: {self.gamma_decay_type}")
        if itr % self.gamma_decay_params["decay_steps"] == 0:
            factor = self.gamma_decay_params["decay_factor"]
            self.gamma = self.gamma * factor
            self.max_step_size = step_size * factor

    def _user_callback_active(self) -> bool:
        return self.iteration_callback is not no_iteration_callback

    def _default_iteration_callback(self, iteration: int, objective_result: ObjectiveResult) -> None:
        try:
            print(format_objective_result_summary(iteration, objective_result))
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    def maximize(self, f: BaseObjective, initial_value: torch.Tensor, rank: int = 0) -> SolverResult:
        """Maximise the dual objective of `f` starting from `initial_value` (reference agd.py:121-229)."""
        if self._fusable(f, initial_value):
            return self._maximize_fused(f, initial_value, rank)
        return self._maximize_generic(f, initial_value, rank)

    @staticmethod
    def _fusable(f, initial_value: torch.Tensor) -> bool:
        from dualip_b200.objectives.matching import (
            MatchingSolverDualObjectiveFunction,
            MatchingSolverDualObjectiveFunctionDistributed,
        )

        if not (isinstance(initial_value, torch.Tensor) and initial_value.is_cuda and initial_value.dtype == torch.float32):
            return False
        if isinstance(f, MatchingSolverDualObjectiveFunctionDistributed):
            return f.device == initial_value.device
        if isinstance(f, MatchingSolverDualObjectiveFunction):
            return not f.is_distributed and f.device == initial_value.device
        from dualip_b200.objectives.matching_fairness import MatchingFairnessDualObjectiveFunction
        from dualip_b200.objectives.miplib import MIPLIB2017ObjectiveFunction

        return isinstance(f, (MIPLIB2017ObjectiveFunction, MatchingFairnessDualObjectiveFunction)) and f.device == initial_value.device

    # -- fused device-resident loop ------------------------------------------------------------------------
    def _maximize_fused(self, f, initial_value: torch.Tensor, rank: int) -> SolverResult:
        loop = FusedAscentLoop(self, f, initial_value, rank)
        try:
            loop.run(1, self.max_iter)
            return loop.finish()
        finally:
            loop.close()

    @staticmethod
    def _callback_result(grad, scal, primal) -> ObjectiveResult:
        """What an iteration callback receives: fresh float32 tensors, like the reference hands out every iteration
        (matching.py:171-187) -- a callback may keep them; the loop's own buffers are overwritten by the next step."""
        return AcceleratedGradientDescent._view_result(grad.clone(), scal, primal, as_float32=True)

    @staticmethod
    def _view_result(grad, scal, primal, as_float32: bool = False) -> ObjectiveResult:
        s = scal.to(torch.float32) if as_float32 else scal
        res = ObjectiveResult(
            dual_gradient=grad,
            dual_objective=s[_IDX["dual_objective"]],
            reg_penalty=s[_IDX["reg_penalty"]],
            dual_val_times_grad=s[_IDX["dual_val_times_grad"]],
            max_pos_slack=s[_IDX["max_pos_slack"]],
            sum_pos_slack=s[_IDX["sum_pos_slack"]],
        )
        if primal is not None:
            res.primal_var = primal
            res.primal_objective = s[_IDX["primal_objective"]]
        return res

    # -- generic host-driven loop (user objectives, CPU tensors) -------------------------------------------
    def _maximize_generic(self, f, initial_value: torch.Tensor, rank: int) -> SolverResult:
        distributed = dist.is_available() and dist.is_initialized()
        everyone = bool(getattr(f, "result_on_all_ranks", False))
        if (isinstance(initial_value, torch.Tensor) and initial_value.device.type == "cpu" and initial_value.dtype == torch.float32
                and initial_value.dim() == 1 and (everyone or not distributed) and os.environ.get("DUALIP_HOST_STEP", "native") != "torch"):
            return self._maximize_host_native(f, initial_value, rank)
        return self._maximize_host_torch(f, initial_value, rank)

    def _maximize_host_native(self, f, initial_value: torch.Tensor, rank: int) -> SolverResult:
        """Host-driven loop with the update done by dualip_agd_host_step (one native call per iteration): float32 CPU
        iterates, any objective with `calculate` / `equality_mask`.  The evaluation point handed to `f.calculate` is a view
        of the optimizer's (pinned) host buffer."""
        import numpy as np

        lib = _native.lib()
        m = initial_value.numel()
        init = initial_value.detach().contiguous()
        eq = f.equality_mask
        eq_u8 = eq.detach().to(device="cpu", dtype=torch.uint8).contiguous() if eq is not None else None
        handle = ctypes.c_void_p()
        _native.check(lib.dualip_agd_host_create(ctypes.byref(handle), m, init.data_ptr(), eq_u8.data_ptr() if eq_u8 is not None else None,
                                                 float(self.initial_step_size), float(self.max_step_size), 15), "dualip_agd_host_create")
        try:
            def view(ptr):
                return torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(m,)))

            x, y = view(lib.dualip_agd_host_x(handle)), view(lib.dualip_agd_host_y(handle))
            beta = self.beta_seq.tolist()
            decay = self.gamma is not None and self.gamma_decay_type is not None
            if decay and self.gamma_decay_type != "step":
                raise ValueError(f"Unsupported gamma decay type: {self.gamma_decay_type}")
            dual_obj_log, step_size_log = [], []
            step = ctypes.c_double(0.0)
            dual_obj, objective_result = 0.0, None
            fast = self._native_host_target(f)
            if fast is not None and not (self.save_primal and self.max_iter >= 1):
                return self._maximize_host_fused(f, fast, lib, handle, x, y, beta, decay, rank)
            for i in range(1, self.max_iter + 1):
                kwargs = {"gamma": self.gamma} if self.gamma is not None else {}
                if i == self.max_iter and self.save_primal:
                    kwargs["save_primal"] = self.save_primal
                # a copy: the objective or the callback may keep what they are given, and x is native memory that the
                # optimizer state frees at the end of this call
                objective_result = f.calculate(dual_val=x.clone(), rank=rank, **kwargs)
                if rank == 0:
                    self.iteration_callback(i, objective_result)
                dual_obj = float(objective_result.dual_objective)
                dual_obj_log.append(dual_obj)
                grad = objective_result.dual_gradient
                if grad.device.type != "cpu" or grad.dtype != torch.float32 or not grad.is_contiguous():
                    grad = grad.detach().to(device="cpu", dtype=torch.float32).contiguous()
                decay_now, factor = 0, 1.0
                if decay and i % self.gamma_decay_params["decay_steps"] == 0:
                    decay_now, factor = 1, float(self.gamma_decay_params["decay_factor"])
                _native.check(lib.dualip_agd_host_step(handle, grad.data_ptr(), float(beta[i - 1]), decay_now, factor,
                                                       ctypes.byref(step)), "dualip_agd_host_step")
                step_size_log.append(step.value)
                if decay_now:  # agd.py:102-109
                    self.gamma = self.gamma * factor
                    self.max_step_size = step.value * factor
            return SolverResult(dual_val=y.clone(), dual_objective=dual_obj, objective_result=objective_result,
                                dual_objective_log=dual_obj_log, step_size_log=step_size_log)
        finally:
            lib.dualip_agd_host_destroy(handle)

    @staticmethod
    def _native_host_target(f):
        """(plan, peer handle or None, b_vec, local objective) when a whole host-buffer iteration can be ONE native call
        (dualip_matching_step_host): a matching objective whose columns are all projected natively, on one GPU or sharded with
        the peer-memory exchange.  COLLECTIVE for sharded objectives (the windows are set up on first use)."""
        from dualip_b200.objectives.matching import (
            MatchingSolverDualObjectiveFunction,
            MatchingSolverDualObjectiveFunctionDistributed,
        )

        if os.environ.get("DUALIP_HOST_FUSED", "1") == "0":
            return None
        if isinstance(f, MatchingSolverDualObjectiveFunctionDistributed):
            local = f.local_objective
            if local.has_block_entries:
                return None
            peer = f.peer_exchange()
            return None if peer is None else (local._plan, peer.handle, f.b_vec, local)
        if type(f) is MatchingSolverDualObjectiveFunction and not f.is_distributed and not f.has_block_entries:
            return (f._plan, None, f.b_vec, f)
        return None

    def _maximize_host_fused(self, f, target, lib, handle, x, y, beta, decay, rank) -> SolverResult:
        """Host-buffer loop with one native call per iteration: lambda host->device, fused kernel (sharded: + exchange through
        peer memory), gradient + scalars device->host, host-side accelerated step.  A callback still sees every iteration's
        result (fresh tensors); without one no Python object is built per iteration."""
        import numpy as np

        plan, peer, b_vec, local = target
        m = x.numel()
        h_grad = torch.empty(m, dtype=torch.float32).pin_memory()
        h_scal = torch.empty(len(_native.SCALAR_FIELDS), dtype=torch.float64).pin_memory()
        scal_np = h_scal.numpy()
        local.check_inputs_unchanged()
        callback = rank == 0 and self._user_callback_active()
        dual_obj_log, step_size_log = [], []
        step = ctypes.c_double(0.0)
        b_ptr = b_vec.data_ptr() if b_vec is not None else None

        def result():
            from dualip_b200.objectives.matching import _host_result

            return _host_result(h_grad.clone(), h_scal.clone(), False)

        objective_result = None
        with torch.cuda.device(local.device):
            stream = torch.cuda.current_stream(local.device).cuda_stream
            for i in range(1, self.max_iter + 1):
                gamma_i = self.gamma if self.gamma is not None else f.gamma
                decay_now, factor = 0, 1.0
                if decay and i % self.gamma_decay_params["decay_steps"] == 0:
                    decay_now, factor = 1, float(self.gamma_decay_params["decay_factor"])
                _native.check(lib.dualip_matching_step_host(plan, peer, handle, b_ptr, float(gamma_i), float(beta[i - 1]), decay_now, factor,
                                                            h_grad.data_ptr(), h_scal.data_ptr(), ctypes.byref(step), stream),
                              "dualip_matching_step_host")
                local.launched()
                dual_obj_log.append(float(scal_np[_IDX["dual_objective"]]))
                step_size_log.append(step.value)
                if callback:
                    objective_result = result()
                    self.iteration_callback(i, objective_result)
                if decay_now:  # agd.py:102-109
                    self.gamma = self.gamma * factor
                    self.max_step_size = step.value * factor
                    f.gamma = self.gamma
        if objective_result is None or not callback:
            objective_result = result() if self.max_iter >= 1 else None
        return SolverResult(dual_val=y.clone(), dual_objective=dual_obj_log[-1] if dual_obj_log else 0.0, objective_result=objective_result,
                            dual_objective_log=dual_obj_log, step_size_log=step_size_log)

    def _maximize_host_torch(self, f, initial_value: torch.Tensor, rank: int) -> SolverResult:
        """The reference's loop op for op (any device / dtype; rank-0 update + broadcasts for reference-style objectives)."""
        grad_history, dual_history, lipschitz_cache = [], [], []
        dual_obj_log, step_size_log = [], []
        x = initial_value.clone()
        y = initial_value.clone()
        equality_mask = f.equality_mask
        everyone_updates = bool(getattr(f, "result_on_all_ranks", False))
        distributed = dist.is_available() and dist.is_initialized()
        dual_obj = 0.0
        objective_result = None
        for i in range(1, self.max_iter + 1):
            kwargs = {"gamma": self.gamma} if self.gamma is not None else {}
            if i == self.max_iter and self.save_primal:
                kwargs["save_primal"] = self.save_primal
            objective_result = f.calculate(dual_val=x, rank=rank, **kwargs)
            if rank == 0 or everyone_updates:
                if rank == 0:
                    self.iteration_callback(i, objective_result)
                dual_obj = objective_result.dual_objective.cpu().item()
                dual_obj_log.append(dual_obj)
                step_size = calculate_step_size(objective_result.dual_gradient, y, grad_history, dual_history,
                                                initial_step_size=self.initial_step_size, max_step_size=self.max_step_size,
                                                lipschitz_cache=lipschitz_cache)
                step_size_log.append(step_size)
                y_new = project_on_nn_cone(x + objective_result.dual_gradient * step_size, equality_mask)
                b_i = self.beta_seq[i - 1]
                x = (y_new * (1.0 - b_i)) + (y * b_i)
                y = y_new
                if self.gamma is not None and self.gamma_decay_type is not None:
                    self._update_gamma(i, step_size)
            if distributed and not everyone_updates:
                dist.broadcast(x, src=0)
                dist.broadcast(y, src=0)
        if rank == 0 or everyone_updates:
            return SolverResult(dual_val=y, dual_objective=dual_obj, objective_result=objective_result,
                                dual_objective_log=dual_obj_log, step_size_log=step_size_log)
        return SolverResult(dual_val=y, dual_objective=0.0, objective_result=objective_result, dual_objective_log=[],
                            step_size_log=[])


class FusedAscentLoop:
    """Device-resident state of one `maximize` call on a CUDA matching objective.

    `step(i)` enqueues iteration i (1-based) on the current stream: the objective's kernel(s) at the evaluation point
    x held by the native optimizer state, then dualip_agd_step.  Nothing synchronises until `finish()`.  bench.py
    drives this class directly so that it can bracket exactly K iterations with CUDA events."""

    def __init__(self, solver: AcceleratedGradientDescent, f, initial_value: torch.Tensor, rank: int = 0):
        from dualip_b200.objectives.matching import MatchingSolverDualObjectiveFunctionDistributed

        self.solver, self.f, self.rank = solver, f, rank
        self.lib = _native.lib()
        self.device = initial_value.device
        self.m = initial_value.numel()
        self.sharded = isinstance(f, MatchingSolverDualObjectiveFunctionDistributed)
        if solver.save_primal and self.sharded:
            raise NotImplementedError("save_primal=True is not yet supported in distributed mode")
        self.decay = solver.gamma is not None and solver.gamma_decay_type is not None
        if self.decay and solver.gamma_decay_type != "step":
            raise ValueError(f"Unsupported gamma decay type: {solver.gamma_decay_type}")
        eq = f.equality_mask
        self._eq_u8 = eq.to(device=self.device, dtype=torch.uint8).contiguous() if eq is not None else None
        init = initial_value.detach().contiguous()
        self.beta = solver.beta_seq.tolist()
        self.handle = ctypes.c_void_p()
        self.steps_done = 0
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            _native.check(self.lib.dualip_agd_create(
                ctypes.byref(self.handle), self.m, self.device.index, init.data_ptr(),
                self._eq_u8.data_ptr() if self._eq_u8 is not None else None, float(solver.initial_step_size),
                float(solver.max_step_size), 15), "dualip_agd_create")
            _native.check(self.lib.dualip_agd_reserve_log(self.handle, max(solver.max_iter, 1)))
            self.x_ptr = self.lib.dualip_agd_x(self.handle)
            self.grad = torch.empty(self.m, dtype=torch.float32, device=self.device)
            self.scal = torch.zeros(len(_native.SCALAR_FIELDS), dtype=torch.float64, device=self.device)
            local = f.local_objective if self.sharded else f
            if hasattr(local, "check_inputs_unchanged"):
                local.check_inputs_unchanged()  # the raw launches below bypass calculate()
            self.block_entries = bool(getattr(local, "has_block_entries", False))  # user-registered projections
            need_partial = self.sharded or self.block_entries
            self.partial = torch.empty(self.m + 2, dtype=torch.float32, device=self.device) if need_partial else None
            self._xbuf = torch.empty(self.m, dtype=torch.float32, device=self.device) if self.block_entries else None
        # sharded, no per-iteration callback: the partial sums are exchanged through peer memory inside the update kernel
        # (the choice must not depend on the rank: every rank takes the same path)
        self.peer = f.peer_exchange() if (self.sharded and not solver._user_callback_active()) else None
        if self.peer is not None and dist.is_available() and dist.is_initialized():
            # the windows may be reused from an earlier maximize(): nobody starts polling flags before every rank is here
            dist.barrier()
        # matching objectives whose columns are all projected natively: evaluation and step in one launch
        from dualip_b200.objectives.matching import MatchingSolverDualObjectiveFunction

        self.one_launch = (not self.block_entries and os.environ.get("DUALIP_ONE_LAUNCH", "1") != "0"
                           and isinstance(local, MatchingSolverDualObjectiveFunction))
        # plan self-tuning needs a host synchronisation now and then: fine for one shard per process (every rank does it at
        # the same iteration), not offered to objectives without a plan
        self._tune = local if hasattr(local, "launched") else None
        self.primal = None
        self.kernel_events = None  # optional list of (start, end) CUDA events per step, for measurement
        self.kernel_events_base = 1
        # CUDA-graph replay (run()): the one-launch iteration with its per-iteration scalars read from a device-resident
        # schedule, so that `chunk` identical launches are captured once and replayed as one submission
        self.graph_chunk = int(os.environ.get("DUALIP_GRAPH_CHUNK", "50"))
        self.graphable = (self.one_launch and not solver._user_callback_active() and (not self.sharded or self.peer is not None)
                          and os.environ.get("DUALIP_GRAPH", "1") != "0" and self.graph_chunk > 0)
        self._graph = ctypes.c_void_p()
        self.graph_launches = 0
        if self.graphable:
            self._install_schedule()

    def _install_schedule(self) -> None:
        """gamma / beta / decay flag of every iteration, exactly as step() would pass them (agd.py:93-109)."""
        solver, n = self.solver, max(self.solver.max_iter, 1)
        g = solver.gamma if solver.gamma is not None else self.f.gamma
        steps = solver.gamma_decay_params["decay_steps"] if self.decay else 0
        factor = float(solver.gamma_decay_params["decay_factor"]) if self.decay else 1.0
        gammas, flags = gamma_schedule(g, n, steps, factor)
        self._gamma_sched = gammas  # [k] = gamma before iteration k+1; [n] = after the last
        beta = (ctypes.c_float * n)(*self.beta[:n])
        with torch.cuda.device(self.device):
            _native.check(self.lib.dualip_agd_set_schedule(self.handle, n, (ctypes.c_double * n)(*gammas[:n]), beta,
                                                           (ctypes.c_uint8 * n)(*flags), factor), "dualip_agd_set_schedule")

    def run(self, first: int, last: int) -> None:
        """Iterations first..last (1-based, inclusive).  Whole chunks go out as CUDA-graph replays once the plan has stopped
        re-cutting its ranges; everything else (the first launches of an objective, a last iteration that writes the primal,
        the remainder, measurement with per-kernel events) is launched one by one through step()."""
        i = first
        while i <= last:
            c = self.graph_chunk
            end = i + c - 1
            if (self.graphable and self.kernel_events is None and end <= last and self.steps_done == i - 1
                    and not (self.solver.save_primal and end >= self.solver.max_iter)
                    and (self._tune is None or self._tune.plan_settled())):
                self._launch_graph(i)
                i = end + 1
            else:
                self.step(i)
                i += 1

    def _launch_graph(self, i: int) -> None:
        solver, f = self.solver, self.f
        local = f.local_objective if self.sharded else f
        with torch.cuda.device(self.device):
            if not self._graph:
                b_ptr = f.b_vec.data_ptr() if f.b_vec is not None else None
                _native.check(self.lib.dualip_ascent_graph_create(
                    ctypes.byref(self._graph), local._plan, self.handle, self.peer.handle if self.peer is not None else None,
                    b_ptr, self.grad.data_ptr(), self.scal.data_ptr(), self.graph_chunk), "dualip_ascent_graph_create")
            _native.check(self.lib.dualip_ascent_graph_launch(self._graph, torch.cuda.current_stream(self.device).cuda_stream),
                          "dualip_ascent_graph_launch")
        self.graph_launches += 1
        end = i + self.graph_chunk - 1
        g = self._gamma_sched[end]
        if self.decay:
            solver.gamma = g
        f.gamma = g
        if self.sharded:
            f.local_objective.gamma = g
        if self._tune is not None:
            self._tune.launched_many(self.graph_chunk)
        if self.peer is not None and self.peer.status_nowait():
            self._peer_timed_out()
        self.steps_done = end

    def step(self, i: int) -> None:
        solver, f = self.solver, self.f
        gamma_i = solver.gamma if solver.gamma is not None else f.gamma
        f.gamma = gamma_i
        last_primal = i == solver.max_iter and solver.save_primal
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            ev = None
            if self.kernel_events is not None and 0 <= i - self.kernel_events_base < len(self.kernel_events):
                ev = self.kernel_events[i - self.kernel_events_base]
                ev[0].record()
            decay_now, factor = 0, 1.0
            if self.decay and i % solver.gamma_decay_params["decay_steps"] == 0:
                decay_now, factor = 1, float(solver.gamma_decay_params["decay_factor"])
            callback = self.rank == 0 and solver._user_callback_active()
            if self.sharded and self.peer is not None:
                f.local_objective.gamma = gamma_i
                if self.one_launch:
                    # ONE launch: shard kernel whose last CTA publishes the partial sums, reads every peer's over NVLink,
                    # runs the objective's tail and takes the step -- no collective call, no second kernel
                    f.local_objective.launch_ascent_step_peer(self.handle, self.peer.handle, f.b_vec.data_ptr(), gamma_i,
                                                              self.grad.data_ptr(), self.scal.data_ptr(), self.beta[i - 1],
                                                              decay_now, factor, i - 1)
                    if ev is not None:
                        ev[1].record()
                else:
                    # shard kernel -> update kernel that reads every peer's partial sums over NVLink
                    f.local_objective.launch_partial(self.x_ptr, gamma_i, self.peer.next_slot())
                    if ev is not None:
                        ev[1].record()
                    _native.check(self.lib.dualip_agd_step_peer(
                        self.handle, self.peer.handle, f.b_vec.data_ptr(), float(gamma_i), self.grad.data_ptr(),
                        self.scal.data_ptr(), float(self.beta[i - 1]), decay_now, factor, i - 1, stream), "dualip_agd_step_peer")
                if (i & 31) == 0 and self.peer.status_nowait():
                    self._peer_timed_out()
            elif self.sharded:
                from dualip_b200.objectives.matching import reduce_partials

                f.local_objective.gamma = gamma_i
                f.local_objective.launch_partial(self.x_ptr, gamma_i, self.partial.data_ptr())
                if self.block_entries:
                    f.local_objective.add_block_entries(self._x_tensor(stream), gamma_i, self.partial)
                if ev is not None:
                    ev[1].record()
                reduce_partials(self.partial)
                if callback:
                    # the callback wants the result before the update: separate epilogue launch, then the plain step
                    f.launch_epilogue(self.partial.data_ptr(), self.x_ptr, gamma_i, self.grad.data_ptr(), self.scal.data_ptr())
                    solver.iteration_callback(i, solver._callback_result(self.grad, self.scal, None))
                    _native.check(self.lib.dualip_agd_step(self.handle, self.grad.data_ptr(), self.scal.data_ptr(),
                                                           float(self.beta[i - 1]), decay_now, factor, i - 1, stream),
                                  "dualip_agd_step")
                else:
                    # objective tail (grad = sum - b, scalars) and optimizer update in ONE launch
                    _native.check(self.lib.dualip_agd_step_sharded(
                        self.handle, self.partial.data_ptr(), f.b_vec.data_ptr(), float(gamma_i), self.grad.data_ptr(),
                        self.scal.data_ptr(), float(self.beta[i - 1]), decay_now, factor, i - 1, stream),
                        "dualip_agd_step_sharded")
            elif self.block_entries:
                # natively projected columns in the fused kernel, user-registered projections through padded blocks, then
                # the objective's tail and the update in one launch (or epilogue -> callback -> update)
                if last_primal:
                    self.primal = torch.empty(f.nnz, dtype=torch.float32, device=self.device)
                f.launch_partial(self.x_ptr, gamma_i, self.partial.data_ptr(), self.primal.data_ptr() if last_primal else None)
                f.add_block_entries(self._x_tensor(stream), gamma_i, self.partial, self.primal if last_primal else None)
                if ev is not None:
                    ev[1].record()
                b_ptr = f.b_vec.data_ptr() if f.b_vec is not None else None
                if callback or last_primal:
                    _native.check(self.lib.dualip_matching_epilogue(self.partial.data_ptr(), self.m, self.x_ptr, b_ptr, float(gamma_i),
                                                                    self.grad.data_ptr(), self.scal.data_ptr(), stream))
                    if callback:
                        solver.iteration_callback(i, solver._callback_result(self.grad, self.scal, self.primal if last_primal else None))
                    _native.check(self.lib.dualip_agd_step(self.handle, self.grad.data_ptr(), self.scal.data_ptr(),
                                                           float(self.beta[i - 1]), decay_now, factor, i - 1, stream), "dualip_agd_step")
                else:
                    _native.check(self.lib.dualip_agd_step_sharded(
                        self.handle, self.partial.data_ptr(), b_ptr, float(gamma_i), self.grad.data_ptr(), self.scal.data_ptr(),
                        float(self.beta[i - 1]), decay_now, factor, i - 1, stream), "dualip_agd_step_sharded")
            elif self.one_launch:
                # evaluation at the optimizer's evaluation point and the accelerated step in ONE launch; a callback sees the
                # result the kernel wrote before it stepped
                if last_primal:
                    self.primal = torch.empty(f.nnz, dtype=torch.float32, device=self.device)
                f.launch_ascent_step(self.handle, gamma_i, self.grad.data_ptr(), self.scal.data_ptr(), self.beta[i - 1],
                                     decay_now, factor, i - 1, self.primal.data_ptr() if last_primal else None)
                if ev is not None:
                    ev[1].record()
                if callback:
                    solver.iteration_callback(i, solver._callback_result(self.grad, self.scal, self.primal if last_primal else None))
            else:
                if last_primal:
                    self.primal = torch.empty(getattr(f, "primal_size", f.nnz), dtype=torch.float32, device=self.device)
                f.launch_calc(self.x_ptr, gamma_i, self.grad.data_ptr(), self.scal.data_ptr(),
                              self.primal.data_ptr() if last_primal else None)
                if ev is not None:
                    ev[1].record()
                if callback:
                    solver.iteration_callback(i, solver._callback_result(self.grad, self.scal, self.primal if last_primal else None))
                _native.check(self.lib.dualip_agd_step(self.handle, self.grad.data_ptr(), self.scal.data_ptr(),
                                                       float(self.beta[i - 1]), decay_now, factor, i - 1, stream),
                              "dualip_agd_step")
            if decay_now:
                solver.gamma = solver.gamma * factor
            if self._tune is not None:
                self._tune.launched()  # re-cuts the plan's per-CTA ranges after a few launches (a stream sync each time)
        self.steps_done = max(self.steps_done, i)

    def _peer_timed_out(self):
        """A rank did not arrive within the time-out: this run is invalid and the windows are unusable from now on."""
        self.f._peer_failed = True
        self.f._peer = None
        raise RuntimeError("peer exchange: a rank (status 1) or a CTA of this rank's own grid (status 2) did not arrive within the "
                           "time-out (ranks out of step?); results of this run are invalid.  DUALIP_PEER_EXCHANGE=0 selects the NCCL path")

    def _x_tensor(self, stream) -> torch.Tensor:
        """The evaluation point as a tensor (a copy of the native state's x), for the tensor-op part of block entries."""
        _native.check(self.lib.dualip_agd_get(self.handle, self._xbuf.data_ptr(), None, stream))
        return self._xbuf

    def current_dual(self) -> torch.Tensor:
        y = torch.empty(self.m, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _native.check(self.lib.dualip_agd_get(self.handle, None, y.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream))
        return y

    def finish(self) -> SolverResult:
        solver, n = self.solver, self.steps_done
        obj_log = (ctypes.c_double * max(n, 1))()
        step_log = (ctypes.c_double * max(n, 1))()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self.lib.dualip_agd_read_log(self.handle, n, obj_log, step_log, stream), "dualip_agd_read_log")
            y = self.current_dual()
            torch.cuda.synchronize(self.device)
            if self.peer is None and hasattr(self._tune, "check_grid_barrier"):
                self._tune.check_grid_barrier()  # all-CTA tail: a grid-wide barrier that timed out invalidates the run
            if self.peer is not None:
                timed_out = self.peer.status()
                if dist.is_available() and dist.is_initialized() and self.peer.world == dist.get_world_size():
                    dist.barrier()  # nobody frees or reuses a window a peer may still be reading
                if timed_out:
                    self._peer_timed_out()
        dual_obj_log = [float(v) for v in obj_log[:n]]
        step_size_log = [float(v) for v in step_log[:n]]
        if self.decay and step_size_log:
            # host mirror of the device-side step cap (agd.py:107), for callers that inspect the optimizer afterwards
            steps = solver.gamma_decay_params["decay_steps"]
            k = (n // steps) * steps
            if k >= 1:
                solver.max_step_size = step_size_log[k - 1] * solver.gamma_decay_params["decay_factor"]
        final = solver._view_result(self.grad, self.scal, self.primal, as_float32=True)
        return SolverResult(dual_val=y, dual_objective=dual_obj_log[-1] if dual_obj_log else 0.0, objective_result=final,
                            dual_objective_log=dual_obj_log, step_size_log=step_size_log)

    def close(self) -> None:
        if self._graph:
            self.lib.dualip_ascent_graph_destroy(self._graph)
            self._graph = ctypes.c_void_p()
        if self.handle:
            self.lib.dualip_agd_destroy(self.handle)
            self.handle = ctypes.c_void_p()
