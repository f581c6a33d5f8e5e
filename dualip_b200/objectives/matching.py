"""Matching objective: same class names, constructor signatures and `calculate` keywords as the reference
(src/dualip/objectives/matching.py), with the body of `calculate` replaced by one call into the C-ABI
(include/dualip_b200.h: dualip_matching_calc; sharded: dualip_matching_partial + all-reduce + dualip_matching_epilogue, or
dualip_matching_calc_peer_host for a host-resident dual).  The Maximizer drives the same plan through
dualip_matching_ascent_step[_peer / _scheduled] (evaluation + step in one launch).

CUDA-only: tensors must live on a CUDA device and be float32.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from dualip_b200 import _native
from dualip_b200.objectives.base import BaseInputArgs, BaseObjective, ObjectiveResult
from dualip_b200.projections.base import ProjectionEntry, project

_IDX = {name: i for i, name in enumerate(_native.SCALAR_FIELDS)}


@dataclass
class MatchingInputArgs(BaseInputArgs):
    """Input arguments of the matching objective (reference matching.py:12-22).

    A and c are `torch.sparse_csc` tensors of shape (m, n) sharing one sparsity pattern; b_vec=None marks a local
    shard whose partial results are reduced by the distributed wrapper."""

    A: torch.Tensor
    c: torch.Tensor
    projection_map: dict[str, ProjectionEntry]
    b_vec: torch.Tensor
    equality_mask: torch.Tensor = None


def calc_grad(dual_grad: torch.Tensor, dual_obj: torch.Tensor, dual_val: torch.Tensor, b_vec: torch.Tensor,
              reg_penalty: torch.Tensor) -> tuple:
    """grad - b and c.x + reg + lambda.(grad - b) (reference matching.py:25-34).  The stock objectives do this in the fused
    kernel's m-length tail; the function is public because the reference's extension recipe calls it
    (docs/demo/matching_complex.rst:139)."""
    dual_grad = dual_grad - b_vec
    dual_obj = dual_obj + reg_penalty + torch.dot(dual_val, dual_grad)
    return dual_grad, dual_obj


def _indices_to_device(indices, device) -> torch.Tensor:
    if isinstance(indices, range):
        if indices.step == 1:
            return torch.arange(indices.start, indices.stop, dtype=torch.int64, device=device)
        return torch.arange(indices.start, indices.stop, indices.step, dtype=torch.int64, device=device)
    if isinstance(indices, torch.Tensor):
        return indices.to(device=device, dtype=torch.int64)
    return torch.as_tensor(np.asarray(indices, dtype=np.int64), device=device)


class _BlockEntry:
    """A projection entry the fused kernel does not implement (a user-registered operator): its columns are projected
    through zero-padded [L x K] blocks, bucketed by column length exactly like the reference (matching.py:87-114,
    utils/sparse_utils.py:133-220).  Index tensors are built once."""

    def __init__(self, key, entry, cols: torch.Tensor, ccol: torch.Tensor, n_rows: int, batching: bool):
        self.key, self.proj_type, self.proj_params = key, entry.proj_type, entry.proj_params
        device = ccol.device
        lengths_all = ccol[1:] - ccol[:-1]
        if batching:
            thresholds, i = [0], 1
            while 2**i <= n_rows:
                thresholds.append(2**i)
                i += 1
            thresholds.append(n_rows + 1)
            ids = torch.bucketize(lengths_all[cols], torch.tensor(thresholds, dtype=lengths_all.dtype, device=device))
            groups = [cols[ids == j] for j in range(1, len(thresholds))]
        else:
            groups = [cols]
        self.buckets = []  # (offset into this entry's entry list, count, idx_in_col, cols_rep, L, K)
        flat, offset = [], 0
        for g in groups:
            K = g.numel()
            if K == 0:
                continue
            starts = ccol[g]
            lengths = ccol[g + 1] - starts
            total = int(lengths.sum().item())
            if total == 0:
                continue
            L = int(lengths.max().item())
            cols_rep = torch.arange(K, device=device).repeat_interleave(lengths)
            prefix = lengths.cumsum(0) - lengths
            idx_in_col = torch.arange(total, device=device) - prefix[cols_rep]
            flat.append(starts[cols_rep] + idx_in_col)
            self.buckets.append((offset, total, idx_in_col, cols_rep, L, K))
            offset += total
        self.entries = torch.cat(flat) if flat else torch.zeros(0, dtype=torch.int64, device=device)


def _build_class_table(ccol: torch.Tensor, n_cols: int, projection_map, batching: bool, n_rows: int = 0, block_entries=None):
    """Turns the reference's projection_map (dict of ProjectionEntry) into the C-ABI class table and a per-column
    class id.  Returns (classes ctypes array, n_classes, col_class uint8 tensor or None when one entry covers all).

    Also derives, per simplex entry, whether its 1-entry columns would sit in a bucket of padded length 1 in the
    reference (matching.py:87-114 thresholds {1,2},{3,4},{5..8},...; batching=False: one bucket), because the
    reference's top-2 shortcut is skipped there (simplex.py:166)."""
    device = ccol.device
    classes = []
    col_class = None
    entries = list(projection_map.items())
    lengths = None
    single_full = False
    if len(entries) == 1:
        ind = entries[0][1].indices
        if isinstance(ind, range) and ind.start == 0 and ind.step == 1 and ind.stop == n_cols:
            single_full = True
        elif not isinstance(ind, range) and len(ind) == n_cols:
            t = _indices_to_device(ind, device)
            single_full = bool((t == torch.arange(n_cols, device=device)).all().item())
    if not single_full:
        # class 0 = identity for columns no entry names (the reference leaves them unprojected)
        classes.append(_native.ProjClass(_native.PROJ_CLAMP, -np.inf, np.inf, 1.0, 1.0, 0))
        col_class = torch.zeros(n_cols, dtype=torch.uint8, device=device)
    for key, entry in entries:
        op = project(entry.proj_type, **entry.proj_params)  # raises ValueError for unknown names, like the reference
        cls = op.native_class() if hasattr(op, "native_class") else None
        idx = None if single_full else _indices_to_device(entry.indices, device)
        if cls is not None and cls.kind >= _native.PROJ_SIMPLEX_BISECT:
            # the fused kernel's bisection handles columns of up to 1024 entries; an entry with a longer column keeps the
            # padded-block route (same operator, dualip_project_block)
            if lengths is None:
                lengths = ccol[1:] - ccol[:-1]
            sel = lengths if idx is None else lengths[idx]
            if sel.numel() and int(sel.max()) > _native.MAX_BISECT_COLUMN:
                cls = None
        if cls is None:
            # not implemented by the fused kernel: its columns produce x = 0 there (clamp to [0, 0]) and are projected
            # through padded blocks by the objective (_BlockEntry)
            if block_entries is None:
                raise ValueError(f"projection '{entry.proj_type}' has no native class")
            cols = torch.arange(n_cols, dtype=torch.int64, device=device) if idx is None else idx
            block_entries.append(_BlockEntry(key, entry, cols, ccol, n_rows, batching))
            cls = _native.ProjClass(_native.PROJ_CLAMP, 0.0, 0.0, 1.0, 1.0, 0)
        if idx is not None and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n_cols):
            raise IndexError(f"projection entry '{key}' names a column outside [0, {n_cols})")
        if cls.kind != _native.PROJ_CLAMP:
            if lengths is None:
                lengths = ccol[1:] - ccol[:-1]
            sel = lengths if idx is None else lengths[idx]
            if sel.numel():
                if batching:
                    unpadded = bool(((sel == 1).any() & ~(sel == 2).any()).item())
                else:
                    unpadded = bool((sel.max() == 1).item())
                if unpadded:
                    cls.flags |= _native.PROJ_FLAG_D1_UNPADDED
        if idx is not None:
            if len(classes) >= 255:
                raise ValueError("at most 254 projection entries are supported")
            if idx.numel():
                if bool((col_class[idx] != 0).any().item()):
                    raise ValueError(
                        f"projection entry '{key}' overlaps an earlier entry; dualip_b200 projects each column once"
                    )
                col_class[idx] = len(classes)
        classes.append(cls)
    arr = (_native.ProjClass * len(classes))(*classes)
    return arr, len(classes), col_class


def _build_pad_table(ccol: torch.Tensor, n_cols: int, projection_map, batching: bool, n_rows: int, n_classes: int):
    """Padded block lengths for `simplex_eq` entries, the only projection whose result depends on them:
    the reference projects a column inside a zero-padded [L x K] block, L = longest column of the length bucket the column
    falls in (matching.py:87-114: buckets (0,2], (2,4], (4,8], ..., (2^k, m]; batching=False: one bucket per entry;
    utils/sparse_utils.py:197,207).  Returns a ctypes int32 array n_classes x 32 indexed by ceil(log2(d)), or None."""
    entries = list(projection_map.items())
    def needs_pad(e):  # simplex_eq (Duchi: SURVEY App. A #4) and both bisection variants (the padding is part of their block)
        return e.proj_type in ("simplex", "simplex_eq") and (e.proj_type == "simplex_eq" or e.proj_params.get("method") == "bisection_search")

    if not any(needs_pad(e) for _, e in entries):
        return None
    table = np.zeros((n_classes, _native.PAD_BUCKETS), dtype=np.int32)
    lengths = ccol[1:] - ccol[:-1]
    first = n_classes - len(entries)  # class 0 is the identity class when the map does not cover every column
    for i, (_, entry) in enumerate(entries):
        if not needs_pad(entry):
            continue
        ind = entry.indices
        full = isinstance(ind, range) and ind.start == 0 and ind.step == 1 and ind.stop == n_cols
        sel = lengths if full else lengths[_indices_to_device(ind, ccol.device)]
        sel = sel[sel > 0]
        if sel.numel() == 0:
            continue
        if batching:
            bkt = torch.ceil(torch.log2(sel.to(torch.float64))).to(torch.int64).clamp_(min=1)  # lengths 1 and 2 share a bucket
            lmax = torch.zeros(_native.PAD_BUCKETS, dtype=torch.int64, device=ccol.device)
            lmax.scatter_reduce_(0, bkt, sel.to(torch.int64), reduce="amax", include_self=True)
            row = lmax.cpu().numpy()
            row[0] = row[1]
            # the reference's last bucket is (2^k, m+1] with 2^k <= m: lengths above 2^k share it whatever their log2
            k = int(np.floor(np.log2(max(n_rows, 1))))
            if k + 1 < _native.PAD_BUCKETS:
                row[k + 1:] = row[k + 1:].max()
        else:
            row = np.full(_native.PAD_BUCKETS, int(sel.max().item()), dtype=np.int64)
        table[first + i] = row
    return (ctypes.c_int32 * table.size)(*table.reshape(-1).tolist())


class MatchingSolverDualObjectiveFunction(BaseObjective):
    """Dual gradient, objective and regularisation penalty of a matching LP on one GPU.

    Drop-in for the reference class of the same name (matching.py:37-188): same constructor, same `calculate`
    keywords, same ObjectiveResult fields.  Construction builds the device-side slab layout
    (dualip_plan_create); `calculate` is one fused kernel launch."""

    def __init__(self, matching_input_args: MatchingInputArgs, gamma: float, batching: bool = True):
        A, c = matching_input_args.A, matching_input_args.c
        if A.layout != torch.sparse_csc or c.layout != torch.sparse_csc:
            raise ValueError("Both A and c must be CSC-format sparse tensors")
        if not A.is_cuda:
            raise RuntimeError("dualip_b200 objectives need CUDA tensors (no CPU fallback); move the inputs to a GPU")
        if A.values().dtype != torch.float32 or c.values().dtype != torch.float32:
            raise TypeError("dualip_b200 is float32-only")
        if A.shape != c.shape or A.values().shape != c.values().shape:
            raise ValueError("A and c must share the same sparsity pattern")
        self.A, self.c = A, c
        self.gamma = gamma
        self.b_vec = matching_input_args.b_vec
        self.projection_map = matching_input_args.projection_map
        self.is_distributed = self.b_vec is None
        self.equality_mask = matching_input_args.equality_mask
        self.batching = batching
        self.device = A.device
        self.m, self.n = int(A.shape[0]), int(A.shape[1])
        if self.b_vec is not None:
            if self.b_vec.device != self.device or self.b_vec.dtype != torch.float32:
                self.b_vec = self.b_vec.to(device=self.device, dtype=torch.float32)
            self.b_vec = self.b_vec.contiguous()

        ccol, row = A.ccol_indices(), A.row_indices()
        if ccol.dtype != row.dtype or ccol.dtype not in (torch.int32, torch.int64):
            raise TypeError("ccol_indices and row_indices must both be int32 or both int64")
        # keep the borrowed value arrays alive for the lifetime of the plan
        self._a_vals = A.values().contiguous()
        self._c_vals = c.values().contiguous()
        self.nnz = int(self._a_vals.numel())
        self._block_entries = []
        classes, n_classes, col_class = _build_class_table(ccol, self.n, self.projection_map, batching, self.m,
                                                           self._block_entries)
        if self._block_entries:
            ec = torch.cat([e.entries for e in self._block_entries])
            self._blk_idx = ec
            self._blk_a, self._blk_c = self._a_vals[ec], self._c_vals[ec]
            self._blk_row = row[ec].to(torch.int64)
        self._classes = classes
        self._pad = _build_pad_table(ccol, self.n, self.projection_map, batching, self.m, n_classes)
        desc = _native.CscDesc(
            n_cols=self.n, nnz=self.nnz, n_rows=self.m, index_bits=32 if ccol.dtype == torch.int32 else 64,
            ccol_dev=ccol.data_ptr(), row_dev=row.data_ptr() if self.nnz else 0,
            a_dev=self._a_vals.data_ptr() if self.nnz else 0, c_dev=self._c_vals.data_ptr() if self.nnz else 0,
            col_class_dev=col_class.data_ptr() if col_class is not None else None,
            classes=ctypes.cast(classes, ctypes.POINTER(_native.ProjClass)), n_classes=n_classes,
            device=self.device.index if self.device.index is not None else torch.cuda.current_device(),
            pad_len=ctypes.cast(self._pad, ctypes.POINTER(ctypes.c_int32)) if self._pad is not None else None,
        )
        handle = ctypes.c_void_p()
        torch.cuda.synchronize(self.device)
        _native.check(_native.lib().dualip_plan_create(ctypes.byref(handle), ctypes.byref(desc)), "dualip_plan_create")
        self._plan = handle
        # the plan holds a COPY of A's and c's values in its own layout; the reference reads the tensors at every call
        # (matching.py:136-142).  An in-place edit after construction would be silently ignored, so it is detected instead.
        self._value_versions = (self._a_vals._version, self._c_vals._version)
        self._launches = 0
        self._rebalance_at = (4, 8, 16, 32, 64, 128) if os.environ.get("DUALIP_REBALANCE", "1") != "0" else ()
        self._scal = torch.zeros(len(_native.SCALAR_FIELDS), dtype=torch.float64, device=self.device)
        self._grad = torch.empty(self.m, dtype=torch.float32, device=self.device)
        self._partial = torch.empty(self.m + 2, dtype=torch.float32, device=self.device)

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan:
            try:
                _native.lib().dualip_plan_destroy(plan)
            except Exception:
                pass
            self._plan = None

    # -- introspection -------------------------------------------------------------------------------------
    def plan_info(self) -> dict:
        buf = (ctypes.c_int64 * 21)()
        _native.check(_native.lib().dualip_plan_info(self._plan, buf, 21))
        names = ["n_slabs", "n_long_cols", "n_ctas", "threads", "smem_bytes", "row_bits", "smem_mode", "slab_elems",
                 "launches_per_calc", "owned_bytes", "n_slab_cols", "nnz", "fixed_point", "fixed_point_bits",
                 "fixed_point_relerr_e12", "staged_degree", "row_scaled", "n_mid_cols", "grid_tail", "grid_barrier_status",
                 "last_launch_grid_tail"]
        return dict(zip(names, list(buf)))

    def check_grid_barrier(self) -> None:
        """Raises if a grid-wide barrier of the all-CTA tail (csrc/grid_tail.cuh) ever timed out on this plan: the CTAs of a
        launch were not co-resident for seconds, and what that launch produced is invalid.  Reads a word in mapped host memory
        (no synchronisation); call it after the stream has been synchronised."""
        buf = (ctypes.c_int64 * 20)()
        _native.check(_native.lib().dualip_plan_info(self._plan, buf, 20))
        if buf[19]:
            raise RuntimeError("dualip_b200: a grid-wide barrier of the fused kernel timed out (status "
                               f"{int(buf[19])}): its CTAs were not co-resident; results of this run are invalid "
                               "(DUALIP_GRID_TAIL=0 selects the single-CTA tail)")

    def algorithmic_bytes(self, save_primal: bool = False) -> int:
        """B_alg of SURVEY.md §8(d): fp32 a + fp32 c + int32 row per nnz, int32 ccol per column, read lambda and b and
        write grad per dual (+4 B/nnz when the primal is written)."""
        return 12 * self.nnz + 4 * (self.n + 1) + 12 * self.m + (4 * self.nnz if save_primal else 0)

    # -- raw launches (device pointers; used by the Maximizer's fused loop) ---------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def check_inputs_unchanged(self) -> None:
        """Raises if A.values() or c.values() were modified in place since the plan was built (e.g. jacobi_precondition called
        after the constructor): the plan's snapshot would no longer describe the tensors."""
        if (self._a_vals._version, self._c_vals._version) != self._value_versions:
            raise RuntimeError("A.values() or c.values() were modified in place after the objective was built; dualip_b200 keeps a "
                               "snapshot of them in its own layout (apply preprocessing first, or build a new objective)")

    def launched(self) -> None:
        """Self-tuning of the plan, called by `calculate` and by the Maximizer's loop after every evaluation: after 4, 8,
        ..., 128 launches the per-CTA slab ranges are re-cut from the per-CTA times the kernel recorded
        (dualip_plan_rebalance; a stream synchronisation each, six in the life of an objective).  Results are unaffected;
        DUALIP_REBALANCE=0 disables it.  The raw launch_* methods do not call it: a caller that drives several shards of
        one exchange group from ONE host thread must not synchronise one of them while its peers have not launched."""
        self._launches += 1
        if self._launches in self._rebalance_at:
            _native.check(_native.lib().dualip_plan_rebalance(self._plan, self._stream()), "dualip_plan_rebalance")

    def plan_settled(self) -> bool:
        """True once no further re-cut of the plan is due (a CUDA graph captured from then on stays valid)."""
        return not self._rebalance_at or self._launches >= self._rebalance_at[-1]

    def launched_many(self, k: int) -> None:
        """Accounts for k launches replayed from a CUDA graph (only taken when the plan has settled)."""
        self._launches += k

    def launch_calc(self, lam_ptr: int, gamma: float, grad_ptr: int, scal_ptr: int, x_ptr: Optional[int] = None,
                    diag_ptr: Optional[int] = None) -> None:
        b_ptr = self.b_vec.data_ptr() if self.b_vec is not None else None
        rc = _native.lib().dualip_matching_calc(self._plan, lam_ptr, b_ptr, float(gamma), grad_ptr, scal_ptr, x_ptr,
                                                diag_ptr, 0, self._stream())
        _native.check(rc, "dualip_matching_calc")

    def launch_partial(self, lam_ptr: int, gamma: float, partial_ptr: int, x_ptr: Optional[int] = None,
                       diag_ptr: Optional[int] = None) -> None:
        rc = _native.lib().dualip_matching_partial(self._plan, lam_ptr, float(gamma), partial_ptr, x_ptr, diag_ptr, 0,
                                                   self._stream())
        _native.check(rc, "dualip_matching_partial")

    def launch_ascent_step(self, agd_handle, gamma: float, grad_ptr: int, scal_ptr: int, beta: float, decay_now: int,
                           decay_factor: float, iter_index: int, x_ptr: Optional[int] = None) -> None:
        """Evaluation at the optimizer's evaluation point + the accelerated step in ONE launch (dualip_matching_ascent_step)."""
        b_ptr = self.b_vec.data_ptr() if self.b_vec is not None else None
        rc = _native.lib().dualip_matching_ascent_step(self._plan, agd_handle, b_ptr, float(gamma), grad_ptr, scal_ptr, x_ptr,
                                                       float(beta), int(decay_now), float(decay_factor), int(iter_index),
                                                       self._stream())
        _native.check(rc, "dualip_matching_ascent_step")

    def launch_ascent_step_peer(self, agd_handle, peer_handle, b_ptr: int, gamma: float, grad_ptr: int, scal_ptr: int, beta: float,
                                decay_now: int, decay_factor: float, iter_index: int) -> None:
        """Sharded twin: shard kernel, exchange through peer memory, tail and step in ONE launch."""
        rc = _native.lib().dualip_matching_ascent_step_peer(self._plan, agd_handle, peer_handle, b_ptr, float(gamma), grad_ptr,
                                                            scal_ptr, float(beta), int(decay_now), float(decay_factor),
                                                            int(iter_index), self._stream())
        _native.check(rc, "dualip_matching_ascent_step_peer")

    @property
    def has_block_entries(self) -> bool:
        return bool(self._block_entries)

    def add_block_entries(self, lam: torch.Tensor, gamma: float, partial: torch.Tensor, x_out: Optional[torch.Tensor] = None) -> None:
        """Adds the columns of user-registered projections to the packed partial sums [sum_j a_rj x_rj (m) | c.x | ||x||^2]
        that dualip_matching_partial produced for the natively projected columns: v = -(a*lambda + c)/gamma on their
        entries (same rounding as the kernel and the reference, matching.py:130-142), the operator on zero-padded blocks
        per length bucket (matching.py:145-150; a new operator object per call, like the reference), row sums
        (matching.py:153) and the two scalars."""
        s32 = torch.tensor(-1.0 / gamma, dtype=torch.float32, device=self.device)
        lam_s = lam * s32
        v = self._blk_a * lam_s[self._blk_row] + self._blk_c * s32
        x = torch.empty_like(v)
        base = 0
        for ent in self._block_entries:
            fn = project(ent.proj_type, **ent.proj_params)
            for off, total, idx_in_col, cols_rep, L, K in ent.buckets:
                block = torch.zeros((L, K), dtype=torch.float32, device=self.device)
                block[idx_in_col, cols_rep] = v[base + off: base + off + total]
                out = fn(block)
                x[base + off: base + off + total] = out[idx_in_col, cols_rep]
            base += ent.entries.numel()
        partial[: self.m].index_add_(0, self._blk_row, self._blk_a * x)
        partial[self.m] += torch.dot(self._blk_c, x)
        partial[self.m + 1] += torch.dot(x, x)
        if x_out is not None:
            x_out[self._blk_idx] = x

    def _check_dual(self, dual_val: torch.Tensor) -> torch.Tensor:
        if dual_val.device != self.device:
            raise RuntimeError(f"dual_val is on {dual_val.device}, objective on {self.device}")
        if dual_val.dtype != torch.float32:
            raise TypeError("dual_val must be float32")
        if dual_val.numel() != self.m:
            raise ValueError(f"dual_val has {dual_val.numel()} entries, expected {self.m}")
        return dual_val.contiguous()

    def _calculate_host(self, dual_val: torch.Tensor) -> ObjectiveResult:
        """Host-buffer evaluation through dualip_matching_calc_host: lambda is copied host->device from pinned memory,
        grad and the scalars come back device->host, and the stream is synchronised before returning CPU tensors."""
        if dual_val.dtype != torch.float32 or dual_val.numel() != self.m:
            raise ValueError(f"dual_val must be float32 with {self.m} entries")
        if not hasattr(self, "_h_lam"):
            self._h_lam = torch.empty(self.m, dtype=torch.float32).pin_memory()
            self._h_grad = torch.empty(self.m, dtype=torch.float32).pin_memory()
            self._h_scal = torch.empty(len(_native.SCALAR_FIELDS), dtype=torch.float64).pin_memory()
        self._h_lam.copy_(dual_val.reshape(-1))
        with torch.cuda.device(self.device):
            rc = _native.lib().dualip_matching_calc_host(
                self._plan, self._h_lam.data_ptr(), self.b_vec.data_ptr() if self.b_vec is not None else None,
                float(self.gamma), self._h_grad.data_ptr(), self._h_scal.data_ptr(), self._stream())
        _native.check(rc, "dualip_matching_calc_host")
        self.launched()
        return _host_result(self._h_grad.clone(), self._h_scal.clone(), self.is_distributed)

    def host_io_bytes(self) -> tuple:
        """(host->device, device->host) bytes per host-buffer evaluation."""
        return 4 * self.m, 4 * self.m + 8 * len(_native.SCALAR_FIELDS)

    # -- the reference-facing call -------------------------------------------------------------------------
    def calculate(self, dual_val: torch.Tensor, gamma: float = None, save_primal: bool = False, **kwargs) -> ObjectiveResult:
        """Same contract as reference matching.py:116-188.  `diagnostics=True` (extra keyword) additionally returns
        the per-column projection branch / support size as `result.projection_diag` (uint8 per nnz position)."""
        if gamma is not None and gamma != self.gamma:
            self.gamma = gamma  # no O(E) rescaling pass: the kernel forms -(a*lambda + c)/gamma in registers
        self.check_inputs_unchanged()
        if isinstance(dual_val, torch.Tensor) and dual_val.device.type == "cpu":
            if save_primal or kwargs.get("diagnostics"):
                raise ValueError("save_primal / diagnostics need a device-resident dual_val")
            if self._block_entries:
                r = self.calculate(dual_val.to(self.device), gamma=None, save_primal=False)
                return _host_result(r.dual_gradient.cpu(), r.scalars64.cpu(), self.is_distributed)
            return self._calculate_host(dual_val)
        lam = self._check_dual(dual_val)
        with torch.cuda.device(self.device):
            grad = torch.empty(self.m, dtype=torch.float32, device=self.device)
            scal = torch.empty(len(_native.SCALAR_FIELDS), dtype=torch.float64, device=self.device)
            x = torch.empty(self.nnz, dtype=torch.float32, device=self.device) if save_primal else None
            diag = None
            if kwargs.get("diagnostics"):
                diag = torch.full((self.nnz,), 255, dtype=torch.uint8, device=self.device)
            if self._block_entries:
                # natively projected columns in the fused kernel, the others through padded blocks, then the m-length tail
                partial = torch.empty(self.m + 2, dtype=torch.float32, device=self.device)
                self.launch_partial(lam.data_ptr(), self.gamma, partial.data_ptr(), x.data_ptr() if x is not None else None,
                                    diag.data_ptr() if diag is not None else None)
                self.add_block_entries(lam, self.gamma, partial, x)
                rc = _native.lib().dualip_matching_epilogue(
                    partial.data_ptr(), self.m, lam.data_ptr(), self.b_vec.data_ptr() if self.b_vec is not None else None,
                    float(self.gamma), grad.data_ptr(), scal.data_ptr(), self._stream())
                _native.check(rc, "dualip_matching_epilogue")
            else:
                self.launch_calc(lam.data_ptr(), self.gamma, grad.data_ptr(), scal.data_ptr(),
                                 x.data_ptr() if x is not None else None, diag.data_ptr() if diag is not None else None)
            s32 = scal.to(torch.float32)
            self.launched()
        if not self.is_distributed:
            res = ObjectiveResult(
                dual_gradient=grad,
                dual_objective=s32[_IDX["dual_objective"]],
                reg_penalty=s32[_IDX["reg_penalty"]],
                dual_val_times_grad=s32[_IDX["dual_val_times_grad"]],
                max_pos_slack=s32[_IDX["max_pos_slack"]],
                sum_pos_slack=s32[_IDX["sum_pos_slack"]],
            )
        else:
            # local-shard mode (matching.py:179-184): raw partial gradient, dual_objective = c.x
            res = ObjectiveResult(
                dual_gradient=grad,
                dual_objective=s32[_IDX["primal_objective"]],
                reg_penalty=s32[_IDX["reg_penalty"]],
            )
        if save_primal:
            res.primal_var = x  # a fresh buffer (the reference aliases its scratch, matching.py:156,186)
            res.primal_objective = s32[_IDX["primal_objective"]].clone()
        if diag is not None:
            res.projection_diag = diag
        res.scalars64 = scal
        return res


def _host_result(grad: torch.Tensor, scal: torch.Tensor, local_mode: bool) -> ObjectiveResult:
    s32 = scal.to(torch.float32)
    if local_mode:
        res = ObjectiveResult(dual_gradient=grad, dual_objective=s32[_IDX["primal_objective"]],
                              reg_penalty=s32[_IDX["reg_penalty"]])
    else:
        res = ObjectiveResult(
            dual_gradient=grad,
            dual_objective=s32[_IDX["dual_objective"]],
            reg_penalty=s32[_IDX["reg_penalty"]],
            dual_val_times_grad=s32[_IDX["dual_val_times_grad"]],
            max_pos_slack=s32[_IDX["max_pos_slack"]],
            sum_pos_slack=s32[_IDX["sum_pos_slack"]],
        )
    res.scalars64 = scal
    return res


def reduce_partials(partial: torch.Tensor) -> torch.Tensor:
    """The one collective of the sharded path: SUM all-reduce of the packed [grad(m) | c.x | ||x||^2] vector
    (replaces three dist.reduce + barrier, reference matching.py:272-277).  No-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial


class MatchingSolverDualObjectiveFunctionDistributed(BaseObjective):
    """Entity-sharded matching objective: one process per GPU, each with its column shard.

    Same constructor as the reference (matching.py:218-245).  Per call: local fused kernel -> ONE all-reduce of
    m+2 floats -> m-length tail on every rank.  Every rank returns the full result (a superset of the reference,
    where only rank 0's is meaningful, matching.py:280-307), so the Maximizer can update redundantly and skip the
    two broadcasts of agd.py:204-206."""

    result_on_all_ranks = True

    def __init__(self, local_matching_input_args: MatchingInputArgs, b_vec: torch.Tensor, gamma: float,
                 host_device=None, batching: bool = True):
        if local_matching_input_args.b_vec is not None:
            raise ValueError("local_matching_input_args.b_vec must be None: b_vec is shared across ranks")
        self.gamma = gamma
        self.host_device = host_device
        self.equality_mask = local_matching_input_args.equality_mask
        self.local_objective = MatchingSolverDualObjectiveFunction(local_matching_input_args, gamma, batching)
        self.device = self.local_objective.device
        self.m = self.local_objective.m
        self.b_vec = b_vec.to(device=self.device, dtype=torch.float32).contiguous()
        self._peer = None
        self._peer_tried = False

    def peer_exchange(self):
        """The NVLink exchange windows of this objective (dualip_b200.utils.peer_exchange), set up on first use.
        COLLECTIVE: every rank must call it at the same point.  None when the ranks cannot map each other's memory (not
        one host, not NCCL, DUALIP_PEER_EXCHANGE=0); the fused loop then uses the NCCL all-reduce."""
        if getattr(self, "_peer_failed", False):
            raise RuntimeError("the exchange windows of this objective timed out in an earlier run; build a new objective "
                               "(or set DUALIP_PEER_EXCHANGE=0 for the NCCL path)")
        if not self._peer_tried:
            from dualip_b200.utils.peer_exchange import PeerExchange

            self._peer_tried = True
            # COLLECTIVE even when this rank cannot take part, so that every rank reaches the same decision
            want = not self.local_objective.has_block_entries  # block entries add their sums with tensor ops
            self._peer = PeerExchange.over_process_group(self.m, self.device, enabled=want)
        return self._peer

    def host_io_bytes(self) -> tuple:
        return 4 * self.m, 4 * self.m + 8 * len(_native.SCALAR_FIELDS)

    def launch_partial_and_reduce(self, lam_ptr: int, gamma: float, partial: torch.Tensor, lam: Optional[torch.Tensor] = None) -> None:
        self.local_objective.launch_partial(lam_ptr, gamma, partial.data_ptr())
        if self.local_objective.has_block_entries:
            self.local_objective.add_block_entries(lam, gamma, partial)
        reduce_partials(partial)

    def launch_epilogue(self, partial_ptr: int, lam_ptr: int, gamma: float, grad_ptr: int, scal_ptr: int) -> None:
        rc = _native.lib().dualip_matching_epilogue(partial_ptr, self.m, lam_ptr, self.b_vec.data_ptr(), float(gamma),
                                                    grad_ptr, scal_ptr, self.local_objective._stream())
        _native.check(rc, "dualip_matching_epilogue")

    def calculate(self, dual_val: torch.Tensor, gamma: float = None, save_primal: bool = False, rank: int = 0) -> ObjectiveResult:
        if save_primal:
            raise NotImplementedError("save_primal=True is not yet supported in distributed mode")  # matching.py:255-256
        if gamma is not None and gamma != self.gamma:
            self.gamma = gamma
            self.local_objective.gamma = gamma
        host_io = isinstance(dual_val, torch.Tensor) and dual_val.device.type == "cpu"
        if host_io:
            if not hasattr(self, "_h_lam"):
                self._h_lam = torch.empty(self.m, dtype=torch.float32).pin_memory()
                self._h_grad = torch.empty(self.m, dtype=torch.float32).pin_memory()
                self._h_scal = torch.empty(len(_native.SCALAR_FIELDS), dtype=torch.float64).pin_memory()
                self._d_lam = torch.empty(self.m, dtype=torch.float32, device=self.device)
            if dual_val.dtype != torch.float32 or dual_val.numel() != self.m:
                raise ValueError(f"dual_val must be float32 with {self.m} entries")
            self._h_lam.copy_(dual_val.reshape(-1))
            peer = self.peer_exchange()  # COLLECTIVE on first use: every rank evaluates in lockstep anyway
            if peer is not None:
                # lambda host->device, shard kernel whose last CTA exchanges the sums through peer memory and runs the m-length
                # tail, grad + scalars device->host: ONE native call, no collective call, no tensor allocation on the device
                self.local_objective.check_inputs_unchanged()
                with torch.cuda.device(self.device):
                    rc = _native.lib().dualip_matching_calc_peer_host(
                        self.local_objective._plan, peer.handle, self._h_lam.data_ptr(), self.b_vec.data_ptr(), float(self.gamma),
                        self._h_grad.data_ptr(), self._h_scal.data_ptr(), self.local_objective._stream())
                if rc != _native.OK and peer.status_nowait():
                    self._peer_failed, self._peer = True, None
                _native.check(rc, "dualip_matching_calc_peer_host")
                self.local_objective.launched()
                return _host_result(self._h_grad.clone(), self._h_scal.clone(), False)
            self._d_lam.copy_(self._h_lam, non_blocking=True)
            dual_val = self._d_lam
        lam = self.local_objective._check_dual(dual_val)
        with torch.cuda.device(self.device):
            partial = torch.empty(self.m + 2, dtype=torch.float32, device=self.device)
            grad = torch.empty(self.m, dtype=torch.float32, device=self.device)
            scal = torch.empty(len(_native.SCALAR_FIELDS), dtype=torch.float64, device=self.device)
            self.launch_partial_and_reduce(lam.data_ptr(), self.gamma, partial, lam)
            self.launch_epilogue(partial.data_ptr(), lam.data_ptr(), self.gamma, grad.data_ptr(), scal.data_ptr())
            self.local_objective.launched()
            if host_io:
                self._h_grad.copy_(grad, non_blocking=True)
                self._h_scal.copy_(scal, non_blocking=True)
                torch.cuda.current_stream(self.device).synchronize()
                return _host_result(self._h_grad.clone(), self._h_scal.clone(), False)
            s32 = scal.to(torch.float32)
        res = ObjectiveResult(
            dual_gradient=grad,
            dual_objective=s32[_IDX["dual_objective"]],
            reg_penalty=s32[_IDX["reg_penalty"]],
            dual_val_times_grad=s32[_IDX["dual_val_times_grad"]],
            max_pos_slack=s32[_IDX["max_pos_slack"]],
            sum_pos_slack=s32[_IDX["sum_pos_slack"]],
        )
        res.scalars64 = scal
        return res
