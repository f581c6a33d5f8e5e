"""CPU restatement (numpy) of the DuaLip dual-ascent hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``dualip_b200/`` may import this module: it is the checker for the CUDA path, used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Every function cites the reference file:line it restates (paths relative to the reference tree,
linkedin/DuaLip v5.0.1 @ 240066ec).  Arithmetic follows the reference's *CPU PyTorch* semantics:
float32 element-wise operations in the reference's order, ``cumsum`` accumulated in float64 and rounded
per element (torch's CPU cumsum uses acc_type<float> = double), Python-float scalars rounded to float32
before they meet a float32 tensor.  ``dtype=np.float64`` turns the same code into the fp64 tie-breaker.

Pinning (see oracle/PINNING.md): checked against the reference's golden vectors
(tests/objectives/test_dualip_matching_simplex.py:129-141, tests/test_agd.py:95-107,
tests/projections/test_simplex.py:270-284) in tests/test_oracle_golden.py, and against outputs of the
reference itself generated in the build container (tests/golden/*.npz, made by tests/golden/make_golden.py).

Deliberate difference from the reference: with a projection map of several entries the reference's
apply_F_to_columns (utils/sparse_utils.py:177,220) overwrites the columns of earlier entries with
uninitialised memory (``torch.empty_like`` + whole-array ``copy_``).  The oracle implements the evident
intent: each entry projects its own columns and leaves the others untouched.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BRANCH_FEASIBLE, BRANCH_SHORTCUT, BRANCH_DUCHI = 0, 1, 2


# --------------------------------------------------------------------------------------
# projections  (reference src/dualip/projections/)
# --------------------------------------------------------------------------------------
def box_proj(x: np.ndarray, lower: float = 0.0, upper: float = 1.0) -> np.ndarray:
    """projections/box.py:15-16: x.clamp(min=lower, max=upper)."""
    dt = x.dtype
    return np.minimum(np.maximum(x, dt.type(lower)), dt.type(upper))


def cone_proj(x: np.ndarray, lower: Optional[float] = None, upper: Optional[float] = None) -> np.ndarray:
    """projections/cone.py:15-28."""
    if lower is not None and upper is not None:
        raise ValueError("Only one of 'lower' or 'upper' should be specified, not both.")
    dt = x.dtype
    if lower is not None:
        return np.maximum(x, dt.type(lower))
    if upper is not None:
        return np.minimum(x, dt.type(upper))
    return x.copy()


def duchi_proj(x: np.ndarray, z: float, inequality: bool, tol: float = 1e-6):
    """projections/simplex.py:126-236 (`_duchi_proj`) on a zero-padded [L, K] block.

    Returns (w, branch[K], rho[K]); branch/rho are the "projection index selection".
    The 10 000-column chunking (:202) does not change any value and is dropped.
    """
    x = np.asarray(x)
    dt = x.dtype
    L, K = x.shape
    assert z > 0, "Simplex radius z must be positive."
    zt = dt.type(z)
    w = np.empty_like(x)
    x = np.maximum(x, dt.type(0.0))  # :148
    branch = np.full(K, BRANCH_DUCHI, dtype=np.int8)
    rho_out = np.zeros(K, dtype=np.int64)
    to_project = np.ones(K, dtype=bool)
    if inequality:
        # :153-159. `z + tol` is a Python float; it is rounded to the tensor dtype for the comparison.
        s = np.zeros(K, dtype=dt)
        for i in range(L):  # row-by-row accumulation in the working dtype
            s = s + x[i]
        feasible = s <= dt.type(float(z) + tol)
        w[:, feasible] = x[:, feasible]
        branch[feasible] = BRANCH_FEASIBLE
        to_project = ~feasible
    if L > 1 and to_project.any():  # :166-193
        idx = np.nonzero(to_project)[0]
        xn = x[:, idx] / zt
        order = np.argsort(-xn, axis=0, kind="stable")
        top0 = np.take_along_axis(xn, order[0:1], axis=0)[0]
        top1 = np.take_along_axis(xn, order[1:2], axis=0)[0]
        short = (top0 - top1) > dt.type(1.0)
        if short.any():
            cols = idx[short]
            sol = np.zeros((L, cols.size), dtype=dt)
            sol[order[0, short], np.arange(cols.size)] = zt
            w[:, cols] = sol
            branch[cols] = BRANCH_SHORTCUT
            rho_out[cols] = 1
            to_project[cols] = False
    if to_project.any():  # :199-234
        idx = np.nonzero(to_project)[0]
        sub = x[:, idx]
        u_sorted = -np.sort(-sub, axis=0)
        css = np.cumsum(u_sorted.astype(np.float64), axis=0).astype(dt)  # double accumulator, rounded per element
        i_f = np.arange(1, L + 1, dtype=dt).reshape(L, 1)
        cond = (u_sorted - (css - zt) / i_f) > 0
        rho0 = (cond.astype(np.int64) * np.arange(L).reshape(L, 1)).max(axis=0)
        css_rho = css[rho0, np.arange(idx.size)]
        theta = (css_rho - zt) / (rho0.astype(dt) + dt.type(1.0))
        w[:, idx] = np.maximum(sub - theta[None, :], dt.type(0.0))
        rho_out[idx] = rho0 + 1
    return w, branch, rho_out


def bisection_proj(x: np.ndarray, z: float, inequality: bool, tol: float = 1e-6, max_iter: int = 50) -> np.ndarray:
    """projections/simplex.py:6-123 (`_proj_via_bisection_search`) on a zero-padded [L, K] block.

    Differences from duchi_proj that are part of the reference and therefore kept: no pre-clamp (negative entries count in
    the column sum, and a column is only "feasible" if all of its entries are >= -tol, :40); the shift subtracts the
    maximum of x/z from the un-normalised x (:86-89) and the root is searched for a sum of 1, not z (:104) -- exact for
    z = 1 only; the search interval [-1, 0] is halved for every active column at once, so all columns stop after the
    same number of steps (:95-118)."""
    x = np.asarray(x)
    dt = x.dtype
    L, K = x.shape
    assert z > 0, "Simplex radius z must be positive."
    zt = dt.type(z)
    w = np.empty_like(x)
    todo = np.ones(K, dtype=bool)
    if inequality:  # :39-47
        s = np.zeros(K, dtype=dt)
        for i in range(L):
            s = s + x[i]
        feasible = (s <= dt.type(float(z) + tol)) & (x >= dt.type(-tol)).all(axis=0)
        w[:, feasible] = x[:, feasible]
        todo = ~feasible
    if L > 1 and todo.any():  # :52-81, top-2 shortcut on x/z
        idx = np.nonzero(todo)[0]
        xn = x[:, idx] / zt
        order = np.argsort(-xn, axis=0, kind="stable")
        top0 = np.take_along_axis(xn, order[0:1], axis=0)[0]
        top1 = np.take_along_axis(xn, order[1:2], axis=0)[0]
        short = (top0 - top1) > dt.type(1.0)
        if short.any():
            cols = idx[short]
            sol = np.zeros((L, cols.size), dtype=dt)
            sol[order[0, short], np.arange(cols.size)] = zt
            w[:, cols] = sol
            todo[cols] = False
    if not todo.any():
        return w
    idx = np.nonzero(todo)[0]  # :85-122
    sub = x[:, idx]
    shifted = sub - (sub / zt).max(axis=0)[None, :]
    lo = np.full(idx.size, -1.0, dtype=dt)
    hi = np.zeros(idx.size, dtype=dt)
    active = np.ones(idx.size, dtype=bool)
    prev = None
    for _ in range(max_iter):
        if not active.any():
            break
        mid = (lo + hi) / dt.type(2.0)
        if prev is not None and np.abs(mid - prev).max() < tol:
            break
        ssum = np.zeros(idx.size, dtype=dt)
        t = np.maximum(shifted - mid[None, :], dt.type(0.0))
        for i in range(L):
            ssum = ssum + t[i]
        high = ssum > dt.type(1.0)
        lo = np.where(high & active, mid, lo)
        hi = np.where(~high & active, mid, hi)
        active = active & ~((hi - lo) < tol)
        prev = mid.copy()
    nu = (lo + hi) / dt.type(2.0)
    w[:, idx] = np.maximum(shifted - nu[None, :], dt.type(0.0)) * zt
    return w


# --------------------------------------------------------------------------------------
# projection map handling  (reference projections/base.py, objectives/matching.py:70-114)
# --------------------------------------------------------------------------------------
@dataclass
class ProjEntry:
    """Mirror of projections/base.py:8-12."""

    proj_type: str = ""
    proj_params: dict = field(default_factory=dict)
    indices: Sequence[int] = field(default_factory=list)


def compute_buckets(ccol: np.ndarray, n_rows: int, indices: np.ndarray, batching: bool = True) -> List[np.ndarray]:
    """objectives/matching.py:87-114.  Thresholds [0,2,4,...,2^k<=m, m+1], torch.bucketize (right-closed),
    empty columns (bucket 0) dropped.  batching=False: one bucket with every listed column (:76-77)."""
    indices = np.asarray(indices, dtype=np.int64)
    if not batching:
        return [indices]
    th = [0]
    i = 1
    while 2**i <= n_rows:
        th.append(2**i)
        i += 1
    th.append(n_rows + 1)
    lengths = np.diff(ccol)
    bucket_ids = np.searchsorted(np.asarray(th), lengths, side="left")  # torch.bucketize(right=False)
    pb = bucket_ids[indices]
    out = []
    for j in range(1, len(th)):
        sel = indices[pb == j]
        if sel.size:
            out.append(sel)
    return out


def pad_table(ccol: np.ndarray, n_rows: int, entries: Sequence[ProjEntry], batching: bool = True) -> np.ndarray:
    """Padded block length L per (entry, ceil(log2(column length))) from the reference's bucket rule
    (objectives/matching.py:87-114 + utils/sparse_utils.py:197): int32 [len(entries), 32].  For checkers that work
    per column (oracle/matching_oracle.c) instead of per padded block."""
    lengths = np.diff(np.asarray(ccol, dtype=np.int64))
    out = np.zeros((len(entries), 32), dtype=np.int32)
    for e, entry in enumerate(entries):
        for cols in compute_buckets(ccol, n_rows, np.asarray(entry.indices, dtype=np.int64), batching):
            ln = lengths[cols]
            ln = ln[ln > 0]
            if ln.size == 0:
                continue
            L = int(ln.max())
            for d in np.unique(ln):
                bkt = 0
                while (1 << bkt) < int(d):
                    bkt += 1
                out[e, bkt] = max(out[e, bkt], L)
    return out


def _apply_entry(vals: np.ndarray, ccol: np.ndarray, entry: ProjEntry, n_rows: int, batching: bool, branch, rho):
    """utils/sparse_utils.py:133-220 (`apply_F_to_columns`): padded [L x K] block per bucket."""
    dt = vals.dtype
    for cols in compute_buckets(ccol, n_rows, np.asarray(entry.indices, dtype=np.int64), batching):
        starts = ccol[cols]
        lengths = ccol[cols + 1] - starts
        total = int(lengths.sum())
        if total == 0:
            continue
        L = int(lengths.max())
        K = cols.size
        cols_rep = np.repeat(np.arange(K), lengths)
        prefix = np.cumsum(lengths) - lengths
        idx_in_col = np.arange(total) - prefix[cols_rep]
        flat = starts[cols_rep] + idx_in_col
        block = np.zeros((L, K), dtype=dt)
        block[idx_in_col, cols_rep] = vals[flat]
        pt, pp = entry.proj_type, entry.proj_params
        if pt == "box":
            out = box_proj(block, **pp)
        elif pt == "cone":
            out = cone_proj(block, **pp)
        elif pt in ("simplex", "simplex_eq"):
            method = pp.get("method", "duchi")
            if method == "bisection_search":
                out = bisection_proj(block, float(pp.get("z", 1.0)), inequality=(pt == "simplex"))
            elif method == "duchi":
                out, br, rh = duchi_proj(block, float(pp.get("z", 1.0)), inequality=(pt == "simplex"))
                nz = lengths > 0
                branch[cols[nz]] = br[nz]
                rho[cols[nz]] = rh[nz]
            else:
                raise ValueError(f"Unsupported projection method: {method}")
        else:
            raise ValueError(f"Unknown projection operator '{pt}'")
        vals[flat] = out[idx_in_col, cols_rep]


@dataclass
class OracleResult:
    dual_gradient: np.ndarray
    dual_objective: float
    reg_penalty: float
    primal_objective: float
    primal_var: np.ndarray
    dual_val_times_grad: Optional[float] = None
    max_pos_slack: Optional[float] = None
    sum_pos_slack: Optional[float] = None
    branch: Optional[np.ndarray] = None  # per column: 0 feasible / 1 shortcut / 2 Duchi / -1 not simplex or empty
    rho: Optional[np.ndarray] = None  # per column support size for branch 1/2


def matching_calculate(
    ccol: np.ndarray,
    row: np.ndarray,
    a: np.ndarray,
    c: np.ndarray,
    n_rows: int,
    projection_map: Dict[str, ProjEntry],
    lam: np.ndarray,
    gamma: float,
    b: Optional[np.ndarray] = None,
    batching: bool = True,
    dtype=np.float32,
) -> OracleResult:
    """objectives/matching.py:116-188 (`calculate`), single device.

    b=None is the "local shard" mode (:56,:179-184): raw partial gradient, dual_objective = c.x only.
    Scalar reductions are returned as Python floats computed in float64 from the working-dtype x
    (the reference's own float32 reductions are order-dependent; the 1e-5 gate is on these values).
    """
    dt = np.dtype(dtype)
    ccol = np.asarray(ccol, dtype=np.int64)
    row = np.asarray(row, dtype=np.int64)
    a = np.asarray(a, dtype=dt)
    c = np.asarray(c, dtype=dt)
    lam = np.asarray(lam, dtype=dt)
    s = dt.type(-1.0 / gamma)  # Python double, rounded when it meets the tensor (:133,:136)
    c_rescaled = s * c  # :66,:133
    scaled = s * lam  # :136
    vals = a * scaled[row]  # :139 -> sparse_utils.py:79
    vals = vals + c_rescaled  # :142
    n = ccol.size - 1
    branch = np.full(n, -1, dtype=np.int8)
    rho = np.zeros(n, dtype=np.int64)
    for _, entry in projection_map.items():  # :145-150
        _apply_entry(vals, ccol, entry, n_rows, batching, branch, rho)
    prod = a * vals  # :153
    grad = np.zeros(n_rows, dtype=np.float64)
    np.add.at(grad, row, prod.astype(np.float64))  # sparse_utils.py:240-242 (order-free in fp64)
    grad = grad.astype(dt)
    xx = float(np.dot(vals.astype(np.float64), vals.astype(np.float64)))
    reg = (gamma / 2.0) * xx  # :157
    cx = float(np.dot(c.astype(np.float64), vals.astype(np.float64)))  # :160
    res = OracleResult(grad, cx, reg, cx, vals, branch=branch, rho=rho)
    if b is not None:
        g = grad - np.asarray(b, dtype=dt)  # :32
        lg = float(np.dot(lam.astype(np.float64), g.astype(np.float64)))
        res.dual_gradient = g
        res.dual_objective = cx + reg + lg  # :33
        res.dual_val_times_grad = lg  # :167
        res.max_pos_slack = float(max(g.max(), 0.0)) if g.size else 0.0  # :168
        res.sum_pos_slack = float(np.maximum(g, 0).astype(np.float64).sum())  # :169
    return res


# --------------------------------------------------------------------------------------
# Maximizer  (reference src/dualip/optimizers/agd.py, agd_utils.py)
# --------------------------------------------------------------------------------------
def compute_beta_seq(max_iter: int) -> np.ndarray:
    """optimizers/agd.py:93-100: t stored as float32, sqrt evaluated in Python double."""
    t = np.zeros(max_iter + 2, dtype=np.float32)
    beta = np.zeros(max_iter, dtype=np.float32)
    for i in range(1, max_iter + 2):
        inner = np.float32(1) + np.float32(4) * (t[i - 1] * t[i - 1])  # float32 tensor arithmetic
        t[i] = np.float32((1 + math.sqrt(float(inner))) / 2)
    for i in range(max_iter):
        beta[i] = (np.float32(1) - t[i + 1]) / t[i + 2]
    return beta


def calculate_step_size(grad, dual, grad_hist: list, dual_hist: list, max_history_length=15, initial_step_size=1e-5,
                        max_step_size=0.1) -> float:
    """optimizers/agd_utils.py:65-89 with :11-62."""
    if len(grad_hist) == max_history_length:
        dual_hist.pop(0)
        grad_hist.pop(0)
    grad_hist.append(np.array(grad, copy=True))
    dual_hist.append(np.array(dual, copy=True))
    ls = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(len(grad_hist) - 1):
            dg = np.float32(np.linalg.norm((grad_hist[i] - grad_hist[i + 1]).astype(np.float64)))
            dy = np.float32(np.linalg.norm((dual_hist[i] - dual_hist[i + 1]).astype(np.float64)))
            ls.append(np.float32(dg) / np.float32(dy))
    if not ls or len(ls) < max_history_length - 1:
        return initial_step_size
    l_max = ls[0]  # Python max(): first element wins ties / NaN comparisons are False
    for v in ls[1:]:
        if v > l_max:
            l_max = v
    if np.isnan(l_max) or np.isinf(l_max):
        return initial_step_size
    cand = 1.0 / float(l_max) if l_max != 0 else max_step_size
    return min(cand, max_step_size)


def project_on_nn_cone(y: np.ndarray, equality_mask: Optional[np.ndarray]) -> np.ndarray:
    """optimizers/agd.py:13-21."""
    p = np.maximum(y, y.dtype.type(0))
    return np.where(equality_mask, y, p) if equality_mask is not None else p


def agd_maximize(calc, initial_value: np.ndarray, max_iter: int, gamma: Optional[float], initial_step_size=1e-5,
                 max_step_size=0.1, gamma_decay_type=None, gamma_decay_params=None, equality_mask=None):
    """optimizers/agd.py:121-229, rank-0 path.  `calc(lam, gamma)` -> (grad ndarray, dual_obj float).

    Returns (y, dual_obj_log, step_size_log, final gamma).  History stores y, not x (:170-172)."""
    dt = initial_value.dtype
    beta = compute_beta_seq(max_iter)
    x = initial_value.copy()
    y = initial_value.copy()
    gh, dh, obj_log, step_log = [], [], [], []
    for i in range(1, max_iter + 1):
        grad, obj = calc(x, gamma)
        obj_log.append(float(obj))
        step = calculate_step_size(grad, y, gh, dh, initial_step_size=initial_step_size, max_step_size=max_step_size)
        step_log.append(step)
        y_new = x + grad * dt.type(step)  # :181
        y_new = project_on_nn_cone(y_new, equality_mask)
        bi = beta[i - 1].astype(dt)
        x = (y_new * (dt.type(1.0) - bi)) + (y * bi)  # :184
        y = y_new
        if gamma is not None and gamma_decay_type is not None:  # :186-187 -> :102-109
            if gamma_decay_type != "step":
                raise ValueError(f"Unsupported gamma decay type: {gamma_decay_type}")
            if i % gamma_decay_params["decay_steps"] == 0:
                f = gamma_decay_params["decay_factor"]
                gamma = gamma * f
                max_step_size = step * f
    return y, obj_log, step_log, gamma


# --------------------------------------------------------------------------------------
# helpers shared by tests and bench
# --------------------------------------------------------------------------------------
def split_columns(n_cols: int, world: int) -> List[Tuple[int, int]]:
    """utils/dist_utils.py:53-61: contiguous column ranges, n//W each, first n%W get one more."""
    base, rem = divmod(n_cols, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def jacobi_precondition(a: np.ndarray, row: np.ndarray, b: np.ndarray, n_rows: int):
    """preprocessing/precondition.py:8-29 + sparse_utils.py:429-450 (returns scaled copies and norms)."""
    dt = a.dtype
    sq = np.zeros(n_rows, dtype=np.float64)
    np.add.at(sq, row, (a * a).astype(np.float64))
    norms = np.sqrt(sq.astype(dt))
    rec = dt.type(1) / norms
    return a * rec[row], b * rec, norms


# --------------------------------------------------------------------------------------
# generic (non-block) LP objective: objectives/miplib.py:60-109
# --------------------------------------------------------------------------------------
def lp_calculate(A: np.ndarray, c: np.ndarray, b: np.ndarray, lower: np.ndarray, upper: np.ndarray, lam: np.ndarray,
                 gamma: float, row_norms: Optional[np.ndarray] = None, dtype=np.float32):
    """Dense restatement of MIPLIB2017ObjectiveFunction.calculate.  lower/upper: per-variable clamp bounds (-inf/+inf =
    open), i.e. the box / cone entries of the projection map applied element-wise (miplib.py:80-90).

    Returns (grad, dual_obj, reg, x, primal_obj)."""
    A = np.asarray(A, dtype=dtype)
    c, b, lam = np.asarray(c, dtype=dtype), np.asarray(b, dtype=dtype), np.asarray(lam, dtype=dtype)
    if row_norms is not None:
        lam = (dtype(1.0) / np.asarray(row_norms, dtype=dtype)) * lam                      # :73-74
    z = dtype(-1.0 / gamma) * (A.T @ lam + c)                                              # :76
    x = np.minimum(np.maximum(z, np.asarray(lower, dtype=dtype)), np.asarray(upper, dtype=dtype))
    resid = A @ x - b
    grad = resid if row_norms is None else (dtype(1.0) / np.asarray(row_norms, dtype=dtype)) * resid   # :92-95
    reg = dtype(gamma / 2.0) * dtype(np.linalg.norm(x.astype(np.float64))) ** 2            # :97
    primal = float(c.astype(np.float64) @ x.astype(np.float64))
    dual_obj = primal + float(reg) + float(lam.astype(np.float64) @ resid.astype(np.float64))   # :99
    return grad, dual_obj, float(reg), x, primal
