"""Round-2 fixtures, produced by the UNMODIFIED reference (oracle/_ref = pip-installed linkedin/DuaLip v5.0.1, or
/root/reference) in the build container.  The GPU box has no reference tree, so the vectors are committed.

    python tests/golden/make_golden_r2.py

* mixed_c3shape.npz — configs[2]'s projection map (simplex z=1 on even entities, box[0,1] on odd ones) on a C3-shaped
  problem from the reference's own generator, Jacobi-preconditioned by the reference, at a LATE dual (60 reference AGD
  iterations).  The reference corrupts a projection map with several entries (utils/sparse_utils.py:177,220), so its
  `calculate` is run once per entry on that entry's column sub-matrix (reference `split`-style column selection) and the
  results are combined: x interleaved, gradients / c.x / ||x||^2 summed.  batching True and False.
* case_simplex_eq.npz — `simplex_eq` through the reference objective, batching on/off.  The result depends on the
  padded length of the bucket a column lands in: columns whose clamped sum is below z get
  (z - sum)/L_bucket added to every entry, where L_bucket is the longest column of the bucket.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def _ref():
    from oracle import make_ref

    make_ref.make()
    make_ref.import_reference()


def _calc(ccol, row, a, c, m, pm, b, lam, gamma, batching):
    import torch
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction

    n = ccol.size - 1
    A = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(a), size=(m, n))
    C = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(c), size=(m, n))
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, None if b is None else torch.from_numpy(b)),
                                              gamma=gamma, batching=batching)
    return obj.calculate(torch.from_numpy(lam), save_primal=True)


def _select_columns(ccol, row, a, c, cols):
    lens = np.diff(ccol)[cols]
    sub_ccol = np.zeros(cols.size + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_ccol[1:])
    pos = np.concatenate([np.arange(ccol[j], ccol[j + 1]) for j in cols] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
    return sub_ccol, row[pos], a[pos], c[pos], pos


def make_mixed():
    import torch
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.optimizers.agd import AcceleratedGradientDescent
    from dualip.preprocessing.precondition import jacobi_precondition
    from dualip.projections.base import create_projection_map

    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "reference_benchmark"))
    from generate_synthetic_data import generate_synthetic_matching_input_args as gen

    n, m, sparsity, gamma = 24_000, 250, 0.04, 1e-3  # mean degree 10 like C3 (sparsity * m)
    args = gen(n, m, sparsity, rng=np.random.default_rng(42))  # rng given: no disk cache
    A, C, b = args.A, args.c, args.b_vec
    jacobi_precondition(A, b)  # in place, like benchmark/benchmark_utils.py:54-56
    ccol = A.ccol_indices().numpy().astype(np.int64)
    row = A.row_indices().numpy().astype(np.int64)
    a = A.values().numpy().astype(np.float32).copy()
    c = C.values().numpy().astype(np.float32).copy()
    b = b.numpy().astype(np.float32).copy()
    # a late dual: 60 iterations of the reference's maximizer with the single-entry simplex map (valid in the reference)
    pm_all = create_projection_map("simplex", {"z": 1.0}, n)
    obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm_all, torch.from_numpy(b)), gamma=gamma)
    solver = AcceleratedGradientDescent(max_iter=60, gamma=gamma, initial_step_size=1e-3, max_step_size=1e-1,
                                        iteration_callback=lambda i, r: None)
    lam = solver.maximize(obj, torch.zeros(m)).dual_val.numpy().astype(np.float32).copy()
    even, odd = np.arange(0, n, 2), np.arange(1, n, 2)
    out = dict(ccol=ccol, row=row, a=a, c=c, b=b, lam=lam, gamma=np.float64(gamma), n_rows=np.int64(m))
    for batching in (True, False):
        tag = "b1" if batching else "b0"
        x = np.zeros(row.size, dtype=np.float32)
        grad = np.zeros(m, dtype=np.float64)
        cx = xx = 0.0
        for cols, ptype, params in ((even, "simplex", {"z": 1.0}), (odd, "box", {"lower": 0.0, "upper": 1.0})):
            sc, sr, sa, scv, pos = _select_columns(ccol, row, a, c, cols)
            r = _calc(sc, sr, sa, scv, m, create_projection_map(ptype, params, cols.size), None, lam, gamma, batching)
            xs = r.primal_var.numpy()
            x[pos] = xs
            grad += r.dual_gradient.numpy().astype(np.float64)  # b_vec=None: raw partial row sums (matching.py:179-184)
            cx += float(np.dot(scv.astype(np.float64), xs.astype(np.float64)))
            xx += float(np.dot(xs.astype(np.float64), xs.astype(np.float64)))
        g = grad.astype(np.float32) - b
        lg = float(np.dot(lam.astype(np.float64), g.astype(np.float64)))
        out[f"x_{tag}"] = x
        out[f"grad_{tag}"] = g
        out[f"scal_{tag}"] = np.array([cx + gamma / 2 * xx + lg, gamma / 2 * xx, cx, lg, max(float(g.max()), 0.0),
                                       float(np.maximum(g, 0).astype(np.float64).sum())])
    print("mixed: nnz", row.size, "simplex columns with sum>z:",
          "x sums", float(out["x_b1"].sum()))
    np.savez_compressed(os.path.join(HERE, "mixed_c3shape.npz"), **out)


def make_simplex_eq():
    from dualip.projections.base import create_projection_map

    rng = np.random.default_rng(20261017)
    n, m, gamma = 3000, 48, 5e-2
    deg = np.clip(rng.poisson(5.0, n), 0, m)
    deg[rng.random(n) < 0.15] = 1   # many 1-entry columns: padded to 2 when a 2-entry column exists
    deg[rng.random(n) < 0.03] = 0
    deg[:4] = (33, 40, 17, 1)        # a few long columns (generic path) sharing buckets
    ccol = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(deg, out=ccol[1:])
    row = np.concatenate([np.sort(rng.choice(m, size=d, replace=False)) for d in deg]).astype(np.int64)
    E = row.size
    # values chosen so that roughly half of the columns have a clamped sum below z and half above
    c = (-rng.random(E) * 0.02).astype(np.float32)
    a = (rng.random(E) + 0.1).astype(np.float32)
    lam = (rng.random(m) * 0.02).astype(np.float32)
    b = np.full(m, 30.0, dtype=np.float32)
    out = dict(ccol=ccol, row=row, a=a, c=c, b=b, lam=lam, gamma=np.float64(gamma), n_rows=np.int64(m),
               proj_type=np.array("simplex_eq"), proj_keys=np.array(["z"]), proj_vals=np.array([1.0]))
    for batching in (True, False):
        r = _calc(ccol, row, a, c, m, create_projection_map("simplex_eq", {"z": 1.0}, n), b, lam, gamma, batching)
        tag = "b1" if batching else "b0"
        out[f"x_{tag}"] = r.primal_var.numpy().copy()
        out[f"grad_{tag}"] = r.dual_gradient.numpy().copy()
        out[f"scal_{tag}"] = np.array([float(r.dual_objective), float(r.reg_penalty), float(r.primal_objective),
                                       float(r.dual_val_times_grad), float(r.max_pos_slack), float(r.sum_pos_slack)])
    x1, x0 = out["x_b1"], out["x_b0"]
    print("simplex_eq: nnz", E, "entries where batching on/off differ:", int((x1 != x0).sum()))
    np.savez_compressed(os.path.join(HERE, "case_simplex_eq.npz"), **out)


if __name__ == "__main__":
    _ref()
    make_mixed()
    make_simplex_eq()
