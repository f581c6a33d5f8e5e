"""Golden vectors shaped like BASELINE.json's configurations, produced by the UNMODIFIED reference in the build container.

    python tests/golden/make_golden_configs.py        # needs /root/reference; writes next to this file

cfg1_movielens_shaped.npz   configs[0]: the MovieLens matching example (examples/movielens_matching/
    movies_lens_matching.py:49-116): a == 1, c = -rating in {0.5,...,5}, one budget per movie, heavy-tailed column
    lengths (mean ~140, two users beyond 1024 ratings).  ml-20m/ratings.csv is not in the container, so the ratings
    are drawn here; everything downstream (CSC assembly, objective, Maximizer) is the reference's.  Box [0,1]
    (BASELINE.json's wording) and simplex z=1 (the example's own map, :163); calculate() at a random lambda and a
    30-iteration run of AcceleratedGradientDescent.maximize.
cfg2_synthetic.npz          configs[1..3]: the reference's benchmark generator (benchmark/generate_synthetic_data.py
    generate_synthetic_matching_input_args, seed 42) at a size the CPU path finishes in seconds; simplex z=1,
    gamma=1e-3, benchmark step sizes (benchmark/config.py:17-18).  Stored: calculate() at lambda=0 and a random lambda,
    a 40-iteration maximize() from zero, the same after the reference's jacobi_precondition (configs[2]'s preprocessing,
    benchmark_utils.py:54-56), and a warm-started second run through run_solver(initial_dual_path=...) (configs[3]).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _import_reference  # noqa: E402


def _calc_outputs(obj, lam_t, tag, out):
    r = obj.calculate(lam_t, save_primal=True)
    out[f"grad_{tag}"] = r.dual_gradient.numpy().copy()
    out[f"x_{tag}"] = r.primal_var.numpy().copy()
    out[f"scal_{tag}"] = np.array(
        [float(r.dual_objective), float(r.reg_penalty), float(r.primal_objective), float(r.dual_val_times_grad),
         float(r.max_pos_slack), float(r.sum_pos_slack)], dtype=np.float64)


def make_movielens_shaped(rng):
    import torch
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.optimizers.agd import AcceleratedGradientDescent
    from dualip.projections.base import create_projection_map

    n_users, n_movies = 500, 3000
    deg = np.clip(rng.lognormal(4.3, 1.0, n_users).astype(np.int64), 20, 1000)  # ml-20m: min 20 ratings per user
    deg[7], deg[311] = 1500, 2600  # ml-20m has users with thousands of ratings: the > 1024 path
    deg[100] = 0
    popularity = rng.lognormal(0, 1.2, n_movies)
    popularity /= popularity.sum()
    ccol = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(deg, out=ccol[1:])
    row = np.concatenate([np.sort(rng.choice(n_movies, size=d, replace=False, p=popularity)) for d in deg]).astype(np.int64)
    E = row.size
    rating = rng.choice(np.arange(0.5, 5.01, 0.5), size=E, p=[.01, .03, .02, .07, .05, .2, .12, .28, .08, .14])
    cval = (-rating).astype(np.float32)  # c = -(scale * rating + shift), scale 1, shift 0 (:81-82)
    aval = np.ones(E, dtype=np.float32)
    b = np.full(n_movies, 0.05, dtype=np.float32)  # per-movie budget scaled with the user count so that rows bind
    lam = (rng.random(n_movies) * 4.0).astype(np.float32)
    gamma = 0.1  # the example's default (:232)
    A = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(aval), size=(n_movies, n_users))
    C = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(cval), size=(n_movies, n_users))
    out = dict(ccol=ccol, row=row, a=aval, c=cval, b=b, lam=lam, gamma=np.float64(gamma), n_rows=np.int64(n_movies))
    for tag, (ptype, pparams) in (("box", ("box", {"lower": 0.0, "upper": 1.0})), ("simplex", ("simplex", {"z": 1}))):
        pm = create_projection_map(ptype, dict(pparams), n_users)
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(b)), gamma=gamma)
        _calc_outputs(obj, torch.from_numpy(lam), tag, out)
        solver = AcceleratedGradientDescent(max_iter=30, gamma=gamma, initial_step_size=1e-3, max_step_size=1e-1,
                                            iteration_callback=lambda i, r: None)
        res = solver.maximize(obj, torch.zeros(n_movies))
        out[f"{tag}_obj_log"] = np.array(res.dual_objective_log, dtype=np.float64)
        out[f"{tag}_step_log"] = np.array(res.step_size_log, dtype=np.float64)
        out[f"{tag}_dual"] = res.dual_val.numpy().copy()
    return out


def make_synthetic(rng):
    import torch

    sys.path.insert(0, REF)
    from benchmark.generate_synthetic_data import generate_synthetic_matching_input_args
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.optimizers.agd import AcceleratedGradientDescent
    from dualip.preprocessing.precondition import jacobi_precondition
    from dualip.run_solver import run_solver
    from dualip.types import ComputeArgs, ObjectiveArgs, SolverArgs

    n, m, sparsity, gamma = 20000, 200, 0.05, 1e-3
    with tempfile.TemporaryDirectory() as tmp:
        args = generate_synthetic_matching_input_args(num_sources=n, num_destinations=m, target_sparsity=sparsity,
                                                      device="cpu", dtype=torch.float32, seed=42, cache_dir=tmp)
        cache_files = {}
        for f in sorted(os.listdir(tmp)):
            if f.endswith("_meta.json"):
                cache_files["meta_name"] = np.array(f)
                cache_files["meta_json"] = np.array(open(os.path.join(tmp, f)).read())
    A, C, b = args.A, args.c, args.b_vec
    out = dict(ccol=A.ccol_indices().numpy().copy(), row=A.row_indices().numpy().copy(), a=A.values().numpy().copy(),
               c=C.values().numpy().copy(), b=b.numpy().copy(), gamma=np.float64(gamma), n_rows=np.int64(m), **cache_files)
    lam = (rng.random(m) * 0.02).astype(np.float32)
    out["lam"] = lam
    for batching, btag in ((True, "b1"), (False, "b0")):  # benchmark/config.py:22 runs with batching off
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, args.projection_map, b), gamma=gamma, batching=batching)
        _calc_outputs(obj, torch.zeros(m), f"zero_{btag}", out)
        _calc_outputs(obj, torch.from_numpy(lam), f"rand_{btag}", out)

    def solve(a_mat, b_vec, name, start=None, iters=40):
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(a_mat, C, args.projection_map, b_vec), gamma=gamma, batching=False)
        solver = AcceleratedGradientDescent(max_iter=iters, gamma=gamma, initial_step_size=1e-3, max_step_size=1e-1,
                                            iteration_callback=lambda i, r: None)
        res = solver.maximize(obj, torch.zeros(m) if start is None else start)
        out[f"{name}_obj_log"] = np.array(res.dual_objective_log, dtype=np.float64)
        out[f"{name}_step_log"] = np.array(res.step_size_log, dtype=np.float64)
        out[f"{name}_dual"] = res.dual_val.numpy().copy()
        return res

    first = solve(A, b, "plain")
    # configs[3]: warm start through run_solver(initial_dual_path=...) (run_solver.py:121-126)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "dual.pt")
        torch.save(first.dual_val.clone(), path)
        res = run_solver(
            input_args=MatchingInputArgs(A, C, args.projection_map, b),
            solver_args=SolverArgs(max_iter=20, gamma=gamma, initial_step_size=1e-3, max_step_size=1e-1, initial_dual_path=path),
            compute_args=ComputeArgs(host_device="cpu"), objective_args=ObjectiveArgs(objective_type="matching"))
        out["warm_obj_log"] = np.array(res.dual_objective_log, dtype=np.float64)
        out["warm_step_log"] = np.array(res.step_size_log, dtype=np.float64)
        out["warm_dual"] = res.dual_val.numpy().copy()
    # configs[2]: Jacobi row scaling changes A's values and b in place (precondition.py:8-29)
    A2 = torch.sparse_csc_tensor(A.ccol_indices().clone(), A.row_indices().clone(), A.values().clone(), size=A.shape)
    b2 = b.clone()
    norms = jacobi_precondition(A2, b2)
    out["jacobi_norms"] = norms.numpy().copy()
    out["jacobi_a"] = A2.values().numpy().copy()
    out["jacobi_b"] = b2.numpy().copy()
    solve(A2, b2, "jacobi")
    return out


def make_bisection(rng):
    """`method="bisection_search"` (projections/simplex.py:6-123): the operator on padded blocks, and through the reference
    objective on a small matching problem (simplex, z = 1; batching on and off: the result depends on the padding)."""
    import torch
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.projections.base import create_projection_map, project
    from make_golden import random_csc

    out = {}
    for L, K, scale in ((1, 40, 3.0), (2, 150, 2.0), (7, 200, 1.0), (16, 200, 0.5), (33, 80, 0.3)):
        x = ((rng.standard_normal((L, K)) + 0.3) * scale).astype(np.float32)
        lens = rng.integers(1, L + 1, size=K)
        for j in range(K):
            x[lens[j]:, j] = 0.0
        out[f"L{L}_x"] = x
        for name in ("simplex", "simplex_eq"):
            for z in (1.0, 2.5):
                out[f"L{L}_{name}_z{z}"] = project(name, z=z, method="bisection_search")(torch.from_numpy(x.copy())).numpy().copy()
    n_cols, n_rows, gamma = 1500, 48, 5e-2
    ccol, row = random_csc(rng, n_cols, n_rows, 7.0)
    E = row.size
    cval = (-np.minimum(rng.lognormal(-4.0, 0.75, E) * rng.lognormal(0, 0.7, E), 0.5) * 10.0).astype(np.float32)
    aval = (rng.lognormal(0, 1, E) * (-cval)).astype(np.float32)
    b = (rng.uniform(0.5, 1.0, n_rows) * 0.05 * n_cols / n_rows).astype(np.float32)
    lam = rng.random(n_rows).astype(np.float32)
    A = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(aval), size=(n_rows, n_cols))
    C = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(cval), size=(n_rows, n_cols))
    pm = create_projection_map("simplex", {"z": 1.0, "method": "bisection_search"}, n_cols)
    out.update(ccol=ccol, row=row, a=aval, c=cval, b=b, lam=lam, gamma=np.float64(gamma), n_rows=np.int64(n_rows))
    for batching, tag in ((True, "b1"), (False, "b0")):
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(b)), gamma=gamma, batching=batching)
        _calc_outputs(obj, torch.from_numpy(lam), tag, out)
    return out


def main():
    _import_reference()
    rng = np.random.default_rng(20261017)
    d = make_movielens_shaped(rng)
    np.savez_compressed(os.path.join(HERE, "cfg1_movielens_shaped.npz"), **d)
    print("cfg1 nnz", d["row"].size, "box obj", d["box_obj_log"][[0, -1]], "simplex obj", d["simplex_obj_log"][[0, -1]])
    d = make_synthetic(rng)
    np.savez_compressed(os.path.join(HERE, "cfg2_synthetic.npz"), **d)
    db = make_bisection(np.random.default_rng(20261018))
    np.savez_compressed(os.path.join(HERE, "projection_bisection.npz"), **db)
    print("bisection: nnz", db["row"].size, "scal", db["scal_b1"][:2], db["scal_b0"][:2])
    print("cfg2 nnz", d["row"].size, "plain", d["plain_obj_log"][[0, -1]], "warm", d["warm_obj_log"][[0, -1]],
          "jacobi", d["jacobi_obj_log"][[0, -1]])


if __name__ == "__main__":
    main()
