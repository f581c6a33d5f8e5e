"""Generates tests/golden/*.npz by running the UNMODIFIED reference (linkedin/DuaLip, /root/reference/src)
in the build container.  The GPU box has no /root/reference, so the vectors are committed.

    python tests/golden/make_golden.py            # needs /root/reference; writes next to this file

Each case stores the inputs (CSC arrays, b, lambda, gamma, projection spec) and the reference outputs of
MatchingSolverDualObjectiveFunction.calculate (dual_gradient, dual_objective, reg_penalty, primal_var, slacks),
plus AGD traces from AcceleratedGradientDescent.maximize.  Mixed projection maps are NOT generated through the
reference objective (its apply_F_to_columns corrupts earlier entries, see oracle/dualip_oracle.py docstring);
for those the reference projections are applied per entry on the reference's own `v` values.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("DUALIP_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    stub = types.ModuleType("mlflow")  # the reference imports mlflow unconditionally; it is not installed here

    def _noop_attr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None

    stub.__getattr__ = _noop_attr
    sys.modules.setdefault("mlflow", stub)
    sys.path.insert(0, os.path.join(REF, "src"))
    import torch  # noqa: F401
    import dualip  # noqa: F401


def random_csc(rng, n_cols, n_rows, mean_deg, max_deg=None, empty_frac=0.05):
    deg = rng.poisson(mean_deg, size=n_cols)
    deg = np.minimum(deg, n_rows if max_deg is None else min(max_deg, n_rows))
    deg[rng.random(n_cols) < empty_frac] = 0
    ccol = np.zeros(n_cols + 1, dtype=np.int64)
    np.cumsum(deg, out=ccol[1:])
    row = np.concatenate([np.sort(rng.choice(n_rows, size=d, replace=False)) for d in deg] + [np.zeros(0, dtype=np.int64)])
    return ccol, row.astype(np.int64)


def make_case(rng, n_cols, n_rows, mean_deg, gamma, lam_scale, proj, max_deg=None, scale_c=1.0):
    import torch
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.projections.base import create_projection_map

    ccol, row = random_csc(rng, n_cols, n_rows, mean_deg, max_deg)
    E = row.size
    cval = (-np.minimum(rng.lognormal(-4.0, 0.75, E) * rng.lognormal(0, 0.7, E), 0.5) * scale_c).astype(np.float32)
    aval = (rng.lognormal(0, 1, E) * (-cval)).astype(np.float32)
    b = (rng.uniform(0.5, 1.0, n_rows) * 0.05 * n_cols / n_rows).astype(np.float32)
    lam = (rng.random(n_rows) * lam_scale).astype(np.float32)
    A = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(aval), size=(n_rows, n_cols))
    C = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(cval), size=(n_rows, n_cols))
    ptype, pparams = proj
    pm = create_projection_map(ptype, dict(pparams), n_cols)
    out = {}
    for batching in (True, False):
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(b)), gamma=gamma, batching=batching)
        r = obj.calculate(torch.from_numpy(lam), save_primal=True)
        tag = "b1" if batching else "b0"
        out[f"grad_{tag}"] = r.dual_gradient.numpy().copy()
        out[f"x_{tag}"] = r.primal_var.numpy().copy()
        out[f"scal_{tag}"] = np.array(
            [float(r.dual_objective), float(r.reg_penalty), float(r.primal_objective), float(r.dual_val_times_grad),
             float(r.max_pos_slack), float(r.sum_pos_slack)], dtype=np.float64)
    out.update(ccol=ccol, row=row, a=aval, c=cval, b=b, lam=lam, gamma=np.float64(gamma), n_rows=np.int64(n_rows),
               proj_type=np.array(ptype), proj_keys=np.array(sorted(pparams.keys())),
               proj_vals=np.array([float(pparams[k]) for k in sorted(pparams.keys())], dtype=np.float64))
    return out


def make_agd_trace(rng):
    """30 iterations of the reference maximizer on a small random simplex LP, with step gamma decay."""
    import torch
    from dualip.objectives.matching import MatchingInputArgs, MatchingSolverDualObjectiveFunction
    from dualip.optimizers.agd import AcceleratedGradientDescent
    from dualip.projections.base import create_projection_map

    n_cols, n_rows = 400, 24
    ccol, row = random_csc(rng, n_cols, n_rows, 5.0)
    E = row.size
    cval = (-rng.random(E)).astype(np.float32)
    aval = (rng.random(E) + 0.1).astype(np.float32)
    b = np.full(n_rows, 3.0, dtype=np.float32)
    A = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(aval), size=(n_rows, n_cols))
    C = torch.sparse_csc_tensor(torch.from_numpy(ccol), torch.from_numpy(row), torch.from_numpy(cval), size=(n_rows, n_cols))
    pm = create_projection_map("simplex", {"z": 1.0}, n_cols)
    out = dict(ccol=ccol, row=row, a=aval, c=cval, b=b, n_rows=np.int64(n_rows))
    for name, kw in (("plain", {}), ("decay", dict(gamma_decay_type="step", gamma_decay_params={"decay_steps": 8, "decay_factor": 0.5}))):
        obj = MatchingSolverDualObjectiveFunction(MatchingInputArgs(A, C, pm, torch.from_numpy(b)), gamma=1e-2)
        solver = AcceleratedGradientDescent(max_iter=40, gamma=1e-2, initial_step_size=1e-3, max_step_size=0.1,
                                            iteration_callback=lambda i, r: None, **kw)
        res = solver.maximize(obj, torch.zeros(n_rows))
        out[f"{name}_obj_log"] = np.array(res.dual_objective_log, dtype=np.float64)
        out[f"{name}_step_log"] = np.array(res.step_size_log, dtype=np.float64)
        out[f"{name}_dual"] = res.dual_val.numpy().copy()
    return out


def make_projection_vectors(rng):
    """Reference projections applied to padded blocks (the ProjectionOperator contract, projections/base.py:30-36)."""
    import torch
    from dualip.projections.base import project

    out = {}
    for L, K, scale in ((1, 50, 3.0), (2, 200, 2.0), (7, 300, 1.0), (16, 300, 0.5), (33, 100, 0.3), (150, 40, 0.05)):
        x = ((rng.standard_normal((L, K)) + 0.3) * scale).astype(np.float32)
        # zero padding at the bottom of random columns, as apply_F_to_columns builds it
        lens = rng.integers(1, L + 1, size=K)
        for j in range(K):
            x[lens[j]:, j] = 0.0
        key = f"L{L}"
        out[f"{key}_x"] = x
        for name, params in (("simplex", {"z": 1.0}), ("simplex", {"z": 2.5}), ("simplex_eq", {"z": 1.0}),
                             ("box", {"lower": 0.0, "upper": 1.0}), ("box", {"lower": -0.5, "upper": 0.25}),
                             ("cone", {"lower": 0.0}), ("cone", {"upper": 0.1}), ("cone", {})):
            tag = name + "".join(f"_{k}{v}" for k, v in sorted(params.items()))
            out[f"{key}_{tag}"] = project(name, **params)(torch.from_numpy(x.copy())).numpy().copy()
    return out


def main():
    _import_reference()
    rng = np.random.default_rng(20260117)
    cases = {
        "simplex_small": make_case(rng, 2000, 50, 8.0, 1e-3, 0.05, ("simplex", {"z": 1.0})),
        "simplex_tight": make_case(rng, 3000, 40, 10.0, 1e-1, 2.0, ("simplex", {"z": 1.0}), scale_c=20.0),
        "simplex_z2": make_case(rng, 1500, 64, 6.0, 5e-2, 1.0, ("simplex", {"z": 2.0}), scale_c=10.0),
        "simplex_long": make_case(rng, 300, 400, 90.0, 5e-2, 1.0, ("simplex", {"z": 1.0}), scale_c=10.0),
        "box": make_case(rng, 2000, 50, 8.0, 1e-2, 2.0, ("box", {"lower": 0.0, "upper": 1.0}), scale_c=0.5),
        "cone_lower": make_case(rng, 1000, 30, 5.0, 1e-2, 4.0, ("cone", {"lower": 0.0})),
    }
    for name, d in cases.items():
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **d)
        print(name, "nnz", d["row"].size)
    np.savez_compressed(os.path.join(HERE, "agd_trace.npz"), **make_agd_trace(rng))
    np.savez_compressed(os.path.join(HERE, "projection_vectors.npz"), **make_projection_vectors(rng))


if __name__ == "__main__":
    main()
